"""The ReCoDe file header (drop-in for pyrecode/recode_header.py): 512-byte v0.2 layout, 321-byte v0.1 layout
(read only), little endian.  Byte-compatible with the reference.
"""
import numpy as np

_U8, _U16, _U32, _U64 = np.uint8, np.uint16, np.uint32, np.uint64

# (name, bytes, dtype) in file order
_FIELDS_V02 = (
    ('uid', 8, _U64), ('version_major', 1, _U8), ('version_minor', 1, _U8), ('is_intermediate', 1, _U8),
    ('reduction_level', 1, _U8), ('rc_operation_mode', 1, _U8), ('is_bit_packed', 1, _U8),
    ('target_bit_depth', 1, _U8), ('nx', 4, _U32), ('ny', 4, _U32), ('nz', 4, _U32),
    ('frame_metadata_size', 1, _U8), ('num_non_standard_frame_metadata', 1, _U8), ('L2_statistics', 1, _U8),
    ('L4_centroiding', 1, _U8), ('compression_scheme', 1, _U8), ('compression_level', 1, _U8),
    ('source_file_type', 1, _U8), ('source_header_length', 2, _U16), ('source_header_position', 1, _U8),
    ('source_file_name', 100, _U8), ('calibration_file_name', 100, _U8),
    ('calibration_threshold_epsilon', 8, _U64), ('has_calibration_data', 1, _U8), ('frame_offset', 4, _U32),
    ('calibration_frame_offset', 4, _U32), ('num_calibration_frames', 4, _U32), ('source_bit_depth', 1, _U8),
    ('source_dtype', 1, _U8), ('target_dtype', 1, _U8), ('checksum', 32, _U8), ('futures', 219, _U8))

_FIELDS_V01 = (
    ('uid', 8, _U64), ('version_major', 1, _U8), ('version_minor', 1, _U8), ('reduction_level', 1, _U8),
    ('rc_operation_mode', 1, _U8), ('target_bit_depth', 1, _U8), ('nx', 2, _U16), ('ny', 2, _U16), ('nz', 4, _U32),
    ('L2_statistics', 1, _U8), ('L4_centroiding', 1, _U8), ('compression_scheme', 1, _U8),
    ('compression_level', 1, _U8), ('source_file_type', 1, _U8), ('source_header_length', 2, _U16),
    ('source_header_position', 1, _U8), ('source_file_name', 100, _U8), ('calibration_file_name', 100, _U8),
    ('calibration_threshold_epsilon', 2, _U16), ('has_calibration_data', 1, _U8), ('frame_offset', 4, _U32),
    ('calibration_frame_offset', 4, _U32), ('num_calibration_frames', 4, _U32), ('source_bit_depth', 1, _U8),
    ('source_dtype', 1, _U8), ('target_dtype', 1, _U8), ('checksum', 32, _U8), ('futures', 42, _U8))

_NAME_FIELDS = ('source_file_name', 'calibration_file_name')
UID = 158966344846346


class ReCoDeHeader:

    def __init__(self, version=0.2):
        self._version = version
        self._rc_header = {}
        self._rc_header_field_defs = []
        self._rc_header_length = 0
        self._get_rc_field_defs()
        self._source_header = None
        self._non_standard_frame_metadata_sizes = {}

    def _get_rc_field_defs(self):
        table = _FIELDS_V01 if self._version < 0.2 else _FIELDS_V02
        self._rc_header_field_defs = [{'name': n, 'bytes': b, 'dtype': d} for n, b, d in table]
        self._rc_header_length = sum(b for _, b, _ in table)

    def create(self, init_params, input_params, is_intermediate):
        """fills the header from the run parameters (pyrecode/recode_header.py:96-163)"""
        ip, h = input_params, self._rc_header
        old = self._version < 0.2
        h['uid'] = UID
        h['version_major'] = 0
        h['version_minor'] = 1 if old else 2
        if not old:
            h['is_intermediate'] = is_intermediate
            h['is_bit_packed'] = 1
            h['frame_metadata_size'] = 0
            h['num_non_standard_frame_metadata'] = 0
        h['reduction_level'] = ip.reduction_level
        h['rc_operation_mode'] = ip.rc_operation_mode
        h['target_bit_depth'] = ip.target_bit_depth
        h['nx'], h['ny'], h['nz'] = ip.nx, ip.ny, ip.nz
        h['L2_statistics'] = ip.L2_statistics
        h['L4_centroiding'] = ip.L4_centroiding
        h['compression_scheme'] = ip.compression_scheme
        h['compression_level'] = ip.compression_level
        h['source_file_type'] = ip.source_file_type
        h['source_header_length'] = ip.source_header_length
        h['source_header_position'] = 0
        h['source_file_name'] = init_params.image_filename
        h['calibration_file_name'] = init_params.calibration_filename
        h['calibration_threshold_epsilon'] = ip.calibration_threshold_epsilon
        h['has_calibration_data'] = ip.keep_calibration_data
        h['frame_offset'] = ip.frame_offset
        h['calibration_frame_offset'] = ip.calibration_frame_offset
        h['num_calibration_frames'] = ip.num_calibration_frames
        h['source_bit_depth'] = ip.source_bit_depth
        h['source_dtype'] = 0 if old else ip.source_data_type      # v0.1 only knows unsigned ints
        h['target_dtype'] = 0 if old else ip.target_data_type
        h['checksum'] = np.zeros(32, dtype=np.uint8)
        h['futures'] = np.zeros(42 if old else 219, dtype=np.uint8)

    @property
    def recode_header_length(self):
        return self._rc_header_length

    def as_dict(self):
        return self._rc_header

    def get(self, field_name):
        if field_name not in self._rc_header:
            raise ValueError('The requested field does not exist in recode header')
        return self._rc_header[field_name]

    def set(self, field_name, value):
        if field_name not in self._rc_header:
            raise ValueError('The requested field does not exist in recode header')
        self._rc_header[field_name] = value

    def update(self, name, value):
        self._rc_header[name] = value

    def get_definition(self, name):
        for field_def in self._rc_header_field_defs:
            if field_def['name'] == name:
                return field_def
        raise ValueError('The requested field does not exist in recode header')

    def load(self, rc_filename, is_intermediate=False):
        if rc_filename == '':
            raise ValueError('ReCoDe filename missing')
        with open(rc_filename, 'rb') as fp:
            head = fp.read(10)
            if len(head) < 10:
                raise ValueError('File too short to hold a ReCoDe header')
            self._version = int(head[8]) + int(head[9]) / 10.0
            self._get_rc_field_defs()
            fp.seek(0, 0)
            for field in self._rc_header_field_defs:
                raw = fp.read(field['bytes'])
                if field['name'] in _NAME_FIELDS:
                    value = ''.join(chr(x) for x in raw)
                elif field['dtype'] == np.uint8 and field['bytes'] != 1:
                    value = np.frombuffer(raw, dtype=np.uint8)
                else:
                    value = int.from_bytes(raw, 'little')
                self._rc_header[field['name']] = value
            if self._version < 0.2:
                self._rc_header['is_intermediate'] = 0 if is_intermediate else 1
                self._rc_header['is_bit_packed'] = 1
                self._rc_header['frame_metadata_size'] = 0
                self._rc_header['num_non_standard_frame_metadata'] = 0
                self._rc_header['source_header_length'] = 0
                self._rc_header['source_dtype'] = 0
                self._rc_header['target_dtype'] = 0
            for _ in range(self._rc_header['num_non_standard_frame_metadata']):
                raw = np.frombuffer(fp.read(100), dtype=np.uint8)
                self._non_standard_frame_metadata_sizes[''.join(chr(x) for x in raw[:-1])] = raw[99]
            self._source_header = fp.read(self._rc_header['source_header_length'])

    def serialize(self, rc_filename):
        if rc_filename == '':
            raise ValueError('ReCoDe filename missing')
        with open(rc_filename, 'wb') as fp:
            self.serialize_to(fp)

    def to_bytes(self):
        out = bytearray()
        for field in self._rc_header_field_defs:
            nb, value = field['bytes'], self._rc_header[field['name']]
            if field['name'] in _NAME_FIELDS:
                out += str(value)[:nb].ljust(nb, ' ').encode('utf-8')[:nb]
            elif field['dtype'] == np.uint8 and nb != 1:
                out += np.asarray(value, dtype=np.uint8)[:nb].tobytes().ljust(nb, b'\0')
            else:
                out += int(value).to_bytes(nb, 'little')
        return bytes(out)

    def serialize_to(self, fp):
        fp.write(self.to_bytes())

    def skip_header(self, rc_fp):
        rc_fp.seek(self._rc_header_length)
        return rc_fp

    def get_frame_data_offset(self, is_intermediate, sz_frame_metadata):
        """offset of frame 0's metadata (intermediate files) or data (merged files), recode_header.py:281-291"""
        if self._rc_header['version_major'] == 0 and self._rc_header['version_minor'] == 1:
            offset = self._rc_header_length
        else:
            offset = self._rc_header_length + self._rc_header['source_header_length'] \
                + len(self._non_standard_frame_metadata_sizes) * 100
        if is_intermediate:
            return offset
        return int(offset + self._rc_header['nz'] * sz_frame_metadata)

    @property
    def source_header(self):
        return self._source_header

    @property
    def non_standard_metadata_sizes(self):
        return self._non_standard_frame_metadata_sizes

    def get_field_position_in_bytes(self, name):
        position = 0
        for field_def in self._rc_header_field_defs:
            if field_def['name'] == name:
                return position
            position += field_def['bytes']
        raise ValueError('The requested field is not defined in the header')

    def print(self):
        print('ReCoDe Header')
        print('-------------')
        for field in self._rc_header_field_defs:
            print(field['name'], '=', self._rc_header[field['name']])

    def validate(self):
        for field in self._rc_header_field_defs:
            if field['name'] not in self._rc_header:
                print('ReCoDe Header Validation Failed: ' + field['name'] + ' is missing.')
                return False
        return True
