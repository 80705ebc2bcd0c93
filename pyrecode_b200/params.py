"""Run parameters of the ReCoDe API (drop-in for pyrecode/params.py).

InitParams   run-level keyword arguments (pyrecode/params.py:7-190)
InputParams  the `key = int` text file, its validation and dtype inference (pyrecode/params.py:193-569,
             keys documented in config/README.md:8-30 of the reference)
"""
from .misc import map_dtype


class InitParams:

    def __init__(self, mode, output_directory, image_filename='', directory_path='', calibration_filename='',
                 params_filename='', validation_frame_gap=-1, log_filename='recode.log', run_name='run',
                 verbosity=0, use_c=False, max_count=-1, chunk_time_in_sec=0):
        self._mode = mode.strip().lower()
        self._output_directory = output_directory
        self._image_filename = image_filename
        self._directory_path = directory_path
        self._calibration_filename = calibration_filename
        self._params_filename = params_filename
        self._validation_frame_gap = validation_frame_gap
        self._log_filename = log_filename
        self._run_name = run_name
        self._verbosity = verbosity
        self._use_c = use_c
        self._max_count = max_count
        self._chunk_time_in_sec = chunk_time_in_sec
        if not self._validate_init_params():
            self.show_usage()
            raise ValueError('Invalid initialization parameters')

    def validate(self):
        self._validate_init_params()

    def _validate_init_params(self):
        if self._output_directory == '':
            print('Output Directory cannot be empty')
            return False
        if self._mode not in ('batch', 'stream'):
            print("Unknown mode: mode can only be 'batch' or 'stream'")
            return False
        if self._mode == 'batch' and self._image_filename == '':
            print('Image filename cannot be empty')
            return False
        self._verbosity = min(max(self._verbosity, 0), 2)
        return True

    mode = property(lambda self: self._mode)
    verbosity = property(lambda self: self._verbosity)
    validation_frame_gap = property(lambda self: self._validation_frame_gap)
    image_filename = property(lambda self: self._image_filename)
    calibration_filename = property(lambda self: self._calibration_filename)
    params_filename = property(lambda self: self._params_filename)
    output_directory = property(lambda self: self._output_directory)
    log_filename = property(lambda self: self._log_filename)
    run_name = property(lambda self: self._run_name)
    use_c = property(lambda self: self._use_c)
    directory_path = property(lambda self: self._directory_path)
    max_count = property(lambda self: self._max_count)
    chunk_time_in_sec = property(lambda self: self._chunk_time_in_sec)

    @staticmethod
    def show_usage():
        print("See documentation at https://github.com/NDLOHGRP/pyReCoDe for usage details")


_KEYS = ('reduction_level', 'rc_operation_mode', 'calibration_threshold_epsilon', 'target_bit_depth',
         'source_bit_depth', 'num_cols', 'num_rows', 'num_frames', 'frame_offset', 'num_calibration_frames',
         'calibration_frame_offset', 'keep_part_files', 'num_threads', 'l2_statistics', 'l4_centroiding',
         'compression_scheme', 'compression_level', 'source_file_type', 'source_header_length',
         'keep_calibration_data', 'calibration_file_type', 'source_data_type', 'target_data_type',
         'source_numpy_dtype', 'target_numpy_dtype')

# (key, allowed values, message) checked in this order, like pyrecode/params.py:227-341
_CHOICES = (
    ('reduction_level', (1, 2, 3, 4), 'Reduction level must be 1, 2, 3 or 4'),
    ('rc_operation_mode', (0, 1), 'RC Operation mode can be 0, 1 or 2'),
)
_CHOICES_LATE = (
    ('keep_part_files', (0, 1), 'Keep part files must be 0 or 1'),
    ('l2_statistics', (0, 1, 2), 'L2 statistics must be 0, 1 or 2'),
    ('l4_centroiding', (0, 1, 2, 3), 'L4 centroiding must be 0, 1, 2 or 3'),
    ('compression_scheme', tuple(range(12)), 'Compression scheme must be 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10 or 11'),
)


class InputParams:

    def __init__(self):
        self._param_map = {k: -1 for k in _KEYS}

    def load(self, params_filename):
        assert params_filename != '', 'Params filename missing'
        with open(params_filename) as fp:
            for line in fp:
                if line == '' or line == '\n' or line.startswith('#'):
                    continue
                parts = line.split('=')
                key = parts[0].strip().lower()
                assert key in self._param_map, 'Unknown parameter: ' + key
                self._param_map[key] = int(parts[1].strip().lower())

    def _validate_input_params(self):
        p = self._param_map
        for key, allowed, msg in _CHOICES:
            if p[key] not in allowed:
                print(msg)
                return False
        if p['calibration_threshold_epsilon'] == '':
            print('Calibration threshold (epsilon) cannot be empty')
            return False
        binary_source = p['source_file_type'] in (0, 3)
        for key, what in (('source_bit_depth', 'Source bit depth'), ('num_cols', 'Number of columns'),
                          ('num_rows', 'Number of rows'), ('num_frames', 'Number of frames')):
            if p[key] == -1 and binary_source:
                print(what + ' cannot be empty when source filetype is binary/other')
                return False
        for key, what in (('frame_offset', 'Frame offset'), ('num_calibration_frames', 'Number of calibration'),
                          ('calibration_frame_offset', 'Calibration frame offset')):
            if not isinstance(p[key], int):
                print(what + ' should be an integer')
                return False
        if p['keep_part_files'] not in (0, 1):
            print('Keep part files must be 0 or 1')
            return False
        if not isinstance(p['num_threads'], int):
            print('Number of threads should be an integer')
            return False
        for key, allowed, msg in _CHOICES_LATE[1:]:
            if p[key] not in allowed:
                print(msg)
                return False
        if int(p['compression_level']) < 0 or int(p['compression_level']) > 22:
            print('Compression level can be from 0 - 22')
            return False
        if p['keep_calibration_data'] not in (0, 1):
            print('Keep dark data cannot be either 0 or 1')
            return False
        if p['source_file_type'] not in (0, 1, 2, 3):
            print('Source file type must be 0, 1, 2 or 3')
            return False
        if binary_source and (p['source_header_length'] == -1 or not isinstance(p['source_header_length'], int)):
            print('Source Header Length cannot be empty or non-integer when source filetype is binary/other')
            return False
        if p['calibration_file_type'] not in (0, 1, 2, 3):
            print('Calibration filetype must be 0, 1, 2 or 3')
            return False
        if p['frame_offset'] < 0:
            p['frame_offset'] = 0
        if p['num_threads'] < 1:
            p['num_threads'] = 1
        if p['source_data_type'] not in (0, 1, 2):
            print('Source data type must be 0, 1, or 2')
            return False
        if p['target_data_type'] not in (0, 1, 2):
            print('Target data type must be 0, 1, or 2')
            return False
        if p['target_bit_depth'] == -1:
            p['target_bit_depth'] = p['source_bit_depth']
        p['source_numpy_dtype'] = map_dtype(p['source_data_type'], p['source_bit_depth'])
        p['target_numpy_dtype'] = map_dtype(p['target_data_type'], p['target_bit_depth'])
        return True

    def validate(self):
        return self._validate_input_params()

    def serialize(self, filename):
        with open(filename, 'w') as f:
            for key in self._param_map:
                f.write(key + ' = ' + str(self._param_map[key]) + '\n')


def _prop(key, settable=True):
    def getter(self):
        return self._param_map[key]

    def setter(self, value):
        self._param_map[key] = value
    return property(getter, setter if settable else None)


# property name -> key (pyrecode/params.py:348-569); the reference exposes setters for a few of them, the
# superset here is harmless
for _name, _key in (('reduction_level', 'reduction_level'), ('rc_operation_mode', 'rc_operation_mode'),
                    ('calibration_threshold_epsilon', 'calibration_threshold_epsilon'),
                    ('target_bit_depth', 'target_bit_depth'), ('source_bit_depth', 'source_bit_depth'),
                    ('num_cols', 'num_cols'), ('num_rows', 'num_rows'), ('num_frames', 'num_frames'),
                    ('nx', 'num_cols'), ('ny', 'num_rows'), ('nz', 'num_frames'), ('frame_offset', 'frame_offset'),
                    ('num_calibration_frames', 'num_calibration_frames'),
                    ('calibration_frame_offset', 'calibration_frame_offset'), ('keep_part_files', 'keep_part_files'),
                    ('num_threads', 'num_threads'), ('l2_statistics', 'l2_statistics'),
                    ('l4_centroiding', 'l4_centroiding'), ('L2_statistics', 'l2_statistics'),
                    ('L4_centroiding', 'l4_centroiding'), ('compression_scheme', 'compression_scheme'),
                    ('compression_level', 'compression_level'), ('keep_calibration_data', 'keep_calibration_data'),
                    ('source_file_type', 'source_file_type'), ('source_header_length', 'source_header_length'),
                    ('calibration_file_type', 'calibration_file_type'), ('source_data_type', 'source_data_type'),
                    ('target_data_type', 'target_data_type')):
    setattr(InputParams, _name, _prop(_key))
for _name in ('source_numpy_dtype', 'target_numpy_dtype'):
    setattr(InputParams, _name, _prop(_name, settable=False))
