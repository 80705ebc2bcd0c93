"""ReCoDeReader / merge_parts -- drop-in for pyrecode/recode_reader.py with decompress + unpack on the GPU.

Same open() / get_next_frame() / get_frame(z) / get_next_frame_raw() protocol and return shapes as the
reference (pyrecode/recode_reader.py:15-493): a frame comes back as
    {frame_id: {'metadata': {...}, 'data': scipy.sparse.coo_matrix[, 'summary_stats': ndarray]}}.
The per-frame work of _get_frame_sparse (recode_reader.py:379-462: zlib.decompress x2 + c_recode.get_frame_sparse)
runs in librecode_b200 (rc_inflate_zlib + rc_unpack_sparse).  All reduction levels and both operation modes are
readable (the reference can only read L1, SURVEY B-2).  Batched extras for throughput: read_frames_dense() and
sum_frames() keep the decoded frames on the device.
"""
import os
import sys

import numpy as np
from scipy.sparse import coo_matrix

from .misc import map_dtype
from .recode_header import ReCoDeHeader
from .structures import ReCoDeStructures


class ReCoDeReader:

    def __init__(self, file, is_intermediate=False, device=None, batch_frames=16):
        self._source_filename = file
        self._current_frame_index = 0
        self._is_intermediate = 1 if is_intermediate else 0
        self._device = device
        self._batch_frames = batch_frames
        self._file_size = None
        self._header = None
        self._frame_metadata = None
        self._seek_table = None
        self._rc_header = None
        self._frame_data_start_position = 0
        self._sz_frame_metadata = None
        self._n_elements_frame_metadata = None
        self._fp = None
        self._structures = None
        self._numpy_dtype = None
        self._engine = None
        self._sm = None

    # ------------------------------------------------------------------------------------------
    def open(self, print_header=True):
        self._rc_header = ReCoDeHeader()
        self._rc_header.load(self._source_filename)
        self._header = self._rc_header.as_dict()
        if print_header:
            self._rc_header.print()
        if self._header['compression_scheme'] != 0:
            raise NotImplementedError('only compression_scheme 0 (zlib / deflate) is supported on the GPU path')
        self._fp = open(self._source_filename, 'rb')
        self._fp.seek(0, 2)
        self._file_size = self._fp.tell()
        self._fp.seek(0, 0)
        self._initialize()
        self._load_seek_table()
        self._numpy_dtype = map_dtype(self._header['target_dtype'], self._header['target_bit_depth'])

    def _initialize(self):
        h = self._header
        self._structures = ReCoDeStructures(h)
        self._sm = self._structures.standard_frame_metadata_structure_for(h['reduction_level'], h['rc_operation_mode'])
        nsm = list(self._rc_header.non_standard_metadata_sizes.values())
        self._sz_frame_metadata = self._structures.get_standard_frame_metadata_size(
            h['reduction_level'], h['rc_operation_mode']) + int(np.sum(nsm))
        self._n_elements_frame_metadata = len(nsm) + len(self._sm)
        self._frame_data_start_position = self._rc_header.get_frame_data_offset(self._is_intermediate,
                                                                                self._sz_frame_metadata)
        return h

    def _get_engine(self):
        if self._engine is None:
            from .engine import ReadEngine          # imports torch + the CUDA library; no CPU fallback
            h = self._header
            if h['target_dtype'] != 0 or not 1 <= h['target_bit_depth'] <= 16:
                raise NotImplementedError('only unsigned targets of 1..16 bits are supported on the GPU path')
            itemsize = 1 if h['target_bit_depth'] <= 8 else 2
            self._engine = ReadEngine(h['ny'], h['nx'], itemsize, h['target_bit_depth'], h['reduction_level'],
                                      h['rc_operation_mode'], max_frames=self._batch_frames, device=self._device)
        return self._engine

    def _load_seek_table(self):
        """merged files: per-frame metadata table right after the header; offsets = exclusive cumsum of the
        frame data sizes (recode_reader.py:127-168, vectorised)"""
        if self._is_intermediate:
            return
        h = self._header
        nf = len(self._sm)
        self._fp.seek(self._rc_header.get_frame_data_offset(True, self._sz_frame_metadata), 0)
        raw = self._fp.read(h['nz'] * nf * 4)
        table = np.frombuffer(raw, dtype='<u4').reshape(h['nz'], nf) if nf else np.zeros((h['nz'], 0), np.uint32)
        names = [f['name'] for f in self._sm]
        self._frame_metadata = [{n: table[z, i] for i, n in enumerate(names)} for z in range(h['nz'])]
        sizes = np.zeros(h['nz'], dtype=np.uint64)
        for i, f in enumerate(self._sm):
            if f['is_frame_size']:
                sizes += table[:, i]
        if h['rc_operation_mode'] == 0:
            sizes += np.uint64(self._structures.binary_image_sz_bytes)
        self._seek_table = np.zeros((h['nz'], 2), dtype=np.uint64)
        self._seek_table[:, 0] = sizes
        if h['nz'] > 1:
            self._seek_table[1:, 1] = np.cumsum(sizes[:-1])

    # ------------------------------------------------------------------------------------------
    def get_header(self):
        return self._rc_header

    def get_source_header(self):
        return self._rc_header.source_header

    def get_true_shape(self):
        return (self._header['nz'], self._header['ny'], self._header['nx'])

    def get_shape(self):
        return (self._header['nz'], self._header['ny'], self._header['nx'])

    def get_dtype(self):
        return self._header['target_dtype']

    def get_sub_volume(self, slice_z, slice_y, slice_x):
        raise NotImplementedError

    def close(self):
        self._fp.close()

    def seek_to_frame_data(self):
        self._frame_data_start_position = self._rc_header.get_frame_data_offset(self._is_intermediate,
                                                                                self._sz_frame_metadata)
        self._fp.seek(0, 2)
        if self._frame_data_start_position <= self._fp.tell():
            self._fp.seek(self._frame_data_start_position, 0)

    def get_file_position(self):
        return self._fp.tell()

    def copy_headers_to(self, target_fp, source_header_length):
        self._fp.seek(0, 0)
        target_fp.write(self._fp.read(self._rc_header.recode_header_length))
        target_fp.write(self._fp.read(source_header_length))

    @property
    def sz_frame_metadata(self):
        return self._sz_frame_metadata

    # ------------------------------------------------------------------------------------------
    def _read_intermediate_metadata(self):
        """-> (frame_id, metadata dict) of the record at the file position, or None at EOF"""
        head = self._fp.read(4)
        if len(head) < 4:
            return None
        frame_id = np.frombuffer(head, dtype=np.uint32)[0]
        d = {}
        for field in self._sm:
            d[field['name']] = np.frombuffer(self._fp.read(field['bytes']), dtype=field['dtype'])[0]
        return frame_id, d

    def _stream_sizes(self, md):
        """(map bytes, value bytes or None) of a frame in the file"""
        h = self._header
        level, mode = h['reduction_level'], h['rc_operation_mode']
        n_map = int(md['bytes_in_compressed_binary_map']) if mode == 1 else self._structures.binary_image_sz_bytes
        if level in (1, 2):
            s = 'pixvals' if level == 1 else 'summary_stats'
            return n_map, int(md[('bytes_in_compressed_' if mode == 1 else 'bytes_in_packed_') + s])
        return n_map, None

    def _get_frame_raw(self, frame_metadata, read_data=True):
        n_map, n_val = self._stream_sizes(frame_metadata)
        if read_data:
            out = {'binary_map': self._fp.read(n_map)}
            if n_val is not None:
                out['pixvals'] = self._fp.read(n_val)
            return out
        self._fp.seek(n_map + (n_val or 0), 1)
        return {'binary_map': None, 'pixvals': None} if n_val is not None else {'binary_map': None}

    def _decode(self, raws, mds):
        """decode a batch of raw frames -> list of (coo_matrix, summary_stats or None)"""
        h = self._header
        level, b = h['reduction_level'], h['target_bit_depth']
        eng = self._get_engine()
        out = []
        for i0 in range(0, len(raws), eng.max_frames):
            part = raws[i0:i0 + eng.max_frames]
            eng.load([r['binary_map'] for r in part], [r['pixvals'] for r in part] if level in (1, 2) else None)
            sizes = eng.check()
            tri = eng.sparse()
            stats = eng.summary_stats(sizes) if level == 2 else [None] * len(part)
            for j, t in enumerate(tri):
                if level == 1:
                    npk = int(mds[i0 + j]['bytes_in_packed_pixvals'])
                    if (t.shape[0] * b + 7) // 8 != npk:
                        raise ValueError('frame has %d foreground pixels but %d packed bytes' % (t.shape[0], npk))
                coo = coo_matrix((t[:, 2], (t[:, 0], t[:, 1])), shape=(h['ny'], h['nx']), dtype=self._numpy_dtype)
                out.append((coo, stats[j]))
        return out

    def _frame_dict(self, md, decoded):
        coo, stats = decoded
        if self._header['reduction_level'] == 2:
            return {'metadata': md, 'data': coo, 'summary_stats': stats}
        return {'metadata': md, 'data': coo}

    def get_frame(self, z):
        if self._is_intermediate:
            raise ValueError("Random acceess is not available for intermediate files")
        if z >= self._header['nz']:
            raise ValueError('Requested frame index is greater than number of frames in dataset')
        self._fp.seek(self._frame_data_start_position + int(self._seek_table[z, 1]), 0)
        if self._file_size - self._fp.tell() == 0:
            self._header['nz'] = self._current_frame_index
            return None
        md = self._frame_metadata[z]
        raw = self._get_frame_raw(md)
        d = self._decode([raw], [md])[0]
        self._current_frame_index = z + 1
        return {z: self._frame_dict(md, d)}

    def _next_raw(self, read_data=True):
        if self._current_frame_index == 0:
            self._fp.seek(self._frame_data_start_position, 0)
        if self._file_size - self._fp.tell() == 0:
            return None
        if self._is_intermediate:
            r = self._read_intermediate_metadata()
            if r is None:
                return None
            frame_id, md = r
        else:
            if self._current_frame_index >= self._header['nz']:
                raise ValueError('Requested frame index is greater than number of frames in dataset')
            frame_id = self._current_frame_index
            md = self._frame_metadata[frame_id]
        raw = self._get_frame_raw(md, read_data=read_data)
        return frame_id, md, raw

    def get_next_frame(self):
        r = self._next_raw()
        if r is None:
            return None
        frame_id, md, raw = r
        d = self._decode([raw], [md])[0]
        self._current_frame_index += 1
        return {frame_id: self._frame_dict(md, d)}

    def get_next_frame_raw(self, read_data=True):
        r = self._next_raw(read_data=read_data)
        if r is None:
            return None
        frame_id, md, raw = r
        self._current_frame_index += 1
        return {frame_id: {'metadata': md, 'data': raw if read_data else self._fp.tell()}}

    # ---- batched extras (device resident results) ------------------------------------------------
    def _next_batch_raw(self, n):
        ids, mds, raws = [], [], []
        while len(raws) < n:
            r = self._next_raw()
            if r is None:
                break
            ids.append(int(r[0]))
            mds.append(r[1])
            raws.append(r[2])
            self._current_frame_index += 1
        return ids, mds, raws

    def read_frames_dense(self, n):
        """next n frames -> (frame ids, CUDA tensor [k, ny, nx] of the target dtype); k <= n at EOF"""
        eng = self._get_engine()
        import torch
        ids, out = [], []
        while len(ids) < n:
            bi, mds, raws = self._next_batch_raw(min(eng.max_frames, n - len(ids)))
            if not raws:
                break
            eng.load([r['binary_map'] for r in raws], [r.get('pixvals') for r in raws]
                     if self._header['reduction_level'] in (1, 2) else None)
            eng.check()
            out.append(eng.dense())
            ids += bi
        if not out:
            return ids, None
        return ids, torch.cat(out, 0)

    def sum_frames(self, n, total=None):
        """live view: adds the next n frames into `total` (uint32 CUDA tensor [ny*nx], created if None) without
        materialising dense frames (examples/ReCoDe_Live_View_MT.ipynb cell 1) -> (frame ids, total)"""
        eng = self._get_engine()
        import torch
        if total is None:
            total = torch.zeros(self._header['ny'] * self._header['nx'], dtype=torch.int32, device=eng.dev)
        ids = []
        while len(ids) < n:
            bi, mds, raws = self._next_batch_raw(min(eng.max_frames, n - len(ids)))
            if not raws:
                break
            eng.load([r['binary_map'] for r in raws], [r.get('pixvals') for r in raws]
                     if self._header['reduction_level'] in (1, 2) else None)
            eng.check()
            eng.dense(total=total, want_dense=False)
            ids += bi
        return ids, total


def merge_parts(folder_path, base_filename, num_parts):
    """Merge <base>_part000.. into one random-access file: header (+ source header) of part 0, the nz x metadata
    table (without frame ids), then the frame payloads in ascending frame id (pyrecode/recode_reader.py:495-595).
    Pure host byte shuffling; part files are walked once."""
    parts = []
    for index in range(num_parts):
        name = os.path.join(folder_path, base_filename + '_part' + '{0:03d}'.format(index))
        reader = ReCoDeReader(name, is_intermediate=True)
        reader.open(print_header=False)
        parts.append(reader)
    first = parts[0]
    header = first.get_header()
    with open(os.path.join(folder_path, base_filename), 'wb') as target:
        first.copy_headers_to(target, header.as_dict()['source_header_length'])
        meta_start = target.tell()
        # index every record (frame id, metadata, file span) without touching payloads
        records = []
        for pi, reader in enumerate(parts):
            reader._current_frame_index = 0
            while True:
                r = reader._next_raw(read_data=False)
                if r is None:
                    break
                frame_id, md, _ = r
                end = reader._fp.tell()
                n_map, n_val = reader._stream_sizes(md)
                size = n_map + (n_val or 0)
                records.append((int(frame_id), pi, end - size, size, md))
                reader._current_frame_index += 1
        records.sort(key=lambda x: (x[0], x[1]))
        target.seek(meta_start + first.sz_frame_metadata * len(records), 0)
        names = [f['name'] for f in first._sm]
        table = np.zeros((len(records), len(names)), dtype='<u4')
        for i, (frame_id, pi, off, size, md) in enumerate(records):
            fp = parts[pi]._fp
            fp.seek(off, 0)
            target.write(fp.read(size))
            for j, n in enumerate(names):
                table[i, j] = md[n]
        target.seek(meta_start, 0)
        target.write(table.tobytes())
        target.seek(header.get_field_position_in_bytes('nz'), 0)
        target.write(len(records).to_bytes(header.get_definition('nz')['bytes'], sys.byteorder))
    for reader in parts:
        reader.close()
