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

    def __init__(self, file, is_intermediate=False, device=None, batch_frames=16, bulk_frames=64, bulk_inflight=6):
        self._source_filename = file
        self._current_frame_index = 0
        self._is_intermediate = 1 if is_intermediate else 0
        self._device = device
        self._batch_frames = batch_frames
        self._bulk_frames = bulk_frames          # frames per batch of read_frames_dense / sum_frames
        self._bulk_inflight = max(1, int(bulk_inflight))   # ... and batches of them in flight (one engine + stream each)
        self._bulk = None
        self._ahead = []                         # decoded read-ahead frames of get_next_frame
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
        for name in ('_io', '_pool'):
            if getattr(self, name, None) is not None:
                getattr(self, name).shutdown()
                setattr(self, name, None)
        self._fp.close()

    def seek_to_frame_data(self):
        self._ahead = []
        self._frame_data_start_position = self._rc_header.get_frame_data_offset(self._is_intermediate,
                                                                                self._sz_frame_metadata)
        self._fp.seek(0, 2)
        if self._frame_data_start_position <= self._fp.tell():
            self._fp.seek(self._frame_data_start_position, 0)

    def rewind(self):
        """back to the first frame (sequential reads start over; the bulk engines and their buffers are kept)"""
        self._current_frame_index = 0
        self.seek_to_frame_data()

    def get_file_position(self):
        return self._ahead[0][3] if self._ahead else self._fp.tell()

    def copy_headers_to(self, target_fp, source_header_length):
        self._ahead = []
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
            pk = None
            if level in (1, 2) and h['rc_operation_mode'] == 1:
                name = 'bytes_in_packed_' + ('pixvals' if level == 1 else 'summary_stats')
                pk = [int(mds[i0 + j][name]) for j in range(len(part))]
            sizes = eng.check(pk)
            tri = eng.sparse_coo()
            stats = eng.summary_stats(sizes) if level == 2 else [None] * len(part)
            for j, (rows, cols, vals) in enumerate(tri):
                if level == 1:
                    npk = int(mds[i0 + j]['bytes_in_packed_pixvals'])
                    if (rows.shape[0] * b + 7) // 8 != npk:
                        raise ValueError('frame has %d foreground pixels but %d packed bytes' % (rows.shape[0], npk))
                coo = coo_matrix((vals, (rows, cols)), shape=(h['ny'], h['nx']), dtype=self._numpy_dtype, copy=False)
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
        self._ahead = []
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

    def _drop_ahead(self):
        """forget the decoded read-ahead of get_next_frame and put the file back where the next frame starts"""
        if self._ahead:
            self._fp.seek(self._ahead[0][3], 0)
            self._ahead = []

    def get_next_frame(self):
        # Frames are decoded `batch_frames` at a time (one GPU round trip) and handed out one by one; the frame index
        # and the file position callers can observe stay those of the next frame not yet handed out.
        if not self._ahead:
            first = self._current_frame_index
            batch = []
            while len(batch) < max(1, int(self._batch_frames)):
                if batch and not self._is_intermediate and self._current_frame_index >= self._header['nz']:
                    break                              # end of a merged file: only the caller's own request may raise
                if self._current_frame_index == 0:
                    self._fp.seek(self._frame_data_start_position, 0)
                pos = self._fp.tell()
                r = self._next_raw()
                if r is None:
                    break
                batch.append((r[0], r[1], r[2], pos))
                self._current_frame_index += 1
            self._current_frame_index = first
            if not batch:
                return None
            dec = self._decode([b[2] for b in batch], [b[1] for b in batch])
            self._ahead = [(b[0], b[1], d, b[3]) for b, d in zip(batch, dec)]
        frame_id, md, d, _ = self._ahead.pop(0)
        self._current_frame_index += 1
        return {frame_id: self._frame_dict(md, d)}

    def get_next_frame_raw(self, read_data=True):
        self._drop_ahead()
        r = self._next_raw(read_data=read_data)
        if r is None:
            return None
        frame_id, md, raw = r
        self._current_frame_index += 1
        return {frame_id: {'metadata': md, 'data': raw if read_data else self._fp.tell()}}

    # ---- batched extras (device resident results) ------------------------------------------------
    def _next_batch_raw(self, n):
        self._drop_ahead()
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

    # Bulk decoding keeps several batches in flight, one ReadEngine + CUDA stream each: the records of a batch are
    # read from the file straight into the engine's pinned block (one readinto per frame of a part file, one per
    # batch of a merged file), copied to the device in one piece and inflated / unpacked asynchronously while the
    # host already stages the next batch.  (The serial inflate of a 16 KiB chunk takes milliseconds whatever the
    # batch size, so throughput comes from the number of chunks in flight.)
    def _bulk_engines(self, n_inflight=None):
        if getattr(self, '_bulk', None) is None:
            from .engine import ReadEngine
            import torch
            h = self._header
            if h['target_dtype'] != 0 or not 1 <= h['target_bit_depth'] <= 16:
                raise NotImplementedError('only unsigned targets of 1..16 bits are supported on the GPU path')
            itemsize = 1 if h['target_bit_depth'] <= 8 else 2
            self._bulk = []
            # the lane-per-chunk inflate is a serial chain per lane (milliseconds per batch at a few per cent of the
            # warp slots): throughput comes from the number of batches in flight
            for _ in range(n_inflight or self._bulk_inflight):
                e = ReadEngine(h['ny'], h['nx'], itemsize, h['target_bit_depth'], h['reduction_level'],
                               h['rc_operation_mode'], max_frames=self._bulk_frames, device=self._device)
                e.stream = torch.cuda.Stream(device=e.dev)
                # pinned staging for one batch, sized from the file's average record (it grows on demand)
                e.block_buffer(int(self._file_size / max(1, h['nz']) * self._bulk_frames * 1.25) + (1 << 20))
                self._bulk.append(e)
        return self._bulk

    def _plan_block(self, n):
        """Locate the records of the next (up to) n frames without reading their payloads, and advance the reader
        past them.  -> dict: ids, start (file offset of the byte range), nbytes, map_off / map_sz / val_off / val_sz
        (offsets relative to start; val_* are None for levels 3 / 4), packed (bytes the value streams inflate to)."""
        self._drop_ahead()
        h = self._header
        level = h['reduction_level']
        two = level in (1, 2)
        vname = 'bytes_in_compressed_' + ('pixvals' if level == 1 else 'summary_stats')
        pname = 'bytes_in_packed_' + ('pixvals' if level == 1 else 'summary_stats')
        pk = []
        if self._current_frame_index == 0:
            self._fp.seek(self._frame_data_start_position, 0)
        ids, moff, msz, voff, vsz = [], [], [], [], []
        start, pos = self._fp.tell(), 0
        fd = self._fp.fileno()
        if self._is_intermediate:
            # records are [frame_id][sizes...][map stream][value stream], back to back: walk the small headers with
            # preads; the whole byte range of the batch is fetched at once afterwards (in parallel slices)
            nf = len(self._sm)
            names = [f['name'] for f in self._sm]
            i_map = names.index('bytes_in_compressed_binary_map')
            i_val = names.index(vname) if two else -1
            i_pk = names.index(pname) if two else -1
            hlen = 4 + 4 * nf
            fpos = start
            while len(ids) < n:
                hdr = os.pread(fd, hlen, fpos)
                if len(hdr) < hlen:
                    break
                rec = np.frombuffer(hdr, dtype='<u4')
                n_map = int(rec[1 + i_map])
                n_val = int(rec[1 + i_val]) if two else 0
                if fpos + hlen + n_map + n_val > self._file_size:
                    raise ValueError('truncated record of frame %d' % int(rec[0]))
                ids.append(int(rec[0]))
                moff.append(fpos - start + hlen); msz.append(n_map)
                if two:
                    voff.append(fpos - start + hlen + n_map); vsz.append(n_val)
                    pk.append(int(rec[1 + i_pk]))
                fpos += hlen + n_map + n_val
            pos = fpos - start
            if ids:
                self._fp.seek(fpos, 0)
                self._current_frame_index += len(ids)
        else:
            z0 = self._current_frame_index
            z1 = min(h['nz'], z0 + n)
            if z1 > z0:
                sizes = self._seek_table[z0:z1, 0].astype(np.int64)
                rel0 = int(self._seek_table[z0, 1])
                start = self._frame_data_start_position + rel0
                pos = int(sizes.sum())
                self._fp.seek(start + pos, 0)                                  # sequential reads continue here
                rel = (self._seek_table[z0:z1, 1].astype(np.int64) - rel0)
                for k, z in enumerate(range(z0, z1)):
                    md = self._frame_metadata[z]
                    n_map = int(md['bytes_in_compressed_binary_map'])
                    ids.append(z)
                    moff.append(int(rel[k])); msz.append(n_map)
                    if two:
                        voff.append(int(rel[k]) + n_map); vsz.append(int(md[vname]))
                        pk.append(int(md[pname]))
                self._current_frame_index = z1
        if not two:
            voff = vsz = None
        return {'ids': ids, 'start': start, 'nbytes': pos, 'map_off': moff, 'map_sz': msz, 'val_off': voff,
                'val_sz': vsz, 'packed': pk or None}

    def _read_block(self, n, eng):
        """Reads the records of up to n frames into eng's pinned block.
        -> (frame ids, nbytes, map_off, map_sz, val_off, val_sz); val_* are None for levels 3 / 4."""
        p = self._plan_block(n)
        self._block_packed = p['packed']
        if p['ids']:
            self._pread_parallel(self._fp.fileno(), eng.block_buffer(p['nbytes'] + 16), p['start'], p['nbytes'])
        return p['ids'], p['nbytes'], p['map_off'], p['map_sz'], p['val_off'], p['val_sz']

    def _pread_parallel(self, fd, buf, offset, nbytes, n_threads=None):
        """file[offset : offset + nbytes] -> buf[:nbytes] (pinned), in slices read by a few threads (preadv releases
        the GIL; one thread copies out of the page cache at only a few GB/s)"""
        mv = memoryview(buf)

        def rd(a, b):
            while a < b:
                k = os.preadv(fd, [mv[a:b]], offset + a)
                if k <= 0:
                    raise ValueError('truncated file')
                a += k

        if nbytes < (8 << 20):
            rd(0, nbytes)
            return
        if n_threads is None:
            # copying out of the page cache is memory-bound: 8 threads reach the host's copy bandwidth (39.5 k frames/s;
            # 16 / 24 / 32 threads: 36.1 / 36.0 / 35.5 k, and the enqueueing main thread waits for a core twice as long)
            try:
                cap = int(os.environ.get('RECODE_B200_READ_THREADS', '8'))
                n_threads = max(4, min(cap, len(os.sched_getaffinity(0))))
            except AttributeError:
                n_threads = 8
        if getattr(self, '_pool', None) is None:
            from concurrent.futures import ThreadPoolExecutor
            self._pool = ThreadPoolExecutor(n_threads)
        step = (nbytes + n_threads - 1) // n_threads
        step = (step + 4095) // 4096 * 4096
        futs = [self._pool.submit(rd, a, min(nbytes, a + step)) for a in range(0, nbytes, step)]
        for f in futs:
            f.result()

    def _bulk_run(self, n, consume):
        """Pipelined decode of the next n frames; consume(engine) enqueues the unpack of a loaded batch on the
        engine's stream.  The file read of batch k + 1 (an I/O thread, preadv releases the GIL) runs while batch k is
        enqueued.  -> frame ids (in file order)"""
        import torch
        if self._header['rc_operation_mode'] != 1:
            raise ValueError('bulk decoding handles compressed files (rc_operation_mode 1) only')
        import time
        engs = self._bulk_engines()
        if getattr(self, '_io', None) is None:
            from concurrent.futures import ThreadPoolExecutor
            self._io = ThreadPoolExecutor(1)
        pending = []                                   # engines whose batch has not been checked yet
        ids = []
        st = self.bulk_stats = {'file_read_s': 0.0, 'enqueue_s': 0.0, 'wait_s': 0.0, 'bytes': 0}
        state = {'k': 0, 'planned': 0}
        fd = self._fp.fileno()

        def timed_read(buf, start, nbytes):
            t = time.perf_counter()
            self._pread_parallel(fd, buf, start, nbytes)
            return time.perf_counter() - t

        def start_read():
            """plan the next batch and start reading it into the next engine's pinned block"""
            if state['planned'] >= n:
                return None
            eng = engs[state['k'] % len(engs)]
            t0 = time.perf_counter()
            if eng in pending:                         # its previous batch: validate before reusing the buffers
                eng.stream.synchronize()
                eng.check(eng.expect_packed)
                pending.remove(eng)
            eng.wait_block_free()
            st['wait_s'] += time.perf_counter() - t0
            plan = self._plan_block(min(eng.max_frames, n - state['planned']))
            if not plan['ids']:
                return None
            fut = self._io.submit(timed_read, eng.block_buffer(plan['nbytes'] + 16), plan['start'], plan['nbytes'])
            state['planned'] += len(plan['ids'])
            state['k'] += 1
            return eng, plan, fut

        cur = start_read()
        while cur is not None:
            nxt = start_read()
            eng, plan, fut = cur
            t0 = time.perf_counter()
            st['file_read_s'] += fut.result()          # time spent in the I/O thread (overlaps the enqueues)
            t1 = time.perf_counter()
            st['wait_s'] += t1 - t0
            eng.expect_packed = plan['packed']
            eng.stream.wait_stream(torch.cuda.current_stream(eng.dev))
            with torch.cuda.stream(eng.stream):
                eng.load_block(plan['nbytes'], plan['map_off'], plan['map_sz'], plan['val_off'], plan['val_sz'])
                consume(eng)
            st['enqueue_s'] += time.perf_counter() - t1
            st['bytes'] += plan['nbytes']
            pending.append(eng)
            ids += plan['ids']
            cur = nxt
        t0 = time.perf_counter()
        for eng in pending:
            eng.stream.synchronize()
            eng.check(eng.expect_packed)
            torch.cuda.current_stream(eng.dev).wait_stream(eng.stream)
        st['wait_s'] += time.perf_counter() - t0
        return ids

    def decode_stage_ms(self, reps=3):
        """Profiling aid (bench.py): one batch of the file decoded ALONE on one engine, timed with CUDA events.
        -> dict: frames, map_inflate_ms (all inflate kernels of the map streams), lanes_ms (k_inflate_lanes of the map
        streams alone), value_inflate_ms, unpack_dense_ms, unpack_sum_ms.  Leaves the reader rewound."""
        import torch
        engs = self._bulk_engines()
        eng = engs[0]
        self.rewind()
        with torch.cuda.device(eng.dev):
            torch.cuda.synchronize()
            bi, nbytes, moff, msz, voff, vsz = self._read_block(eng.max_frames, eng)
            eng.expect_packed = self._block_packed
            n = len(bi)
            out = {'frames': n}
            if n == 0:
                return out
            dense = torch.empty((n, self._header['ny'], self._header['nx']), dtype=eng.t_dtype, device=eng.dev)
            total = torch.zeros(self._header['ny'] * self._header['nx'], dtype=torch.int32, device=eng.dev)
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
            eng.ctx.profile_enable(2)
            acc = {'map_inflate_ms': 0.0, 'lanes_ms': 0.0, 'value_inflate_ms': 0.0, 'unpack_dense_ms': 0.0,
                   'unpack_sum_ms': 0.0}
            for _ in range(reps + 1):
                ev[0].record()
                eng.load_block(nbytes, moff, msz, voff, vsz, maps_only=True)
                ev[1].record()
                torch.cuda.synchronize()
                d = eng.ctx.profile_read_detail()
                eng.load_block(nbytes, moff, msz, voff, vsz)
                ev[2].record()
                eng.dense(out=dense)
                ev[3].record()
                eng.dense(total=total, want_dense=False)
                ev[4].record()
                torch.cuda.synchronize()
                if _ == 0:
                    continue                           # warm-up
                acc['map_inflate_ms'] += ev[0].elapsed_time(ev[1])
                acc['lanes_ms'] += d[1] if len(d) > 1 else 0.0
                acc['value_inflate_ms'] += max(ev[1].elapsed_time(ev[2]) - ev[0].elapsed_time(ev[1]), 0.0)
                acc['unpack_dense_ms'] += ev[2].elapsed_time(ev[3])
                acc['unpack_sum_ms'] += ev[3].elapsed_time(ev[4])
            eng.ctx.profile_enable(0)
            eng.check(eng.expect_packed)
            out.update({k: v / reps for k, v in acc.items()})
        self.rewind()
        return out

    def read_frames_dense(self, n):
        """next n frames -> (frame ids, CUDA tensor [k, ny, nx] of the target dtype); k <= n at EOF"""
        import torch
        if self._header['rc_operation_mode'] != 1:
            return self._read_frames_dense_serial(n)
        # one output tensor for the whole request; every batch is unpacked straight into its slice
        h = self._header
        engs = self._bulk_engines()
        left = min(n, max(h['nz'] - self._current_frame_index, 0)) if h['nz'] > 0 else n
        out = torch.empty((max(left, 1), h['ny'], h['nx']), dtype=engs[0].t_dtype, device=engs[0].dev)
        done = [0]

        def consume(eng):
            k = eng.n
            if done[0] + k > out.shape[0]:
                raise RuntimeError('more frames in the file than its header announces')
            eng.dense(out=out[done[0]:done[0] + k])
            done[0] += k

        ids = self._bulk_run(min(n, out.shape[0]), consume)
        if not ids:
            return ids, None
        return ids, out[:len(ids)]

    def sum_frames(self, n, total=None):
        """live view: adds the next n frames into `total` (uint32 CUDA tensor [ny*nx], created if None) without
        materialising dense frames (examples/ReCoDe_Live_View_MT.ipynb cell 1) -> (frame ids, total)"""
        import torch
        if self._header['rc_operation_mode'] != 1:
            return self._sum_frames_serial(n, total)
        engs = self._bulk_engines()
        if total is None:
            total = torch.zeros(self._header['ny'] * self._header['nx'], dtype=torch.int32, device=engs[0].dev)
        # the unpack kernel accumulates with atomics, so batches on different streams may share `total`
        ids = self._bulk_run(n, lambda eng: eng.dense(total=total, want_dense=False))
        return ids, total

    def _read_frames_dense_serial(self, n):
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

    def _sum_frames_serial(self, n, total=None):
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
