"""ReCoDeWriter -- drop-in for pyrecode/recode_writer.py with the per-frame hot path on the GPU.

Same constructor, start() / run(data) / close() protocol, part-file naming, header, record layout and
run_metrics keys as the reference (pyrecode/recode_writer.py:24-619).  The difference is inside run(): instead of
calling _reduce_compress once per frame (recode_writer.py:383-387), frames are processed in batches by
rc_reduce_compress (include/recode_b200.h), which returns the finished part-file records.

Scope: unsigned sources of 1..16 bits with source_bit_depth == target_bit_depth, zlib
(compression_scheme 0), compression_level 0..9, in-memory data or raw binary files.  Anything else raises
NotImplementedError -- there is no CPU fallback.
"""
import math
import os
import sys
import warnings
from datetime import datetime, timedelta
from pathlib import Path

import numpy as np

from .misc import rc_cfg as rc
from .params import InitParams, InputParams
from .recode_header import ReCoDeHeader
from .structures import ReCoDeStructures

_WORKSPACE_BUDGET = 12 << 30       # bytes of device workspace a writer may claim
_BATCH_INPUT_TARGET = 512 << 20    # raw frame bytes per launch batch


class ReCoDeWriter:

    def __init__(self, image_filename, dark_data=None, dark_filename='', output_directory='', input_params=None,
                 params_filename='', mode='batch', validation_frame_gap=-1, log_filename='recode.log', run_name='run',
                 verbosity=0, use_c=False, max_count=-1, chunk_time_in_sec=0, node_id=0, buffer_size_in_frames=10.0,
                 device=None, batch_frames=None, merged=False):
        self._init_params = InitParams(mode, output_directory, image_filename=image_filename,
                                       calibration_filename=dark_filename, params_filename=params_filename,
                                       validation_frame_gap=validation_frame_gap, log_filename=log_filename,
                                       run_name=run_name, verbosity=verbosity, use_c=use_c)
        if input_params is None:
            self._input_params = InputParams()
            self._input_params.load(Path(self._init_params.params_filename))
        else:
            self._input_params = input_params
        if not self._input_params.validate():
            raise ValueError('Invalid input params')
        ip = self._input_params
        if self._init_params.use_c:
            if ip.source_numpy_dtype != np.uint16 or ip.target_numpy_dtype != np.uint16:
                raise ValueError('use_c=True can only be used if source and target dtypes are both unsigned 16-bit')
        self._check_scope()

        self._rc_header = ReCoDeHeader()
        self._rc_header.create(self._init_params, ip, True)
        self._rc_header.set('source_header_length',
                            1024 if ip.source_file_type in (rc.FILE_TYPE_MRC, rc.FILE_TYPE_SEQ) else 0)
        if self._init_params.verbosity > 0:
            self._rc_header.print()
        if not self._rc_header.validate():
            raise ValueError('Invalid ReCoDe header created')
        self._header = self._rc_header.as_dict()

        self._src_dtype = ip.source_numpy_dtype
        if dark_data is None:
            if ip.calibration_file_type == rc.FILE_TYPE_BINARY:
                t = np.fromfile(self._init_params.calibration_filename, dtype=self._src_dtype,
                                count=self._header['ny'] * self._header['nx']).reshape(self._header['ny'],
                                                                                     self._header['nx'])
            elif ip.calibration_file_type in (rc.FILE_TYPE_MRC, rc.FILE_TYPE_SEQ):
                raise NotImplementedError('MRC / SEQ calibration files need mrcfile / pims; pass dark_data instead')
            else:
                raise NotImplementedError("No implementation available for loading calibration file of type 'Other'")
        else:
            t = np.squeeze(dark_data[0]) if np.ndim(dark_data) > 2 else np.asarray(dark_data)
        if self._header['ny'] != t.shape[0] or self._header['nx'] != t.shape[1]:
            raise RuntimeError('Data and Calibration frames have different shapes')
        eps = ip.calibration_threshold_epsilon
        if not (0 <= eps <= np.iinfo(t.dtype if t.dtype.kind == 'u' else self._src_dtype).max):
            raise OverflowError('calibration_threshold_epsilon %d out of bounds for %s' % (eps, t.dtype))
        # dark + eps is evaluated in the calibration frame's dtype, then cast to the source dtype
        # (recode_writer.py:126-137); for an unsigned dark frame that is a wrapping add in that dtype
        self._calibration_frame = t
        if t.dtype != self._src_dtype:
            warnings.warn('Calibration data type not same as source. Attempting to cast.')
            thr = (t + t.dtype.type(eps)).astype(self._src_dtype)
            self._calibration_frame = t.astype(self._src_dtype)
            self._thr_host = thr
        else:
            self._thr_host = None

        # merged=True (an extension): a single batch-mode writer emits the random-access layout of
        # merge_parts directly -- header, nz x metadata table, payloads -- instead of a part file to be merged later
        self._merged = bool(merged)
        if self._merged and (mode != 'batch' or ip.num_threads != 1 or node_id != 0):
            raise ValueError("merged=True needs mode='batch', num_threads = 1 and node_id = 0")
        self._merged_table = []
        self._merged_table_pos = None
        self._node_id = node_id
        self._device = device
        self._batch_frames = batch_frames
        self._buffer_size_in_frames = buffer_size_in_frames
        self._structures = ReCoDeStructures(self._header)
        self._engine = None
        self._intermediate_file_name = None
        self._intermediate_file = None
        self._validation_file_name = None
        self._validation_file = None
        self._frame_sz = None
        self._chunk_offset = None
        self._num_frames_in_part = None
        self._is_first_chunk = True
        self._vc_roi = {'x_start': None, 'y_start': None, 'nx': None, 'ny': None}
        self._vc_n_pixels = None
        self._vc_dose_rate = 0.0
        self._vc_engine = None
        self._val_engine = None
        self._source_shape = None

    # ------------------------------------------------------------------------------------------
    def _check_scope(self):
        ip = self._input_params
        if ip.source_data_type != 0 or ip.target_data_type != 0:
            raise NotImplementedError('only unsigned integer sources / targets are supported on the GPU path')
        if not 1 <= ip.source_bit_depth <= 16:
            raise NotImplementedError('source_bit_depth must be in 1..16 on the GPU path')
        if ip.target_bit_depth != ip.source_bit_depth:
            raise NotImplementedError('target_bit_depth must equal source_bit_depth: the format packs with the '
                                      'source depth and unpacks with the target depth')
        if ip.compression_scheme != 0:
            raise NotImplementedError('only compression_scheme 0 (zlib / deflate) is supported on the GPU path')
        if ip.rc_operation_mode == 1 and ip.compression_level > 9:
            raise ValueError('Bad compression level')          # what zlib.compress raises in the reference
        if ip.nx > 65535 or ip.ny > 65535:
            raise NotImplementedError('nx, ny must be <= 65535')

    def _make_engine(self):
        from .engine import WriteEngine          # imports torch + the CUDA library; fails loudly without a GPU
        ip = self._input_params
        ny, nx = self._header['ny'], self._header['nx']
        itemsize = np.dtype(self._src_dtype).itemsize
        frame_bytes = ny * nx * itemsize
        F = self._batch_frames or max(1, min(64, _BATCH_INPUT_TARGET // frame_bytes))
        import ctypes
        from . import _native
        while True:
            cfg = _native.make_config(ny, nx, itemsize, ip.source_bit_depth, ip.reduction_level,
                                      ip.rc_operation_mode, ip.L2_statistics, ip.L4_centroiding,
                                      min(ip.compression_level, 9), F)
            if _native.lib().rc_workspace_bytes(ctypes.byref(cfg)) <= _WORKSPACE_BUDGET or F == 1:
                break
            F = max(1, F // 2)
        # a record larger than the raw frame is an error in the reference (recode_writer.py:565-566); the batch
        # buffer therefore never needs more than F raw frames (+ headers)
        cap = F * (frame_bytes + 64) + 4096
        eng = WriteEngine(ny, nx, itemsize, ip.source_bit_depth, ip.reduction_level, ip.rc_operation_mode,
                          ip.L2_statistics, ip.L4_centroiding, min(ip.compression_level, 9), max_frames=F,
                          device=self._device, records_capacity=cap, n_slots=2)
        if self._thr_host is not None:
            import torch
            eng.thr = torch.from_numpy(np.ascontiguousarray(self._thr_host)).to(eng.dev)
        else:
            eng.set_threshold(self._calibration_frame, ip.calibration_threshold_epsilon)
        for slot in eng.slots:
            slot.ctx.profile_enable(True)
        return eng

    # ------------------------------------------------------------------------------------------
    def start(self):
        """Create the part file and the device buffers (recode_writer.py:184-240)."""
        if self._init_params.mode == 'batch':
            base_filename = Path(self._init_params.image_filename).stem
        else:
            base_filename = self._init_params.run_name
        self._intermediate_file_name = os.path.join(
            self._init_params.output_directory,
            base_filename + '.rc' + str(self._input_params.reduction_level) +
            ('' if self._merged else '_part' + '{0:03d}'.format(self._node_id)))
        self._intermediate_file = open(self._intermediate_file_name, 'wb')
        if self._rc_header.as_dict()['nz'] < 0:
            # num_frames = -1 ("take the count from the source"): the reference cannot serialize that placeholder
            # (recode_header.py:274 raises OverflowError); the true count is written by close() either way
            self._rc_header.set('nz', 0)
        self._rc_header.serialize_to(self._intermediate_file)
        self._intermediate_file.flush()
        if self._init_params.validation_frame_gap > 0:
            self._validation_file_name = os.path.join(
                self._init_params.output_directory,
                base_filename + '_part' + '{0:03d}'.format(self._node_id) + '_validation_frames.bin')
            self._validation_file = open(self._validation_file_name, 'wb')
        self._frame_sz = self._header['ny'] * self._header['nx'] * np.dtype(self._src_dtype).itemsize
        self._engine = self._make_engine()
        self._chunk_offset = 0
        self._num_frames_in_part = 0
        self._vc_roi['nx'] = min(self._header['nx'], 128)
        self._vc_roi['ny'] = min(self._header['ny'], 128)
        self._vc_roi['x_start'] = math.floor((self._header['nx'] - self._vc_roi['nx']) / 2.0)
        self._vc_roi['y_start'] = math.floor((self._header['ny'] - self._vc_roi['ny']) / 2.0)
        self._vc_n_pixels = self._vc_roi['nx'] * self._vc_roi['ny']

    def _do_sanity_checks(self, data=None):
        if data is None:
            ft = self._input_params.source_file_type
            if ft == rc.FILE_TYPE_BINARY:
                self._source_shape = (self._header['nz'], self._header['ny'], self._header['nx'])
            elif ft == rc.FILE_TYPE_SEQ:
                # frames present in the chunk right now (the header's count may be ahead of the file while the
                # acquisition is writing: recode_writer.py:330-347 falls back to frame-by-frame reads for that)
                from .em_reader import SEQReader
                with SEQReader(self._init_params.image_filename) as f:
                    self._source_shape = tuple(f.shape)
                    if np.dtype(f.dtype).itemsize != np.dtype(self._src_dtype).itemsize:
                        raise RuntimeError('Sequence file holds %s pixels, the params say %s'
                                           % (np.dtype(f.dtype).name, np.dtype(self._src_dtype).name))
            elif ft == rc.FILE_TYPE_MRC:
                raise NotImplementedError('MRC sources need mrcfile; pass data= instead')
            else:
                raise NotImplementedError("No implementation available for loading calibration file of type 'Other'")
        else:
            self._source_shape = tuple(data.shape)
        if self._source_shape[1] != self._header['ny']:
            raise RuntimeError('Expected height does not match height in source file')
        if self._source_shape[2] != self._header['nx']:
            raise RuntimeError('Expected width does not match width in source file')
        if self._input_params.num_frames == -1:
            self._header['nz'] = self._source_shape[0]
        elif self._input_params.num_frames > self._source_shape[0]:
            raise RuntimeError('Number of frames requested in config file is larger than available in source file')
        else:
            self._header['nz'] = self._input_params.num_frames

    def run(self, data=None):
        """Process this node's share of a chunk of frames (recode_writer.py:292-428).  `data` is
        [nz, ny, nx]: a numpy array, or a CUDA torch tensor of the source dtype (no host round trip); None reads the
        source file (raw binary, or the SEQ chunk of a streaming session).

        run_metrics keeps the reference's keys, but the GPU path has other stage boundaries (rc_profile_read), summed
        over the run's batches as GPU time:
          frame_thresholding_and_counting_time   threshold compare + binary-map packing + value compaction (one kernel)
          frame_binary_image_packing_time        0 (fused into the kernel above)
          frame_pixel_intensity_packing_time     the rest of the reduction: puddle labelling, statistics / centroids,
                                                 bit packing
          frame_binary_image_compression_time    deflate of BOTH streams (maps and values run as two groups of one stage)
          frame_pixel_intensity_compression_time 0 (included above)
          frame_time                             their sum (+ record assembly)
        With two batches in flight the stages of different batches overlap, so the sum can exceed run_time."""
        import torch
        run_metrics = {}
        self._do_sanity_checks(data)
        if self._is_first_chunk and data is None and self._input_params.source_file_type == rc.FILE_TYPE_SEQ:
            # the source header follows the ReCoDe header (recode_writer.py:267-270); for SEQ sources the reference
            # stores 1024 zero bytes (em_reader.py:296-300)
            self._intermediate_file.write(bytes(1024))
            self._intermediate_file.flush()
        self._is_first_chunk = False

        if self._init_params.mode == 'batch':
            n_frames_in_chunk = self._input_params.nz
        elif self._init_params.mode == 'stream':
            n_frames_in_chunk = self._source_shape[0]
        else:
            raise ValueError("Invalid input params: mode. Can be 'batch' or 'stream'.")
        # the reference's partition rule (recode_writer.py:320-322)
        n_frames_per_thread = int(math.ceil((n_frames_in_chunk * 1.0) / (self._input_params.num_threads * 1.0)))
        frame_offset = self._node_id * n_frames_per_thread
        available_frames = min(n_frames_per_thread, max(n_frames_in_chunk - frame_offset, 0))

        stt = datetime.now()
        seq = None
        if data is None and self._input_params.source_file_type == rc.FILE_TYPE_SEQ:
            # read batch by batch straight into the engine's pinned staging buffers (below)
            from .em_reader import SEQReader
            seq = SEQReader(self._init_params.image_filename)
            available_frames = min(available_frames, max(seq.shape[0] - frame_offset, 0))
        elif data is None:
            itemsize = np.dtype(self._src_dtype).itemsize
            off = self._input_params.source_header_length + \
                (self._input_params.frame_offset + frame_offset) * self._frame_sz
            data = np.fromfile(self._init_params.image_filename, dtype=self._src_dtype,
                               count=available_frames * self._header['ny'] * self._header['nx'], offset=off)
            available_frames = data.size // (self._header['ny'] * self._header['nx'])
            data = data[:available_frames * self._header['ny'] * self._header['nx']].reshape(
                available_frames, self._header['ny'], self._header['nx'])
            del itemsize
        else:
            data = data[frame_offset:frame_offset + available_frames]
        on_device = isinstance(data, torch.Tensor) and data.is_cuda
        if seq is None and not on_device and isinstance(data, np.ndarray) and data.dtype != self._src_dtype:
            warnings.warn('Source data type either not as specified or does not match params specs. Attempting to cast.')
            data = data.astype(self._src_dtype)
        run_metrics['run_data_read_time'] = datetime.now() - stt

        keys = ('frame_thresholding_and_counting_time', 'frame_binary_image_packing_time',
                'frame_pixel_intensity_packing_time', 'frame_binary_image_compression_time',
                'frame_pixel_intensity_compression_time', 'frame_time')
        gpu_ms = dict.fromkeys(keys, 0.0)
        run_start = datetime.now()
        eng = self._engine
        F = eng.max_frames
        gap = self._init_params.validation_frame_gap
        # two batches in flight: the copies and the file write of one batch overlap the kernels of the next
        def finish(ticket):
            slot, first_id, batch, n = ticket
            rec, offs, counts, _, _ = eng.collect(slot)
            sizes = np.diff(offs)
            if sizes.size and int(sizes.max()) > self._frame_sz:
                raise ValueError('Buffer size smaller than compressed data size')
            if self._merged:
                # payloads only; the [sizes...] of every record go to the metadata table written by close()
                hl = 4 + 4 * n_meta
                table = np.empty((n, n_meta), dtype='<u4')
                mv = memoryview(rec)
                parts = []
                for i in range(n):
                    o = int(offs[i])
                    table[i] = np.frombuffer(mv[o + 4:o + hl], dtype='<u4')
                    parts.append(mv[o + hl:int(offs[i + 1])])
                self._merged_table.append(table)
                self._intermediate_file.writelines(parts)
            else:
                self._intermediate_file.write(rec)
            st = eng.slots[slot].ctx.profile_read()
            if len(st) >= 4:
                gpu_ms['frame_thresholding_and_counting_time'] += st[0]
                gpu_ms['frame_pixel_intensity_packing_time'] += st[1]
                gpu_ms['frame_binary_image_compression_time'] += st[2]
                gpu_ms['frame_time'] += sum(st[:4])
            if gap > 0:
                for i in range(n):
                    if (first_id + i) % gap == 0:
                        self._validation_frame(batch[i], run_metrics)

        n_meta = len(self._structures.standard_frame_metadata_structure_for(self._header['reduction_level'],
                                                                            self._header['rc_operation_mode']))
        if self._merged:
            if self._merged_table_pos is not None:
                raise RuntimeError('merged=True writes the whole dataset in one run()')
            # the table comes first in the file: reserve it now that the frame count is known
            self._merged_table_pos = self._intermediate_file.tell()
            self._intermediate_file.write(bytes(4 * n_meta * available_frames))
        pending = None
        for b0 in range(0, available_frames, F):
            n = min(F, available_frames - b0)
            first_id = self._chunk_offset + frame_offset + b0
            if seq is not None:
                batch = eng.pinned_input(n)                   # the next slot's staging buffer (its last batch is done)
                seq.read_into(batch.numpy().view(self._src_dtype), frame_offset + b0, frame_offset + b0 + n)
            else:
                batch = data[b0:b0 + n]
            ticket = (eng.submit(batch, first_frame_id=first_id), first_id, batch, n)
            if pending is not None:
                finish(pending)
            pending = ticket
        if pending is not None:
            finish(pending)
        if seq is not None:
            seq.close()
        self._intermediate_file.flush()
        for k in keys:
            run_metrics[k] = timedelta(milliseconds=gpu_ms[k])
        if run_metrics['frame_time'] == timedelta(0):
            run_metrics['frame_time'] = timedelta(microseconds=1)

        self._chunk_offset += n_frames_in_chunk
        self._num_frames_in_part += available_frames
        run_metrics['run_time'] = datetime.now() - run_start
        run_metrics['run_frames'] = available_frames
        return run_metrics

    def _validation_frame(self, frame, run_metrics):
        """dump the raw frame and estimate the dose rate: puddles in the central <=128 x 128 ROI of the frame's
        binary map divided by the ROI area (recode_writer.py:402-415).  Rare path; still on the GPU."""
        import torch
        from .engine import WriteEngine
        fh = frame.cpu().numpy() if isinstance(frame, torch.Tensor) else np.asarray(frame)
        self._validation_file.write(fh.tobytes())
        # a private one-frame engine (own context, workspace, pinned staging and device buffers): the batch engine's
        # slots hold the batches that are still in flight while a finished batch is validated
        if self._val_engine is None:
            be = self._engine
            self._val_engine = WriteEngine(be.ny, be.nx, be.itemsize, be.bit_depth, 3, max_frames=1, device=self._device)
            self._val_engine.thr = be.thr
        maps, _, _ = self._val_engine.reduce(fh[None])
        ny, nx = self._header['ny'], self._header['nx']
        bits = np.unpackbits(np.frombuffer(maps[0], dtype=np.uint8), bitorder='little')[:ny * nx].reshape(ny, nx)
        roi = self._vc_roi
        crop = bits[roi['y_start']:roi['y_start'] + roi['ny'], roi['x_start']:roi['x_start'] + roi['nx']]
        if self._vc_engine is None:
            self._vc_engine = WriteEngine(roi['ny'], roi['nx'], 2, 16, 2, max_frames=1, device=self._device)
        _, k = self._vc_engine.labels([np.packbits(crop.ravel(), bitorder='little').tobytes()])
        self._vc_dose_rate = int(k[0]) / self._vc_n_pixels
        run_metrics.setdefault('run_dose_rates', []).append(self._vc_dose_rate)

    def close(self):
        """rewrite the header with the true frame count and close the part file (recode_writer.py:589-603)"""
        self._rc_header.update('nz', self._num_frames_in_part)
        if self._merged and self._merged_table:
            self._intermediate_file.seek(self._merged_table_pos)
            self._intermediate_file.write(np.concatenate(self._merged_table).tobytes())
        self._intermediate_file.seek(0)
        self._rc_header.serialize_to(self._intermediate_file)
        self._intermediate_file.close()
        if self._init_params.validation_frame_gap > 0:
            self._validation_file.close()


def print_run_metrics(run_metrics):
    for key in run_metrics:
        if key.startswith('frame_'):
            print(key, "\t", run_metrics[key] / run_metrics['run_frames'], "\t",
                  run_metrics[key] / run_metrics['frame_time'])
        elif key == 'run_dose_rates':
            print(key, "\t", run_metrics[key], "\t", 'Avg.=', np.mean(run_metrics[key]))
        else:
            print(key, "\t", run_metrics[key])


if __name__ == "__main__":
    import argparse
    parser = argparse.ArgumentParser(description='ReCoDe writer (GPU)')
    parser.add_argument('--image_filename', default='')
    parser.add_argument('--calibration_file', default='')
    parser.add_argument('--out_dir', default='')
    parser.add_argument('--params_file', default='')
    parser.add_argument('--node_id', type=int, default=0)
    args = parser.parse_args()
    writer = ReCoDeWriter(args.image_filename, dark_filename=args.calibration_file, output_directory=args.out_dir,
                          params_filename=args.params_file, mode='batch', node_id=args.node_id)
    writer.start()
    metrics = writer.run()
    writer.close()
    print_run_metrics(metrics)
    sys.exit(0)
