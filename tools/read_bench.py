"""Read-path timing (not a test): decode L1 / L2 records of 4096 x 4096 frames to dense frames and to the summed
live-view image, own (chunk-parallel) streams and stock-zlib (serial per stream) streams.

    python tests/read_bench.py
"""
import os
import sys
import time
import zlib

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from oracle import oracle as orc
from pyrecode_b200.engine import ReadEngine, WriteEngine

NY = NX = 4096
F = int(os.environ.get('RB_FRAMES', '16'))
for level, kind in ((1, 'l1'), (2, 'l2')):
    dark = orc.synth_dark(NY, NX)
    frames = orc.synth_frames(kind, 2, NY, NX, dark, seed=1234)
    we = WriteEngine(NY, NX, 2, 12, level, 1, 0, 0, 1, max_frames=F, records_capacity=F * (NY * NX // 2))
    we.set_threshold(dark, 20)
    batch = np.stack([frames[i % 2] for i in range(F)])
    rec, offs, counts, _, _ = we.reduce_compress(batch)
    rec = bytes(rec)
    maps, vals = [], []
    for f in range(F):
        r = rec[int(offs[f]):int(offs[f + 1])]
        h = np.frombuffer(r[:16], '<u4')
        maps.append(r[16:16 + h[1]])
        vals.append(r[16 + h[1]:16 + h[1] + h[2]])
    del we
    re_ = ReadEngine(NY, NX, 2, 12, level, 1, max_frames=F)
    for it in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        re_.load(maps, vals); t1 = time.perf_counter(); torch.cuda.synchronize(); t1b = time.perf_counter()
        re_.check(); t2 = time.perf_counter()
        d = re_.dense(); torch.cuda.synchronize(); t3 = time.perf_counter()
        tot = torch.zeros(NY * NX, dtype=torch.int32, device='cuda')
        re_.dense(total=tot, want_dense=False); torch.cuda.synchronize(); t4 = time.perf_counter()
    print('L%d read: host staging + launch %.2f ms, inflate wait %.2f ms, check %.2f, dense %.2f ms, sum %.2f ms for %d frames'
          ' -> %.0f frames/s (dense path)' % (level, (t1 - t0) * 1e3, (t1b - t1) * 1e3, (t2 - t1b) * 1e3, (t3 - t2) * 1e3,
                                              (t4 - t3) * 1e3, F, F / (t3 - t0)))
    zm = [zlib.compress(zlib.decompress(m), 1) for m in maps]
    zv = [zlib.compress(zlib.decompress(v), 1) for v in vals]
    torch.cuda.synchronize(); t0 = time.perf_counter(); re_.load(zm, zv); torch.cuda.synchronize(); t1 = time.perf_counter()
    print('   foreign (stock zlib) streams: load + inflate %.2f ms' % ((t1 - t0) * 1e3))
    del re_

# ---- through the reader API: an L2 part file on tmpfs -> dense frames / live-view sum (bulk path)
import tempfile
from pyrecode_b200.recode_reader import ReCoDeReader
from pyrecode_b200.recode_writer import ReCoDeWriter
from pyrecode_b200.params import InputParams

NZ = int(os.environ.get('RB_FILE_FRAMES', '256'))
tmp = tempfile.mkdtemp(dir='/dev/shm' if os.path.isdir('/dev/shm') else None)
ip = InputParams()
for k, v in dict(l4_centroiding=0, source_file_type=0, num_frames=NZ, source_header_length=0, calibration_frame_offset=0,
                 compression_scheme=0, calibration_file_type=0, compression_level=1, l2_statistics=0,
                 calibration_threshold_epsilon=20, frame_offset=0, num_threads=1, rc_operation_mode=1,
                 num_calibration_frames=1, reduction_level=2, keep_calibration_data=1, source_bit_depth=12,
                 target_bit_depth=12, keep_part_files=0, num_rows=NY, num_cols=NX, source_data_type=0,
                 target_data_type=0).items():
    ip._param_map[k] = v
dark = orc.synth_dark(NY, NX)
fr = orc.synth_frames('l2', 4, NY, NX, dark, seed=1234)
w = ReCoDeWriter('rb', dark_data=dark[None], output_directory=tmp, input_params=ip, mode='batch', node_id=0)
w.start()
w.run(np.stack([fr[i % 4] for i in range(NZ)]))
w.close()
path = os.path.join(tmp, 'rb.rc2_part000')
print('L2 part file: %d frames, %.1f MB' % (NZ, os.path.getsize(path) / 1e6))
for bulk in (32, 64, 128):
    for what in ('sum', 'dense'):
        best = 0
        for it in range(3):
            r = ReCoDeReader(path, is_intermediate=True, bulk_frames=bulk)
            r.open(print_header=False)
            r._bulk_engines()
            torch.cuda.synchronize(); t0 = time.perf_counter()
            if what == 'sum':
                ids, total = r.sum_frames(NZ)
            else:
                ids, dense = r.read_frames_dense(min(NZ, 128))
            torch.cuda.synchronize(); t1 = time.perf_counter()
            if len(ids) / (t1 - t0) > best:
                best, stats = len(ids) / (t1 - t0), dict(r.bulk_stats)
            r.close()
            del r
            if what == 'dense':
                del dense
        print('reader API, bulk_frames=%d, %s: %.0f frames/s  (file read %.1f ms, enqueue %.1f ms, GPU wait %.1f ms, %.0f MB)'
              % (bulk, what, best, stats['file_read_s'] * 1e3, stats['enqueue_s'] * 1e3, stats['wait_s'] * 1e3,
                 stats['bytes'] / 1e6))
os.remove(path)
