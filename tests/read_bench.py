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
F = 16
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
