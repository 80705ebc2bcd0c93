"""Times the decode of reference-style streams (ONE stock-zlib block sequence per stream, no sync markers -- what
pyrecode's own writer produces) on the GPU: every stream is decoded by one lane of k_inflate_serial through a 32 KiB
shared-memory history ring, so the figure of merit is streams in flight, not one stream's latency.
usage: python tools/foreign_inflate.py [n_streams ...]   (needs a GPU)"""
import sys
import time
import zlib

sys.path.insert(0, '.')
import numpy as np
import torch

from pyrecode_b200._native import Context
from pyrecode_b200.engine import _stage_streams
from pyrecode_b200.synth import synth_dark, synth_frames
from oracle import oracle as orc


def main():
    counts = [int(a) for a in sys.argv[1:]] or [32, 512]
    ny = nx = 4096
    dark = synth_dark(ny, nx)
    frames = synth_frames('l2', 4, ny, nx, dark, seed=1234, bit_depth=12)
    thr = orc.make_threshold(dark, 20)
    maps = [np.packbits((f > thr).reshape(-1), bitorder='little').tobytes() for f in frames]
    comp = [zlib.compress(m, 1) for m in maps]
    t0 = time.perf_counter()
    for c in comp:
        zlib.decompress(c)
    core_ms = 1e3 * (time.perf_counter() - t0) / len(comp)
    ctx = Context(0)
    cap = len(maps[0])
    stride = (cap + 15) // 16 * 16 + 16
    for n in counts:
        streams = [comp[i % len(comp)] for i in range(n)]
        d_in, d_off, d_sz, _ = _stage_streams(ctx, streams)
        out = ctx.empty(n * stride + 64)
        out_bytes = ctx.zeros(n, torch.int32)
        status = ctx.zeros(n, torch.int32)
        best = None
        for rep in range(2):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ws = ctx.inflate_zlib(d_in, d_off, d_sz, n, out, stride, out_bytes, status)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            best = ms if best is None else min(best, ms)
        assert not status.cpu().numpy().any()
        o = out.cpu().numpy()
        for i in (0, n - 1):
            assert o[i * stride:i * stride + cap].tobytes() == maps[i % len(maps)]
        print('streams %d  inflated MiB each %.1f  gpu_ms %.1f  -> %.0f streams/s (%.2f GB/s inflated); one host core '
              '%.2f ms per stream -> %.0f streams/s; gpu = %.1f cores' %
              (n, cap / 2 ** 20, best, n / best * 1e3, n * cap / best / 1e6, core_ms, 1e3 / core_ms,
               n / best * core_ms), flush=True)


if __name__ == '__main__':
    main()
