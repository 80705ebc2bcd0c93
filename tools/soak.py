"""Soak run of the pipelined write path against the CPU oracle (test infrastructure; compute-sanitizer is not available on
the GPU pool, so this is the race check we can run): random geometries, occupancies, thresholds, levels and batch sizes,
3 or 4 batches in flight with several rounds of slot reuse, EVERY record of every slot's last launch inflated with stock
zlib and compared with the oracle.
usage: python tools/soak.py [iterations] [seed]   (needs a GPU)"""
import sys
import time
import zlib

sys.path.insert(0, '.')
import numpy as np
import torch

from oracle import oracle as orc
from pyrecode_b200.engine import WriteEngine


def frames_of(rng, n, ny, nx, dark, p, grow, bit_depth):
    vmax = (1 << bit_depth) - 1
    out = np.empty((n, ny, nx), dtype=np.uint16)
    for z in range(n):
        f = dark.astype(np.int64) + np.rint(rng.normal(0.0, 3.0, size=(ny, nx))).astype(np.int64)
        ev = rng.random((ny, nx)) < p
        f[ev] += rng.integers(50, vmax // 4, size=int(ev.sum()))
        for dy, dx in ((0, 1), (1, 0), (1, 1), (1, -1)):
            nb = np.zeros_like(ev)
            if dx >= 0:
                nb[dy:, dx:] = ev[:ny - dy, :nx - dx]
            else:
                nb[dy:, :dx] = ev[:ny - dy, -dx:]
            nb &= rng.random((ny, nx)) < grow
            f[nb] += rng.integers(25, vmax // 8, size=int(nb.sum()))
        np.clip(f, 0, vmax, out=f)
        out[z] = f.astype(np.uint16)
    return out


def check(rec, offs, counts, F, level, first_id, expect, shift, tag):
    rec = memoryview(rec)
    for i in range(F):
        r = bytes(rec[int(offs[i]):int(offs[i + 1])])
        m_ref, v_ref, n_ref = expect[(i + shift) % len(expect)]
        hdr = np.frombuffer(r[:16 if level <= 2 else 8], dtype='<u4')
        assert hdr[0] == first_id + i, (tag, i, 'frame id')
        if level <= 2:
            assert len(r) == 16 + hdr[1] + hdr[2], (tag, i, 'record length')
            assert zlib.decompress(r[16:16 + hdr[1]]) == m_ref, (tag, i, 'map')
            assert zlib.decompress(r[16 + hdr[1]:]) == v_ref and hdr[3] == len(v_ref), (tag, i, 'values')
        else:
            assert len(r) == 8 + hdr[1], (tag, i, 'record length')
            assert zlib.decompress(r[8:]) == m_ref, (tag, i, 'centroid map')
        assert int(counts[i]) == n_ref, (tag, i, 'count')


def main(iters=12, seed=2026):
    rng = np.random.default_rng(seed)
    geoms = [(4096, 4096), (4096, 4096), (2048, 4096), (1000, 1200), (1536, 2048), (777, 4100)]
    t00 = time.time()
    for it in range(iters):
        ny, nx = geoms[int(rng.integers(len(geoms)))]
        level = int(rng.choice([2, 2, 1, 4]))
        b = int(rng.choice([12, 12, 10, 14]))
        p = float(rng.choice([0.001, 0.004, 0.0075, 0.015, 0.03]))
        grow = float(rng.choice([0.2, 0.5, 0.8]))
        eps = int(rng.choice([6, 12, 20]))            # 6 = two sigma of the read noise: speckle everywhere
        slots = int(rng.choice([3, 4]))
        F = int(rng.choice([5, 8, 16])) if ny * nx > 8e6 else int(rng.choice([7, 16, 32]))
        stat, cent = int(rng.choice([1, 2])), int(rng.choice([0, 2, 3]))
        distinct = 3
        dark = orc.synth_dark(ny, nx)
        frames = frames_of(rng, distinct, ny, nx, dark, p, grow, b)
        thr = orc.make_threshold(dark, eps)
        expect = [orc.reduce_frame(f, thr, level, b, l2_statistics=stat, l4_centroiding=cent) for f in frames]
        eng = WriteEngine(ny, nx, 2, b, level, 1, stat, cent, 1, max_frames=F,
                          records_capacity=F * (ny * nx * 2 // 2 + 4096), n_slots=slots)
        eng.set_threshold(dark, eps)
        host = torch.empty((slots, F, ny, nx), dtype=torch.uint16).pin_memory()
        hv = host.numpy()
        for k in range(slots):
            for i in range(F):
                hv[k, i] = frames[(i + k) % distinct]
        d_frames = host.to(eng.dev)
        torch.cuda.synchronize()
        cur = torch.cuda.current_stream()
        rounds = int(rng.integers(3, 7))
        for sl in eng.slots:
            sl.stream.wait_stream(cur)
        for s in range(rounds * slots):
            sl = eng.slots[s % slots]
            with torch.cuda.stream(sl.stream):
                eng.launch(d_frames[s % slots], F, 1000 + s * F, s % slots)
        for sl in eng.slots:
            cur.wait_stream(sl.stream)
        torch.cuda.synchronize()
        fg = 0
        for k, sl in enumerate(eng.slots):
            assert int(sl.status.cpu()[0]) == 0, 'status'
            offs = sl.offsets.cpu().numpy()
            counts = sl.counts.cpu().numpy()
            rec = sl.records[:int(offs[F])].cpu().numpy()
            check(rec, offs, counts, F, level, 1000 + ((rounds - 1) * slots + k) * F, expect, k, 'it %d slot %d' % (it, k))
            fg += int(counts.sum())
        print('it %2d ok: %4dx%4d L%d b=%d p=%.4f grow=%.1f eps=%d stat=%d cent=%d F=%d slots=%d rounds=%d  counts/frame %.0f'
              % (it, ny, nx, level, b, p, grow, eps, stat, cent, F, slots, rounds, fg / (slots * F)), flush=True)
        del eng, d_frames, host
        torch.cuda.empty_cache()
    print('SOAK_OK %d iterations in %.0f s' % (iters, time.time() - t00))


if __name__ == '__main__':
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 12, int(sys.argv[2]) if len(sys.argv) > 2 else 2026)
