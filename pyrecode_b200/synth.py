"""Synthetic detector frames of SURVEY.md section 8d (bench.py, tests): a fixed-pattern dark frame, Gaussian read noise
and sparse electron events -- Bernoulli(0.02) single pixels for L1, puddles (a centre plus random E / S / SE neighbours)
at p = 0.0075 for L2 and p = 0.005 ("low dose") for L4.  Pure numpy; seeded."""
import numpy as np


def synth_dark(ny, nx, seed=7):
    rng = np.random.default_rng(seed)
    return (100 + rng.integers(0, 8, size=(ny, nx))).astype(np.uint16)


def synth_frames(kind, nz, ny, nx, dark, seed=1234, bit_depth=12):
    """kind: 'l1' Bernoulli(0.02) events; 'l2' puddle model p=0.0075; 'l4' low-dose p=0.005."""
    rng = np.random.default_rng(seed)
    scale = (1 << bit_depth) / 4096.0
    vmax = (1 << bit_depth) - 1
    out = np.empty((nz, ny, nx), dtype=np.uint16)
    for z in range(nz):
        f = dark.astype(np.int64) + np.rint(rng.normal(0.0, 3.0, size=(ny, nx))).astype(np.int64)
        if kind == 'l1':
            ev = rng.random((ny, nx)) < 0.02
            f[ev] += (rng.integers(50, 1000, size=int(ev.sum())) * scale).astype(np.int64)
        else:
            p = 0.0075 if kind == 'l2' else 0.005
            ev = rng.random((ny, nx)) < p
            amp = np.zeros((ny, nx), dtype=np.int64)
            amp[ev] = (rng.integers(50, 1000, size=int(ev.sum())) * scale).astype(np.int64)
            f += amp
            for dy, dx in ((0, 1), (1, 0), (1, 1)):
                nb = np.zeros_like(ev)
                nb[dy:, dx:] = ev[:ny - dy, :nx - dx]
                nb &= rng.random((ny, nx)) < 0.5
                f[nb] += (rng.integers(25, 500, size=int(nb.sum())) * scale).astype(np.int64)
        np.clip(f, 0, vmax, out=f)
        out[z] = f.astype(np.uint16)
    return out
