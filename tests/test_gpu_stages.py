"""Parity of every CUDA stage against the CPU oracle (oracle/) and the golden fixtures, through the C ABI."""
import os
import zlib

import numpy as np
import pytest

from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def first_diff(a, b):
    a = np.frombuffer(a, np.uint8) if isinstance(a, (bytes, bytearray)) else np.asarray(a).ravel()
    b = np.frombuffer(b, np.uint8) if isinstance(b, (bytes, bytearray)) else np.asarray(b).ravel()
    if a.size != b.size:
        return 'size %d != %d' % (a.size, b.size)
    d = np.nonzero(a != b)[0]
    return 'equal' if d.size == 0 else 'first diff at %d (%r vs %r), %d diffs' % (d[0], a[d[0]], b[d[0]], d.size)


def make_frames(rng, n, ny, nx, dtype, vmax, occ, dark_lo=0, dark_hi=4):
    dark = rng.integers(dark_lo, dark_hi, size=(ny, nx)).astype(dtype)
    f = np.broadcast_to(dark, (n, ny, nx)).astype(np.int64)
    m = rng.random((n, ny, nx)) < occ
    f[m] += rng.integers(1, vmax + 1, size=int(m.sum()))
    return np.clip(f, 0, vmax).astype(dtype), dark


def engine(ny, nx, itemsize, b, level, mode=1, l2=0, l4=0, clevel=1, F=4):
    from pyrecode_b200.engine import WriteEngine
    return WriteEngine(ny, nx, itemsize, b, level, mode, l2, l4, clevel, max_frames=F)


SHAPES = [(1, 1), (3, 5), (37, 53), (64, 96), (128, 256), (512, 512), (300, 1000)]


@pytest.mark.parametrize('ny,nx', SHAPES)
@pytest.mark.parametrize('b,itemsize', [(12, 2), (16, 2), (9, 2), (8, 1), (5, 1), (1, 1)])
def test_l1_reduce(ny, nx, b, itemsize):
    rng = np.random.default_rng(ny * 1000 + nx + b)
    dt = np.uint8 if itemsize == 1 else np.uint16
    frames, dark = make_frames(rng, 3, ny, nx, dt, (1 << b) - 1, 0.1)
    frames[1] = 0                      # empty frame
    frames[2] = (1 << b) - 1           # every pixel foreground (dark < max)
    eng = engine(ny, nx, itemsize, b, 1, F=3)
    eng.set_threshold(dark, 2)
    maps, packed, counts = eng.reduce(frames)
    thr = orc.make_threshold(dark, 2, dtype=dt)
    for f in range(3):
        m, v, n = orc.reduce_frame(frames[f].astype(np.uint16), thr.astype(np.uint16), 1, b)
        assert counts[f] == n
        assert maps[f] == m, 'map frame %d: %s' % (f, first_diff(maps[f], m))
        assert packed[f] == v, 'packed frame %d: %s' % (f, first_diff(packed[f], v))


def test_l1_threshold_wraps():
    # dark + eps wraps in uint16: reproduce the wrapped threshold
    rng = np.random.default_rng(5)
    dark = rng.integers(65000, 65536, size=(32, 64)).astype(np.uint16)
    frames = rng.integers(0, 65536, size=(2, 32, 64)).astype(np.uint16)
    eng = engine(32, 64, 2, 16, 1, F=2)
    eng.set_threshold(dark, 700)
    maps, packed, counts = eng.reduce(frames)
    thr = (dark + np.uint16(700)).astype(np.uint16)
    for f in range(2):
        m, v, n = orc.reduce_frame(frames[f], thr, 1, 16)
        assert maps[f] == m and packed[f] == v and counts[f] == n


def test_l1_full_frame_4096():
    dark = orc.synth_dark(4096, 4096)
    frames = orc.synth_frames('l1', 2, 4096, 4096, dark, seed=1234)
    eng = engine(4096, 4096, 2, 12, 1, F=2)
    eng.set_threshold(dark, 20)
    maps, packed, counts = eng.reduce(frames)
    thr = orc.make_threshold(dark, 20)
    for f in range(2):
        m, v, n = orc.reduce_frame(frames[f], thr, 1, 12)
        assert counts[f] == n
        assert maps[f] == m, first_diff(maps[f], m)
        assert packed[f] == v, first_diff(packed[f], v)


def test_l1_golden_reference_streams(gold_dir):
    # GPU streams == the streams the unmodified reference wrote (inflated with stock zlib)
    z = np.load(os.path.join(gold_dir, 'gold_a_input.npz'))
    data, dark, eps = z['data'], z['dark'], int(z['eps'])
    nz, ny, nx = data.shape
    eng = engine(ny, nx, 2, 12, 1, F=nz)
    eng.set_threshold(dark, eps)
    maps, packed, counts = eng.reduce(data)
    recs = []
    for node in range(3):
        recs += orc.parse_part_file(os.path.join(gold_dir, 'gold_a.rc1_part%03d' % node))[1]
    assert [r['frame_id'] for r in recs] == list(range(nz))
    for f in range(nz):
        assert maps[f] == recs[f]['map'] and packed[f] == recs[f]['vals']


@pytest.mark.parametrize('ny,nx', SHAPES)
def test_l3_reduce(ny, nx):
    rng = np.random.default_rng(ny + nx)
    frames, dark = make_frames(rng, 2, ny, nx, np.uint16, 4095, 0.2)
    eng = engine(ny, nx, 2, 12, 3, F=2)
    eng.set_threshold(dark, 1)
    maps, packed, counts = eng.reduce(frames)
    thr = orc.make_threshold(dark, 1)
    for f in range(2):
        m, v, n = orc.reduce_frame(frames[f], thr, 3, 12)
        assert maps[f] == m and counts[f] == n and packed[f] == b''


@pytest.mark.parametrize('ny,nx,occ', [(1, 1, 1.0), (3, 5, 0.5), (37, 53, 0.3), (64, 96, 0.1), (128, 256, 0.45),
                                       (512, 512, 0.02), (512, 512, 0.6), (300, 1000, 0.15), (200, 200, 1.0)])
def test_ccl_labels(ny, nx, occ):
    rng = np.random.default_rng(int(occ * 100) + ny)
    n = 2
    binary = rng.random((n, ny, nx)) < occ
    eng = engine(ny, nx, 2, 12, 2, F=n)
    maps = [orc.pack_map(binary[f]).tobytes() for f in range(n)]
    labels, k = eng.labels(maps)
    for f in range(n):
        olab, ok = orc.label8(binary[f])
        assert k[f] == ok
        assert np.array_equal(labels[f], olab), first_diff(labels[f], olab)


def test_ccl_labels_structured():
    # serpentine / spiral / comb shapes: long union chains across tiles
    ny, nx = 96, 300
    img = np.zeros((ny, nx), dtype=bool)
    img[::4, :] = True
    img[2::8, -1] = True
    img[6::8, 0] = True
    img2 = np.zeros((ny, nx), dtype=bool)
    img2[:, ::2] = True
    img2[-1, :] = True
    img3 = np.zeros((ny, nx), dtype=bool)
    for i in range(min(ny, nx)):
        img3[i, i] = True
        img3[i, nx - 1 - i] = True
    eng = engine(ny, nx, 2, 12, 2, F=3)
    labels, k = eng.labels([orc.pack_map(x).tobytes() for x in (img, img2, img3)])
    for f, x in enumerate((img, img2, img3)):
        olab, ok = orc.label8(x)
        assert k[f] == ok and np.array_equal(labels[f], olab)


def test_ccl_golden_scipy(gold_dir):
    z = np.load(os.path.join(gold_dir, 'gold_d_ccl.npz'))
    for tag in ('small', 'tall', 'dense'):
        frame, lab = z[tag + '_frame'], z[tag + '_labels']
        ny, nx = frame.shape
        eng = engine(ny, nx, 2, 16, 2, F=1)
        labels, k = eng.labels([orc.pack_map(frame > 0).tobytes()])
        assert k[0] == lab.max() and np.array_equal(labels[0], lab), tag


@pytest.mark.parametrize('stat', [0, 2])
@pytest.mark.parametrize('ny,nx,b,occ', [(37, 53, 12, 0.2), (128, 256, 12, 0.1), (512, 512, 12, 0.03), (64, 96, 16, 0.4),
                                         (300, 1000, 9, 0.15)])
def test_l2_reduce(ny, nx, b, occ, stat):
    rng = np.random.default_rng(ny + b + stat)
    frames, dark = make_frames(rng, 2, ny, nx, np.uint16, (1 << b) - 1, occ)
    eng = engine(ny, nx, 2, b, 2, l2=stat, F=2)
    eng.set_threshold(dark, 3)
    maps, packed, counts = eng.reduce(frames)
    thr = orc.make_threshold(dark, 3)
    for f in range(2):
        m, v, n = orc.reduce_frame(frames[f], thr, 2, b, l2_statistics=stat)
        assert counts[f] == n
        assert maps[f] == m
        assert packed[f] == v, first_diff(packed[f], v)



@pytest.mark.parametrize('stat', [0, 2])
@pytest.mark.parametrize('ny,nx,b,occ', [(37, 53, 8, 0.2), (128, 256, 5, 0.1), (512, 512, 8, 0.03), (300, 1000, 5, 0.15),
                                         (257, 4096, 8, 0.02)])
def test_l2_reduce_uint8_source(ny, nx, b, occ, stat):
    """itemsize-1 sources (what misc.map_dtype selects for bit depths <= 8, pyrecode/misc.py:41-71): raw values and
    thresholds are bytes; the per-puddle sum is kept modulo 2^b by the packer"""
    rng = np.random.default_rng(ny + b + stat + 77)
    frames, dark = make_frames(rng, 2, ny, nx, np.uint8, (1 << b) - 1, occ)
    eng = engine(ny, nx, 1, b, 2, l2=stat, F=2)
    eng.set_threshold(dark, 3)
    maps, packed, counts = eng.reduce(frames)
    thr = orc.make_threshold(dark, 3, dtype=np.uint8)
    for f in range(2):
        m, v, n = orc.reduce_frame(frames[f].astype(np.uint16), thr.astype(np.uint16), 2, b, l2_statistics=stat)
        assert counts[f] == n
        assert maps[f] == m
        assert packed[f] == v, first_diff(packed[f], v)


def _cross_tile_frames(ny, nx, rng, vmax):
    """Frames whose puddles cross the 32768-pixel tile boundaries of the labelling kernel in every way:
    full-height vertical lines, diagonals, U shapes closed only in a later tile, combs, one tile dense enough
    to overflow the shared-memory labelling next to sparse tiles."""
    f = np.zeros((4, ny, nx), np.uint16)
    val = lambda shape: rng.integers(1, vmax + 1, size=shape).astype(np.uint16)
    # 0: vertical lines + both diagonals over the whole frame
    for c in range(3, nx, 37):
        f[0, :, c] = val(ny)
    d = np.arange(min(ny, nx))
    f[0, d, d] = val(d.size)
    f[0, d, nx - 1 - d] = val(d.size)
    # 1: U shapes: two separate columns joined only at the bottom row, and an inverted comb from the top row
    for c in range(2, nx - 8, 40):
        f[1, : ny - 1, c] = val(ny - 1)
        f[1, : ny - 1, c + 4] = val(ny - 1)
        f[1, ny - 1, c:c + 5] = val(5)
    f[1, 0, :] = 0
    # 2: sparse random everywhere, but a dense band (overflow tile) in the middle rows touching its neighbours
    m = rng.random((ny, nx)) < 0.02
    f[2][m] = val(int(m.sum()))
    r0 = ny // 2
    band = rng.random((min(40, ny - r0), nx)) < 0.55
    f[2, r0:r0 + band.shape[0]][band] = val(int(band.sum()))
    # 3: horizontal serpentine: rows fully set every 3rd row, joined alternately at the left / right end
    for i, r in enumerate(range(0, ny - 3, 3)):
        f[3, r, :] = val(nx)
        f[3, r + 1: r + 3, 0 if i % 2 else nx - 1] = val(2)
    return f


@pytest.mark.parametrize('stat', [0, 2])
@pytest.mark.parametrize('ny,nx', [(512, 512), (300, 1000), (257, 4096), (1030, 96), (200, 333), (9, 9000)])
def test_l2_cross_tile_puddles(ny, nx, stat):
    rng = np.random.default_rng(ny * 7 + nx + stat)
    frames = _cross_tile_frames(ny, nx, rng, 4095)
    dark = np.zeros((ny, nx), np.uint16)
    eng = engine(ny, nx, 2, 12, 2, l2=stat, F=4)
    eng.set_threshold(dark, 0)
    maps, packed, counts = eng.reduce(frames)
    for f in range(4):
        m, v, n = orc.reduce_frame(frames[f], dark, 2, 12, l2_statistics=stat)
        assert counts[f] == n, 'frame %d: %d puddles, expected %d' % (f, counts[f], n)
        assert maps[f] == m
        assert packed[f] == v, 'frame %d: %s' % (f, first_diff(packed[f], v))


@pytest.mark.parametrize('ny,nx', [(512, 512), (300, 1000), (1030, 96)])
def test_l4_cross_tile_puddles(ny, nx):
    rng = np.random.default_rng(ny * 11 + nx)
    frames = _cross_tile_frames(ny, nx, rng, 4095)
    dark = np.zeros((ny, nx), np.uint16)
    eng = engine(ny, nx, 2, 12, 4, F=4)
    eng.set_threshold(dark, 0)
    cents = eng.centroids(frames)
    maps, packed, counts = eng.reduce(frames)
    for f in range(4):
        lab, k = orc.label8(frames[f] > 0)
        oc = orc.l4_centroids(lab, frames[f], k, 0)
        assert cents[f].shape == oc.shape
        assert np.array_equal(cents[f].view(np.uint32), oc.view(np.uint32)), first_diff(cents[f].view(np.uint32), oc.view(np.uint32))
        m, v, n = orc.reduce_frame(frames[f], dark, 4, 12)
        assert counts[f] == n == k
        assert maps[f] == m, first_diff(maps[f], m)



@pytest.mark.parametrize('level,stat', [(2, 0), (2, 2), (4, 0)])
@pytest.mark.parametrize('ny,nx,b', [(512, 512, 8), (300, 1000, 5), (257, 4096, 8)])
def test_cross_tile_puddles_uint8_source(ny, nx, b, level, stat):
    rng = np.random.default_rng(ny * 13 + nx + stat + level)
    frames = _cross_tile_frames(ny, nx, rng, (1 << b) - 1).astype(np.uint8)
    dark = np.zeros((ny, nx), np.uint8)
    eng = engine(ny, nx, 1, b, level, l2=stat, F=4)
    eng.set_threshold(dark, 0)
    maps, packed, counts = eng.reduce(frames)
    for f in range(4):
        m, v, n = orc.reduce_frame(frames[f].astype(np.uint16), dark.astype(np.uint16), level, b, l2_statistics=stat)
        assert counts[f] == n, 'frame %d: %d puddles, expected %d' % (f, counts[f], n)
        assert maps[f] == m, 'frame %d: %s' % (f, first_diff(maps[f], m))
        if level == 2:
            assert packed[f] == v, 'frame %d: %s' % (f, first_diff(packed[f], v))


def test_l2_synthetic_4096():
    dark = orc.synth_dark(4096, 4096)
    frames = orc.synth_frames('l2', 1, 4096, 4096, dark, seed=1234)
    eng = engine(4096, 4096, 2, 12, 2, F=1)
    eng.set_threshold(dark, 20)
    maps, packed, counts = eng.reduce(frames)
    m, v, n = orc.reduce_frame(frames[0], orc.make_threshold(dark, 20), 2, 12)
    assert counts[0] == n and maps[0] == m and packed[0] == v


@pytest.mark.parametrize('mode', [0, 2, 3])
@pytest.mark.parametrize('ny,nx,b,occ', [(37, 53, 12, 0.2), (128, 256, 12, 0.1), (3000, 24, 16, 0.22), (64, 96, 16, 0.45)])
def test_l4_centroids_and_map(ny, nx, b, occ, mode):
    rng = np.random.default_rng(ny + b + mode)
    frames, dark = make_frames(rng, 2, ny, nx, np.uint16, (1 << b) - 1, occ)
    eng = engine(ny, nx, 2, b, 4, l4=mode, F=2)
    eng.set_threshold(dark, 3)
    thr = orc.make_threshold(dark, 3)
    cents = eng.centroids(frames)
    maps, packed, counts = eng.reduce(frames)
    for f in range(2):
        binary = orc.binarize(frames[f], thr)
        lab, k = orc.label8(binary)
        oc = orc.l4_centroids(lab, frames[f], k, mode)
        assert cents[f].shape == oc.shape
        assert np.array_equal(cents[f].view(np.uint32), oc.view(np.uint32)), first_diff(cents[f].view(np.uint32), oc.view(np.uint32))
        m, v, n = orc.reduce_frame(frames[f], thr, 4, b, l4_centroiding=mode)
        assert counts[f] == n == k
        assert maps[f] == m, first_diff(maps[f], m)


def test_l4_golden_live_centroids(gold_dir):
    # bit-exact against the live reference's get_centroids_2D_nb, including float32 rounding (sums > 2^24)
    z = np.load(os.path.join(gold_dir, 'gold_d_ccl.npz'))
    for tag in ('small', 'tall', 'dense'):
        frame, cen = z[tag + '_frame'], z[tag + '_centroids']
        ny, nx = frame.shape
        eng = engine(ny, nx, 2, 16, 4, F=1)
        eng.set_threshold(np.zeros((ny, nx), np.uint16), 0)
        got = eng.centroids(frame[None])[0]
        assert got.shape == cen.shape and np.array_equal(got.view(np.uint32), cen.view(np.uint32)), tag


def payload_cases():
    rng = np.random.default_rng(11)
    cases = {'empty': b'', 'one': b'\x07', 'zeros': bytes(100000), 'ff': b'\xff' * 40000}
    for occ in (0.0005, 0.02, 0.5):
        cases['map%g' % occ] = np.packbits(rng.random(1 << 20) < occ, bitorder='little').tobytes()
    cases['rand16384'] = rng.integers(0, 256, 16384, dtype=np.uint8).tobytes()
    cases['rand16385'] = rng.integers(0, 256, 16385, dtype=np.uint8).tobytes()
    cases['rand50001'] = rng.integers(0, 256, 50001, dtype=np.uint8).tobytes()
    cases['packed12'] = orc.bit_pack(rng.integers(1, 1000, 70000).astype(np.uint16), 12).tobytes()
    cases['marker'] = b'\x00\x00\xff\xff' * 5000 + rng.integers(0, 256, 30000, dtype=np.uint8).tobytes()
    skew = np.concatenate([np.full(int(1.6 ** i) + 1, i, np.uint8) for i in range(24)])
    rng.shuffle(skew)
    cases['skew'] = skew.tobytes()
    return cases


@pytest.mark.parametrize('level', [1, 9])
def test_deflate_long_streams(ctx, level):
    """streams of more than 256 chunks (the per-stream prefix / Adler-32 combine loops in tiles of 256 chunks) and a
    very sparse one (zero runs far beyond one composite run token); stock zlib verifies payload and checksum"""
    from pyrecode_b200.engine import deflate_batch, inflate_batch
    rng = np.random.default_rng(77)
    big = np.packbits(rng.random(5 * (1 << 23) + 1000) < 0.03, bitorder='little').tobytes()        # 5.2 MB: 321 chunks
    sparse = np.packbits(rng.random(1 << 25) < 0.00002, bitorder='little').tobytes()               # 4 MB, ~80 set bits
    mixed = big[:700001] + bytes(300000) + rng.integers(0, 256, 123457, dtype=np.uint8).tobytes() + sparse[:2000001]
    payloads = [big, sparse, mixed]
    comp = deflate_batch(ctx, payloads, level)
    for c, d in zip(comp, payloads):
        assert zlib.decompress(c) == d
    assert len(comp[1]) < 60000           # 256 chunks: what is left is one block header (~200 B) per chunk
    back, st = inflate_batch(ctx, comp, max(len(p) for p in payloads))
    assert not st.any() and all(b == d for b, d in zip(back, payloads))


@pytest.mark.parametrize('level', [1, 0, 9])
def test_deflate_inflates_with_stock_zlib(ctx, level):
    from pyrecode_b200.engine import deflate_batch
    cases = payload_cases()
    names = list(cases)
    comp = deflate_batch(ctx, [cases[k] for k in names], level)
    for k, c in zip(names, comp):
        assert zlib.decompress(c) == cases[k], k
        assert len(c) <= ctx.deflate_bound(len(cases[k])), k
    if level == 1:
        # ratio sanity on the data this path exists for: not worse than zlib level 1 on a 2 % map
        i = names.index('map0.02')
        assert len(comp[i]) < len(zlib.compress(cases['map0.02'], 1))


def test_inflate_stock_zlib_streams(ctx):
    from pyrecode_b200.engine import inflate_batch
    cases = payload_cases()
    streams, want = [], []
    for k, d in cases.items():
        for lvl in (0, 1, 6, 9):
            streams.append(zlib.compress(d, lvl))
            want.append(d)
    co = zlib.compressobj(6, zlib.DEFLATED, 15, 9, zlib.Z_FIXED)
    streams.append(co.compress(cases['map0.02']) + co.flush())
    want.append(cases['map0.02'])
    co = zlib.compressobj(1)
    s = b''
    for i in range(0, len(cases['skew']), 10000):          # sync-flushed foreign stream with other chunking
        s += co.compress(cases['skew'][i:i + 10000]) + co.flush(zlib.Z_SYNC_FLUSH)
    streams.append(s + co.flush())
    want.append(cases['skew'])
    out, st = inflate_batch(ctx, streams, max(len(w) for w in want))
    for i, (o, w) in enumerate(zip(out, want)):
        assert st[i] == 0, (i, st[i])
        assert o == w, (i, first_diff(o, w))


def test_inflate_own_streams_and_errors(ctx):
    from pyrecode_b200.engine import deflate_batch, inflate_batch
    cases = payload_cases()
    names = list(cases)
    # level 1: one code per batch (identical chunk headers across streams); 9: one code per stream; 0: stored pieces
    for level in (9, 0, 1):
        comp = deflate_batch(ctx, [cases[k] for k in names], level)
        out, st = inflate_batch(ctx, comp, max(len(v) for v in cases.values()))
        for k, o, s in zip(names, out, st):
            assert s == 0 and o == cases[k], (level, k)
    bad = bytearray(comp[names.index('map0.02')])
    bad[len(bad) // 2] ^= 0x55
    trunc = comp[names.index('map0.5')][:1000]
    out, st = inflate_batch(ctx, [bytes(bad), trunc, b'\x00' * 20], 1 << 17)
    assert all(s != 0 for s in st)
    out, st = inflate_batch(ctx, [comp[names.index('zeros')]], 1000)       # output capacity too small
    assert st[0] != 0


def test_bit_pack_unpack_flat(ctx, gold_dir):
    import torch
    z = np.load(os.path.join(gold_dir, 'gold_d_pack.npz'))
    vals = z['vals']
    dv = torch.from_numpy(vals).cuda()
    for b in range(1, 17):
        ref = z['b%d' % b]
        out = torch.zeros((len(ref) + 7) // 4 * 4, dtype=torch.uint8, device='cuda')
        ctx.bit_pack(b, dv, len(vals), out)
        back = torch.zeros(len(vals), dtype=torch.int64, device='cuda')
        ctx.bit_unpack(b, out, len(vals), back)
        torch.cuda.synchronize()
        assert np.array_equal(out.cpu().numpy()[:len(ref)], ref), b
        assert np.array_equal(back.cpu().numpy().astype(np.uint64), (vals & ((1 << b) - 1)).astype(np.uint64)), b


def parse_records(buf, offs, level, mode, map_bytes):
    out = []
    for i in range(len(offs) - 1):
        r = bytes(buf[int(offs[i]):int(offs[i + 1])])
        fid = int.from_bytes(r[:4], 'little')
        if mode == 1:
            n1 = int.from_bytes(r[4:8], 'little')
            if level <= 2:
                n2 = int.from_bytes(r[8:12], 'little')
                npk = int.from_bytes(r[12:16], 'little')
                assert len(r) == 16 + n1 + n2
                m, v = zlib.decompress(r[16:16 + n1]), zlib.decompress(r[16 + n1:])
                assert len(v) == npk
            else:
                assert len(r) == 8 + n1
                m, v = zlib.decompress(r[8:]), b''
        else:
            if level <= 2:
                npk = int.from_bytes(r[4:8], 'little')
                assert len(r) == 8 + map_bytes + npk
                m, v = r[8:8 + map_bytes], r[8 + map_bytes:]
            else:
                assert len(r) == 4 + map_bytes
                m, v = r[4:], b''
        out.append((fid, m, v))
    return out


@pytest.mark.parametrize('level', [1, 2, 3, 4])
@pytest.mark.parametrize('mode', [1, 0])
@pytest.mark.parametrize('ny,nx', [(37, 53), (512, 512), (300, 1000)])
def test_reduce_compress_records(level, mode, ny, nx):
    rng = np.random.default_rng(level * 10 + mode + ny)
    n = 5
    frames, dark = make_frames(rng, n, ny, nx, np.uint16, 4095, 0.05)
    frames[3] = 0
    eng = engine(ny, nx, 2, 12, level, mode=mode, F=n)
    eng.set_threshold(dark, 4)
    rec, offs, counts, h2d, d2h = eng.reduce_compress(frames, first_frame_id=17)
    thr = orc.make_threshold(dark, 4)
    got = parse_records(rec, offs, level, mode, (ny * nx + 7) // 8)
    for f in range(n):
        m, v, cnt = orc.reduce_frame(frames[f], thr, level, 12)
        assert got[f][0] == 17 + f
        assert got[f][1] == m, (f, first_diff(got[f][1], m))
        assert got[f][2] == v, (f, first_diff(got[f][2], v))
        assert counts[f] == cnt


def test_reduce_compress_records_overflow():
    rng = np.random.default_rng(3)
    frames = rng.integers(0, 4096, size=(2, 256, 256)).astype(np.uint16)      # incompressible, all foreground
    from pyrecode_b200.engine import WriteEngine
    eng = WriteEngine(256, 256, 2, 12, 1, max_frames=2, records_capacity=4096)
    eng.set_threshold(np.zeros((256, 256), np.uint16), 0)
    with pytest.raises(ValueError):
        eng.reduce_compress(frames)


def test_unpack_golden_triples(gold_dir):
    from pyrecode_b200.engine import ReadEngine
    z = np.load(os.path.join(gold_dir, 'gold_d_unpack.npz'))
    ny, nx, b = int(z['ny']), int(z['nx']), int(z['b'])
    for level in (1, 3):
        eng = ReadEngine(ny, nx, 2, b, level, rc_operation_mode=0, max_frames=2)
        eng.load([z['map'].tobytes()] * 2, [z['packed'].tobytes()] * 2 if level == 1 else None)
        tri = eng.sparse()
        for t in tri:
            assert np.array_equal(t, z['triples_l%d' % level])


@pytest.mark.parametrize('ny,nx,b,itemsize', [(37, 53, 12, 2), (512, 512, 12, 2), (300, 1000, 16, 2), (64, 96, 7, 1),
                                              (128, 256, 8, 1)])
def test_read_path_roundtrip(ny, nx, b, itemsize):
    import torch
    from pyrecode_b200.engine import ReadEngine
    rng = np.random.default_rng(ny + b)
    dt = np.uint8 if itemsize == 1 else np.uint16
    n = 4
    frames, dark = make_frames(rng, n, ny, nx, dt, (1 << b) - 1, 0.08)
    frames[2] = 0
    eng = engine(ny, nx, itemsize, b, 1, F=n)
    eng.set_threshold(dark, 1)
    rec, offs, counts, _, _ = eng.reduce_compress(frames)
    cm, cv = [], []
    for i in range(n):
        r = bytes(rec[int(offs[i]):int(offs[i + 1])])
        n1 = int.from_bytes(r[4:8], 'little')
        cm.append(r[16:16 + n1])
        cv.append(r[16 + n1:])
    rd = ReadEngine(ny, nx, itemsize, b, 1, max_frames=n)
    rd.load(cm, cv)
    rd.check()
    thr = (dark + dt(1)).astype(dt)
    model = np.where(frames > thr, frames - thr, 0).astype(dt)
    total = torch.zeros(ny * nx, dtype=torch.int32, device='cuda')
    dense = rd.dense(total=total)
    torch.cuda.synchronize()
    assert np.array_equal(dense.cpu().numpy(), model)
    assert np.array_equal(total.cpu().numpy().reshape(ny, nx), model.astype(np.int64).sum(axis=0))
    tri = rd.sparse()
    for f in range(n):
        r, c = np.nonzero(frames[f] > thr)
        assert np.array_equal(tri[f][:, 0], r) and np.array_equal(tri[f][:, 1], c)
        assert np.array_equal(tri[f][:, 2], model[f][r, c])


def test_read_reference_written_file(gold_dir):
    # streams compressed by the reference (stock zlib, dynamic blocks) decode on the GPU to the input model
    from pyrecode_b200.engine import ReadEngine
    z = np.load(os.path.join(gold_dir, 'gold_a_input.npz'))
    data, dark, eps = z['data'], z['dark'], int(z['eps'])
    nz, ny, nx = data.shape
    h, recs = orc.parse_part_file(os.path.join(gold_dir, 'gold_a.rc1_part000'))
    rd = ReadEngine(ny, nx, 2, 12, 1, max_frames=len(recs))
    rd.load([r['cmap'] for r in recs], [r['cvals'] for r in recs])
    rd.check()
    dense = rd.dense().cpu().numpy()
    thr = dark + np.uint16(eps)
    for i, r in enumerate(recs):
        f = r['frame_id']
        assert np.array_equal(dense[i], np.where(data[f] > thr, data[f] - thr, 0))


def test_l4_synthetic_4096():
    # BASELINE config 4 at full size: low-dose frame, centroid map and puddle count against the oracle
    dark = orc.synth_dark(4096, 4096)
    frames = orc.synth_frames('l4', 1, 4096, 4096, dark, seed=4321)
    eng = engine(4096, 4096, 2, 12, 4, F=1)
    eng.set_threshold(dark, 20)
    maps, packed, counts = eng.reduce(frames)
    m, v, n = orc.reduce_frame(frames[0], orc.make_threshold(dark, 20), 4, 12)
    assert counts[0] == n and maps[0] == m


def test_randomized_geometries_all_levels():
    """seeded sweep over odd geometries (tile boundaries in the middle of rows, widths beyond the labelling halo,
    single rows / columns), bit depths, occupancies and statistics, every level against the oracle"""
    rng = np.random.default_rng(20261018)
    shapes = [(1, 70000 // 7), (2049, 17), (33, 1025), (64, 8448), (5, 8449), (700, 257), (129, 255), (1, 40000),
              (40000, 1), (96, 4096), (17, 33000)]
    for case in range(40):
        ny, nx = shapes[case % len(shapes)] if case < 16 else (int(rng.integers(1, 400)), int(rng.integers(1, 3000)))
        b = int(rng.choice([9, 12, 16, 8, 5]))
        isz, dt = (1, np.uint8) if b <= 8 else (2, np.uint16)
        occ = float(rng.choice([0.002, 0.02, 0.08, 0.3, 0.7]))
        level = int(rng.choice([1, 2, 2, 4, 4, 3]))
        l2 = int(rng.choice([0, 2]))
        l4 = int(rng.choice([0, 2, 3]))
        frames, dark = make_frames(rng, 2, ny, nx, dt, (1 << b) - 1, occ)
        eng = engine(ny, nx, isz, b, level, l2=l2, l4=l4, F=2)
        eng.set_threshold(dark, 2)
        maps, packed, counts = eng.reduce(frames)
        thr = orc.make_threshold(dark, 2, dtype=dt).astype(np.uint16)
        for f in range(2):
            m, v, n = orc.reduce_frame(frames[f].astype(np.uint16), thr, level, b, l2_statistics=l2, l4_centroiding=l4)
            tag = 'case %d: %dx%d b=%d occ=%g level=%d l2=%d l4=%d frame %d' % (case, ny, nx, b, occ, level, l2, l4, f)
            assert counts[f] == n, tag
            assert maps[f] == m, tag + ' map ' + first_diff(maps[f], m)
            if level <= 2:
                assert packed[f] == v, tag + ' values ' + first_diff(packed[f], v)


def test_read_path_full_frames_4096():
    """full-size property test of the read path: L1 records of 4096 x 4096 frames (129 chunks per map stream, decoded
    one chunk per lane) -> dense frames == where(frame > thr, frame - thr, 0); live-view sum == their sum"""
    import torch
    from pyrecode_b200.engine import ReadEngine, WriteEngine
    ny = nx = 4096
    F = 3
    dark = orc.synth_dark(ny, nx)
    frames = np.stack(orc.synth_frames('l1', F, ny, nx, dark, seed=99))
    thr = orc.make_threshold(dark, 20)
    we = WriteEngine(ny, nx, 2, 12, 1, 1, 0, 0, 1, max_frames=F, records_capacity=F * (ny * nx // 2))
    we.set_threshold(dark, 20)
    rec, offs, counts, _, _ = we.reduce_compress(frames)
    rec = bytes(rec)
    maps, vals = [], []
    for f in range(F):
        r = rec[int(offs[f]):int(offs[f + 1])]
        h = np.frombuffer(r[:16], '<u4')
        maps.append(r[16:16 + h[1]])
        vals.append(r[16 + h[1]:16 + h[1] + h[2]])
    del we
    re_ = ReadEngine(ny, nx, 2, 12, 1, 1, max_frames=F)
    re_.load(maps, vals)
    re_.check()
    total = torch.zeros(ny * nx, dtype=torch.int32, device='cuda')
    dense = re_.dense(total=total).cpu().numpy()
    want = np.where(frames > thr, frames - thr, 0).astype(np.uint16)
    assert np.array_equal(dense, want)
    assert np.array_equal(total.cpu().numpy().reshape(ny, nx).astype(np.int64), want.astype(np.int64).sum(0))
