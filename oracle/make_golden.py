"""Generate tests/golden/ by running the UNMODIFIED reference in this container.

Run once in the build container (the only place /root/reference exists):

    python oracle/make_golden.py

It imports pyrecode straight from /root/reference (read-only; no bytecode is
written) with the native extension compiled by oracle/Makefile into oracle/_ref/,
drives the reference's own ReCoDeWriter / ReCoDeReader / merge_parts / numba and
C kernels on small seeded inputs, asserts that the CPU oracle (oracle/oracle.py)
reproduces every output bit for bit, and stores inputs + reference outputs as
fixtures.  The fixtures travel to the GPU box; the reference does not.

The only deviation from "as shipped" is ``numpy.int = int`` injected before the
import (SURVEY B-8: the reference uses an alias removed from numpy >= 1.24).
"""
import io
import os
import sys
import shutil
import tempfile
import contextlib

import numpy as np

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(ROOT, 'tests', 'golden')
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(HERE, '_ref'))
sys.path.insert(0, '/root/reference')

np.int = int  # noqa  (SURVEY B-8)

import scipy.ndimage as nd  # noqa: E402
from oracle import oracle as orc  # noqa: E402

with contextlib.redirect_stdout(io.StringIO()):
    from pyrecode.recode_writer import ReCoDeWriter, _pack_binary_frame, _bit_pack  # noqa: E402
    from pyrecode.recode_reader import ReCoDeReader, merge_parts  # noqa: E402
    from pyrecode.params import InputParams  # noqa: E402
    from pyrecode.utils.converters import get_centroids_2D_nb  # noqa: E402
    import c_recode  # noqa: E402


def make_params(tmp, **kw):
    base = dict(l4_centroiding=0, source_file_type=0, num_frames=1, source_header_length=0,
                calibration_frame_offset=0, compression_scheme=0, calibration_file_type=0, compression_level=1,
                l2_statistics=0, calibration_threshold_epsilon=0, frame_offset=0, num_threads=1,
                rc_operation_mode=1, num_calibration_frames=1, reduction_level=1, keep_calibration_data=1,
                source_bit_depth=12, target_bit_depth=12, keep_part_files=0, num_rows=8, num_cols=8,
                source_data_type=0, target_data_type=0)
    base.update(kw)
    path = os.path.join(tmp, 'params.txt')
    with open(path, 'w') as f:
        for k, v in base.items():
            f.write('%s = %d\n' % (k, v))
    ip = InputParams()
    ip.load(path)
    return ip, base


def ref_write(tmp, name, data, dark, nodes, **kw):
    """reference writer, one instance per node (what ReCoDeNode does, recode_server.py:688-736)."""
    nz, ny, nx = data.shape
    files = []
    for node in range(nodes):
        ip, base = make_params(tmp, num_frames=nz, num_rows=ny, num_cols=nx, num_threads=nodes, **kw)
        with contextlib.redirect_stdout(io.StringIO()):
            w = ReCoDeWriter(name, dark_data=dark, output_directory=tmp, input_params=ip, mode='batch',
                             node_id=node)
            w.start()
            w.run(data)
            w.close()
        files.append(os.path.join(tmp, '%s.rc%d_part%03d' % (name, base['reduction_level'], node)))
    return files, base


def ref_read_all(path, intermediate):
    with contextlib.redirect_stdout(io.StringIO()):
        r = ReCoDeReader(path, is_intermediate=intermediate)
        r.open(print_header=False)
        out = {}
        nz = r.get_shape()[0]
        for _ in range(nz):
            f = r.get_next_frame()
            if f is None:
                break
            fid = list(f.keys())[0]
            coo = f[fid]['data']
            out[int(fid)] = (np.asarray(coo.row), np.asarray(coo.col), np.asarray(coo.data),
                             np.asarray(coo.todense()))
        r.close()
    return out


def sparse_frames(rng, nz, ny, nx, dtype, vmax, p=0.12):
    d = np.zeros((nz, ny, nx), dtype=np.int64)
    m = rng.random((nz, ny, nx)) < p
    d[m] = rng.integers(1, vmax + 1, size=int(m.sum()))
    return d.astype(dtype)


def check_oracle_vs_part(files, data, dark, eps, base):
    """the pinning step: oracle stage outputs == inflated reference streams"""
    thr = orc.make_threshold(dark, eps, dtype=data.dtype)
    level, b = base['reduction_level'], base['source_bit_depth']
    for node, path in enumerate(files):
        h, recs = orc.parse_part_file(path)
        off, cnt = orc.partition(data.shape[0], base['num_threads'], node)
        assert h['nz'] == cnt == len(recs), (h['nz'], cnt, len(recs))
        for i, rec in enumerate(recs):
            fid = off + i
            assert rec['frame_id'] == fid
            m, v, n = orc.reduce_frame(data[fid], thr, level, b)
            assert m == rec['map'], 'map mismatch'
            if level == 1:
                assert v == rec['vals'], 'packed values mismatch'
                assert rec['metadata']['bytes_in_packed_pixvals'] == len(v)


def make_converters_golden():
    """gold_e_converters.npz: the reference's recalibrate_l1 / l1_to_l4_converter (pyrecode/utils/converters.py)
    run live on seeded L1 frame dictionaries; the oracle restatement is asserted equal first."""
    from scipy.sparse import coo_matrix
    with contextlib.redirect_stdout(io.StringIO()):
        from pyrecode.utils.converters import recalibrate_l1, l1_to_l4_converter
    rng = np.random.default_rng(777)
    nz, ny, nx = 4, 80, 80
    # puddle-like L1 frames: event centres with random E / S / SE neighbours, 12-bit dark-subtracted values
    fr = np.zeros((nz, ny, nx), dtype=np.uint16)
    for z in range(nz):
        c = rng.random((ny, nx)) < 0.03
        v = np.where(c, rng.integers(50, 1000, (ny, nx)), 0)
        for dy, dx in ((0, 1), (1, 0), (1, 1)):
            m = np.roll(np.roll(c, dy, 0), dx, 1) & (rng.random((ny, nx)) < 0.5)
            v = np.where(m & (v == 0), rng.integers(25, 500, (ny, nx)), v)
        fr[z] = v
    fr[3] = 0                                                   # an empty frame
    frames = {10 + z: {'metadata': {'n': z}, 'data': coo_matrix(fr[z])} for z in range(nz)}
    orig = rng.integers(90, 110, (ny, nx)).astype(np.uint16)
    new = rng.integers(80, 125, (ny, nx)).astype(np.uint16)
    eps = 2.5
    with contextlib.redirect_stdout(io.StringIO()):
        rec = recalibrate_l1(frames, original_calibration_frame=orig, new_calibration_frame=new, epsilon=eps)
        l4 = l1_to_l4_converter(frames, (ny, nx))
    rec_d = np.stack([np.asarray(rec[k]['data'].todense()) for k in frames])
    l4_d = np.stack([np.asarray(l4[k]['data'].todense()) for k in frames])
    for i, k in enumerate(frames):
        assert np.array_equal(rec_d[i], orc.recalibrate_l1_frame(fr[i], orig, new, eps)), 'recalibrate oracle'
        assert np.array_equal(l4_d[i], orc.l1_to_l4_frame(fr[i], 0, True)), 'l1_to_l4 oracle'
    np.savez_compressed(os.path.join(GOLD, 'gold_e_converters.npz'), frames=fr, ids=np.array(list(frames)), orig=orig,
                        new=new, eps=eps, recalibrated=rec_d, l4=l4_d)
    return 'gold_e_converters.npz: recalibrate_l1 and l1_to_l4_converter (weighted_average) on %d frames of %dx%d' % (
        nz, ny, nx)


def make_calibration_golden():
    """gold_f_calibration.npz: the reference's make_calibration_frames (pyrecode/utils/calibration.py:87-138) and its
    numba _median_std_nb run live on a seeded dark stack.  The module imports pims (absent here) only to open the .seq
    file: a stand-in module whose open() returns the in-memory stack lets everything else run unmodified."""
    import types
    stack = []
    pims = types.ModuleType('pims')
    pims.open = lambda path: stack
    sys.modules['pims'] = pims
    with contextlib.redirect_stdout(io.StringIO()):
        from pyrecode.utils import calibration as cal
    rng = np.random.default_rng(4242)
    nF, ny, nx = 41, 40, 40                               # odd and (below) even frame counts
    dark = rng.integers(95, 110, (ny, nx))
    d = dark[None] + np.round(rng.normal(0, 3, (nF + 1, ny, nx)))
    ev = rng.random((nF + 1, ny, nx)) < 0.01
    d = np.where(ev, d + rng.integers(50, 1000, (nF + 1, ny, nx)), d)
    d[:, 3, 5] = rng.integers(0, 4000, nF + 1)            # a hot pixel: median outside any narrow window
    d[:, 7, 7] = 0
    d = d.astype(np.uint16)
    out = {}
    for tag, n in (('odd', nF), ('even', nF + 1)):
        m, s_ = cal._median_std_nb(d[:n], ny, nx)
        om, os_ = orc.median_std(d[:n])
        assert np.array_equal(m, om) and np.allclose(s_, os_, rtol=1e-6, atol=0), 'median/std oracle'
        out['med_' + tag], out['std_' + tag] = m, s_
    tmp = tempfile.mkdtemp(prefix='recode_cal_')
    stack.extend(list(d[:nF]))
    with contextlib.redirect_stdout(io.StringIO()):
        cal.make_calibration_frames('x.seq', np.uint16, nF, 10, 4, savepath=tmp, filename_prefix='c')
    thr = np.stack([np.fromfile(os.path.join(tmp, 'c__dark_ref_%d.bin' % i), dtype=np.uint16).reshape(ny, nx)
                    for i in range(4)])
    shutil.rmtree(tmp)
    # "accurate" per-pixel thresholds: a busier stack so that more than one event per pixel is expected
    ev2 = rng.random((nF, ny, nx)) < 0.08
    d2 = np.where(ev2, d[:nF].astype(np.int64) + rng.integers(50, 1000, (nF, ny, nx)), d[:nF]).astype(np.uint16)
    m2, _ = cal._median_std_nb(d2, ny, nx)
    for k in (2, 5):
        a = cal._get_pixel_thresh_2(d2, ny, nx, k, m2)
        with np.errstate(all='ignore'):
            assert np.array_equal(a, orc.pixel_thresholds(d2, m2, k), equal_nan=True), 'pixel thresholds oracle'
        out['acc_k%d' % k] = a
    del stack[:]
    stack.extend(list(d2))
    tmp = tempfile.mkdtemp(prefix='recode_cal_')
    with contextlib.redirect_stdout(io.StringIO()):
        cal.make_calibration_frames('x.seq', np.uint16, nF, 10, 4, savepath=tmp, filename_prefix='c', use_acc=True,
                                    sigma_acc=2)
    assert os.path.exists(os.path.join(tmp, 'c__dark_ref_2A.bin')), 'the busy stack should reach expected_n_events >= 2'
    out['acc_file'] = np.fromfile(os.path.join(tmp, 'c__dark_ref_2A.bin'), dtype=np.uint16).reshape(ny, nx)
    out['thresholds2'] = np.stack([np.fromfile(os.path.join(tmp, 'c__dark_ref_%d.bin' % i), dtype=np.uint16).reshape(ny, nx)
                                   for i in range(4)])
    shutil.rmtree(tmp)
    np.savez_compressed(os.path.join(GOLD, 'gold_f_calibration.npz'), stack=d, stack2=d2, n_odd=nF, thresholds=thr, **out)
    return 'gold_f_calibration.npz: _median_std_nb (odd / even frame counts) and make_calibration_frames thresholds'


def main():
    if len(sys.argv) > 1 and sys.argv[1] == 'calibration':
        print(make_calibration_golden())
        return
    if len(sys.argv) > 1 and sys.argv[1] == 'converters':      # add this fixture without regenerating the others
        print(make_converters_golden())
        return
    if os.path.isdir(GOLD):
        shutil.rmtree(GOLD)
    os.makedirs(GOLD)
    tmp = tempfile.mkdtemp(prefix='recode_golden_')
    rng = np.random.default_rng(20261018)
    report = []

    # ---- A: L1, 12 bit, 3 nodes, ragged pixel count, dark pattern + eps, merged file ---------
    nz, ny, nx = 8, 37, 53                         # 1961 px: nx*ny % 8 == 1; 8 frames over 3 nodes = 3,3,2
    dark = rng.integers(90, 110, size=(ny, nx)).astype(np.uint16)
    eps = 6
    data = dark[None].astype(np.int64) + rng.integers(-8, 8, size=(nz, ny, nx))
    ev = rng.random((nz, ny, nx)) < 0.07
    data[ev] += rng.integers(20, 3900, size=int(ev.sum()))
    data = np.clip(data, 0, 4095).astype(np.uint16)
    files, base = ref_write(tmp, 'gold_a', data, dark, 3, calibration_threshold_epsilon=eps)
    check_oracle_vs_part(files, data, dark, eps, base)
    with contextlib.redirect_stdout(io.StringIO()):
        merge_parts(tmp, 'gold_a.rc1', 3)
    merged = os.path.join(tmp, 'gold_a.rc1')
    dense_model = np.where(data > (dark + eps), data - (dark + eps), 0).astype(np.uint16)
    got = ref_read_all(merged, False)
    assert sorted(got) == list(range(nz))
    thr = orc.make_threshold(dark, eps)
    for z in range(nz):
        assert np.array_equal(got[z][3], dense_model[z])
        m, v, n = orc.reduce_frame(data[z], thr, 1, 12)
        tri = orc.unpack_sparse(ny, nx, 12, m, v, 1)
        assert np.array_equal(tri[:, 0], got[z][0]) and np.array_equal(tri[:, 1], got[z][1])
        assert np.array_equal(tri[:, 2], got[z][2])
        assert np.array_equal(orc.unpack_dense(ny, nx, 12, m, v, 1), dense_model[z])
    h, mrecs = orc.parse_merged_file(merged)
    assert h['nz'] == nz and h['is_intermediate'] == 1
    for f in files + [merged]:
        shutil.copy(f, GOLD)
    np.savez_compressed(os.path.join(GOLD, 'gold_a_input.npz'), data=data, dark=dark, eps=eps)
    report.append('A  L1 12-bit 8x37x53, 3 nodes + merged: oracle == reference (maps, packed values, triples, dense)')

    # ---- B: byte-aligned depths (tobytes shortcut, recode_writer.py:463-464) -----------------
    for b, dt in ((8, np.uint8), (16, np.uint16)):
        nz, ny, nx = 2, 24, 40
        dark_b = rng.integers(0, 4, size=(ny, nx)).astype(dt)
        data_b = sparse_frames(rng, nz, ny, nx, dt, (1 << b) - 1)
        name = 'gold_b%d' % b
        files, base = ref_write(tmp, name, data_b, dark_b, 1, calibration_threshold_epsilon=1,
                                source_bit_depth=b, target_bit_depth=b)
        if b == 16:
            check_oracle_vs_part(files, data_b, dark_b, 1, base)
        else:                                          # uint8 source: compare through widened arrays
            thr8 = orc.make_threshold(dark_b, 1, dtype=np.uint8)
            h, recs = orc.parse_part_file(files[0])
            for z, rec in enumerate(recs):
                m, v, n = orc.reduce_frame(data_b[z].astype(np.uint16), thr8.astype(np.uint16), 1, 8)
                assert m == rec['map'] and v == rec['vals']
        got = ref_read_all(files[0], True)
        thr_b = (dark_b + dt(1)).astype(dt)
        for z in range(nz):
            assert np.array_equal(got[z][3], np.where(data_b[z] > thr_b, data_b[z] - thr_b, 0))
        shutil.copy(files[0], GOLD)
        np.savez_compressed(os.path.join(GOLD, name + '_input.npz'), data=data_b, dark=dark_b, eps=1)
    report.append('B  L1 8-bit (uint8 source) and 16-bit part files: oracle == reference')

    # ---- C: L3 (map only), mode 1 and mode 0; L1 16-bit mode 0 -------------------------------
    nz, ny, nx = 3, 16, 48
    dark_c = np.full((ny, nx), 10, dtype=np.uint16)
    data_c = sparse_frames(rng, nz, ny, nx, np.uint16, 4095) + np.uint16(5)
    for mode in (1, 0):
        name = 'gold_c_l3m%d' % mode
        files, base = ref_write(tmp, name, data_c, dark_c, 1, reduction_level=3, rc_operation_mode=mode,
                                calibration_threshold_epsilon=2)
        check_oracle_vs_part(files, data_c, dark_c, 2, base)
        shutil.copy(files[0], GOLD)
    files, base = ref_write(tmp, 'gold_c_l1m0', data_c, dark_c, 1, reduction_level=1, rc_operation_mode=0,
                            calibration_threshold_epsilon=2, source_bit_depth=16, target_bit_depth=16)
    check_oracle_vs_part(files, data_c, dark_c, 2, base)
    shutil.copy(files[0], GOLD)
    np.savez_compressed(os.path.join(GOLD, 'gold_c_input.npz'), data=data_c, dark=dark_c, eps=2)
    report.append('C  L3 mode 1 / mode 0 and L1 16-bit mode 0 part files: oracle == reference')

    # ---- D: stage vectors ---------------------------------------------------------------------
    # D1 numba packers for every bit depth 1..16
    vals = rng.integers(0, 65536, size=77).astype(np.uint16)
    packs = {}
    for b in range(1, 17):
        ref = np.asarray(_bit_pack(vals, b))
        assert np.array_equal(ref, orc.bit_pack(vals, b)), b
        assert np.array_equal(orc.bit_unpack(ref, vals.size, b), vals & ((1 << b) - 1))
        packs['b%d' % b] = ref
    bm = rng.random((19, 23)) < 0.3
    ref_map = np.asarray(_pack_binary_frame(bm, (bm.size + 7) // 8))
    assert np.array_equal(ref_map, orc.pack_map(bm))
    assert np.array_equal(ref_map, np.packbits(bm.ravel(), bitorder='little'))
    np.savez_compressed(os.path.join(GOLD, 'gold_d_pack.npz'), vals=vals, bm=bm, ref_map=ref_map, **packs)

    # D2 labels (scipy, the reference's call) + weighted centroids (live numba function), including
    #    a tall frame with 16-bit values so float32 accumulation rounds (sum v*r > 2^24)
    cases = {}
    for tag, (ny2, nx2, p, vmax) in dict(small=(64, 96, 0.10, 4095), tall=(3000, 24, 0.22, 65535),
                                         dense=(40, 40, 0.45, 4095)).items():
        frame = np.zeros((ny2, nx2), dtype=np.uint16)
        m = rng.random((ny2, nx2)) < p
        frame[m] = rng.integers(1, vmax + 1, size=int(m.sum()))
        binary = frame > 0
        lab, k = nd.label(binary, structure=nd.generate_binary_structure(2, 2))
        olab, ok = orc.label8(binary)
        assert ok == k and np.array_equal(olab, lab), tag
        cen = np.asarray(get_centroids_2D_nb(lab, binary, frame, 0, 'weighted_average'), dtype=np.float64)
        ocen = orc.l4_centroids(lab, frame, k, 0)
        assert cen.shape == (k, 2)
        assert np.array_equal(cen.astype(np.float32), ocen) and np.array_equal(cen, ocen.astype(np.float64)), tag
        cases[tag + '_frame'] = frame
        cases[tag + '_labels'] = lab.astype(np.int32)
        cases[tag + '_centroids'] = cen.astype(np.float32)
    np.savez_compressed(os.path.join(GOLD, 'gold_d_ccl.npz'), **cases)

    # D3 c_recode.get_frame_sparse triples (L1 and the value-1 behaviour of other levels)
    ny3, nx3, b3 = 21, 35, 11
    bm3 = rng.random((ny3, nx3)) < 0.2
    v3 = rng.integers(0, 1 << b3, size=int(bm3.sum())).astype(np.uint16)
    map3 = np.packbits(bm3.ravel(), bitorder='little').tobytes()
    pk3 = orc.bit_pack(v3, b3).tobytes()
    rd = c_recode.Reader()
    rd.create_buffers(ny3, nx3, b3)
    tri = {}
    for level in (1, 3):
        buf = bytearray(ny3 * nx3 * 3 * 8)
        n = rd.get_frame_sparse(level, map3, pk3 + b'\0' * 8, buf)
        t = np.frombuffer(buf, dtype=np.uint64, count=n * 3).reshape(n, 3).copy()
        assert np.array_equal(t, orc.unpack_sparse(ny3, nx3, b3, map3, pk3, level))
        tri['triples_l%d' % level] = t
    np.savez_compressed(os.path.join(GOLD, 'gold_d_unpack.npz'), map=np.frombuffer(map3, np.uint8),
                        packed=np.frombuffer(pk3, np.uint8), ny=ny3, nx=nx3, b=b3, **tri)
    report.append('D  numba packers (b=1..16), scipy labels, live weighted centroids (incl. >2^24 sums), '
                  'c_recode triples: oracle == reference')

    shutil.rmtree(tmp)
    report.append(make_converters_golden())
    report.append(make_calibration_golden())
    with open(os.path.join(GOLD, 'README.md'), 'w') as f:
        f.write('# Golden fixtures\n\nGenerated by `python oracle/make_golden.py` in the build container from the '
                'unmodified reference at `/root/reference` (numpy %s, scipy %s). Every line below was asserted '
                'during generation:\n\n' % (np.__version__, __import__('scipy').__version__))
        for line in report:
            f.write('- ' + line + '\n')
        f.write('\nFiles named `gold_*.rc*` are byte-for-byte outputs of the reference writer / merge_parts.\n')
    for line in report:
        print(line)
    total = sum(os.path.getsize(os.path.join(GOLD, x)) for x in os.listdir(GOLD))
    print('fixtures: %d files, %d bytes' % (len(os.listdir(GOLD)), total))


if __name__ == '__main__':
    main()
