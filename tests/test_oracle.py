"""CPU: the oracle (oracle/) against the golden fixtures generated from the unmodified reference."""
import os

import numpy as np
import pytest

from oracle import oracle as orc


def _parts(gold_dir, stem, n):
    recs = []
    for node in range(n):
        recs += orc.parse_part_file(os.path.join(gold_dir, '%s_part%03d' % (stem, node)))[1]
    return recs


def test_l1_reference_streams(gold_dir):
    z = np.load(os.path.join(gold_dir, 'gold_a_input.npz'))
    data, dark, eps = z['data'], z['dark'], int(z['eps'])
    thr = orc.make_threshold(dark, eps)
    recs = _parts(gold_dir, 'gold_a.rc1', 3)
    assert [r['frame_id'] for r in recs] == list(range(data.shape[0]))
    for r in recs:
        m, v, n = orc.reduce_frame(data[r['frame_id']], thr, 1, 12)
        assert m == r['map'] and v == r['vals']
        assert r['metadata']['bytes_in_packed_pixvals'] == len(v) == (n * 12 + 7) // 8


def test_merged_file_layout(gold_dir):
    z = np.load(os.path.join(gold_dir, 'gold_a_input.npz'))
    h, recs = orc.parse_merged_file(os.path.join(gold_dir, 'gold_a.rc1'))
    parts = _parts(gold_dir, 'gold_a.rc1', 3)
    assert h['nz'] == z['data'].shape[0] == len(recs)
    for a, b in zip(recs, parts):
        assert a['map'] == b['map'] and a['vals'] == b['vals'] and a['metadata'] == b['metadata']


@pytest.mark.parametrize('b,dt', [(8, np.uint8), (16, np.uint16)])
def test_byte_aligned_depths(gold_dir, b, dt):
    z = np.load(os.path.join(gold_dir, 'gold_b%d_input.npz' % b))
    data, dark = z['data'], z['dark']
    thr = orc.make_threshold(dark, int(z['eps']), dtype=dt)
    h, recs = orc.parse_part_file(os.path.join(gold_dir, 'gold_b%d.rc1_part000' % b))
    assert h['source_bit_depth'] == b
    for r in recs:
        m, v, n = orc.reduce_frame(data[r['frame_id']].astype(np.uint16), thr.astype(np.uint16), 1, b)
        assert m == r['map'] and v == r['vals']
        dense = orc.unpack_dense(h['ny'], h['nx'], b, m, v, 1)
        assert np.array_equal(dense, np.where(data[r['frame_id']] > thr, data[r['frame_id']] - thr, 0))


@pytest.mark.parametrize('name,level,mode,b', [('gold_c_l3m1.rc3', 3, 1, 12), ('gold_c_l3m0.rc3', 3, 0, 12),
                                               ('gold_c_l1m0.rc1', 1, 0, 16)])
def test_l3_and_reduce_only(gold_dir, name, level, mode, b):
    z = np.load(os.path.join(gold_dir, 'gold_c_input.npz'))
    thr = orc.make_threshold(z['dark'], int(z['eps']))
    h, recs = orc.parse_part_file(os.path.join(gold_dir, name + '_part000'))
    assert (h['reduction_level'], h['rc_operation_mode']) == (level, mode)
    for r in recs:
        m, v, n = orc.reduce_frame(z['data'][r['frame_id']], thr, level, b)
        assert m == r['map']
        if level == 1:
            assert v == r['vals']
    # build_record restates the record layout: mode 0 records are reproduced byte for byte
    if mode == 0:
        raw = open(os.path.join(gold_dir, name + '_part000'), 'rb').read()[orc.HEADER_LEN:]
        mine = b''.join(orc.build_record(r['frame_id'], level, 0, r['map'], r['vals'] or b'') for r in recs)
        assert mine == raw


def test_packers(gold_dir):
    z = np.load(os.path.join(gold_dir, 'gold_d_pack.npz'))
    for b in range(1, 17):
        assert np.array_equal(orc.bit_pack(z['vals'], b), z['b%d' % b])
        assert np.array_equal(orc.bit_unpack(z['b%d' % b], z['vals'].size, b), z['vals'] & ((1 << b) - 1))
    assert np.array_equal(orc.pack_map(z['bm']), z['ref_map'])


def test_labels_and_centroids(gold_dir):
    z = np.load(os.path.join(gold_dir, 'gold_d_ccl.npz'))
    for tag in ('small', 'tall', 'dense'):
        frame = z[tag + '_frame']
        lab, k = orc.label8(frame > 0)
        assert np.array_equal(lab, z[tag + '_labels'])
        cen = orc.l4_centroids(lab, frame, k, 0)
        assert np.array_equal(cen.view(np.uint32), z[tag + '_centroids'].view(np.uint32))


def test_labels_against_scipy():
    import scipy.ndimage as nd
    rng = np.random.default_rng(3)
    s = nd.generate_binary_structure(2, 2)
    for occ in (0.02, 0.3, 0.55, 0.9):
        b = rng.random((97, 131)) < occ
        lab, k = nd.label(b, structure=s)
        olab, ok = orc.label8(b)
        assert ok == k and np.array_equal(olab, lab)


def test_unpack_triples(gold_dir):
    z = np.load(os.path.join(gold_dir, 'gold_d_unpack.npz'))
    for level in (1, 3):
        t = orc.unpack_sparse(int(z['ny']), int(z['nx']), int(z['b']), z['map'].tobytes(), z['packed'].tobytes(), level)
        assert np.array_equal(t, z['triples_l%d' % level])


def test_hand_checked_l2_l4():
    # 5 x 6 frame, two puddles: {(0,0),(1,1)} values 10, 30 and {(3,4),(3,5),(4,4)} values 7, 9, 200
    f = np.zeros((5, 6), np.uint16)
    f[0, 0], f[1, 1], f[3, 4], f[3, 5], f[4, 4] = 10, 30, 7, 9, 200
    thr = np.zeros((5, 6), np.uint16)
    lab, k = orc.label8(f > 0)
    assert k == 2 and lab[0, 0] == lab[1, 1] == 1 and lab[3, 4] == lab[3, 5] == lab[4, 4] == 2
    assert list(orc.l2_stats(lab, f, k, 0)) == [30, 200]
    assert list(orc.l2_stats(lab, f, k, 2)) == [40, 216]
    c = orc.l4_centroids(lab, f, k, 0)
    assert np.allclose(c[0], [30 / 40, 30 / 40]) and np.allclose(c[1], [(7 * 3 + 9 * 3 + 200 * 4) / 216, (7 * 4 + 9 * 5 + 200 * 4) / 216])
    m = orc.centroid_map(c, 5, 6)
    assert m.sum() == 2 and m[1, 1] == 1 and m[4, 4] == 1
    assert np.array_equal(orc.l4_centroids(lab, f, k, 2), np.array([[1, 1], [4, 4]], np.float32))
    assert np.allclose(orc.l4_centroids(lab, f, k, 3), [[0.5, 0.5], [10 / 3, 13 / 3]])
    m1, v1, n1 = orc.reduce_frame(f, thr, 2, 12, l2_statistics=2)
    assert n1 == 2 and v1 == orc.bit_pack(np.array([40, 216], np.uint16), 12).tobytes()


def test_partition_rule():
    # recode_writer.py:320-322
    assert [orc.partition(9, 3, i) for i in range(3)] == [(0, 3), (3, 3), (6, 3)]
    assert [orc.partition(8, 3, i) for i in range(3)] == [(0, 3), (3, 3), (6, 2)]
    assert [orc.partition(2, 4, i) for i in range(4)] == [(0, 1), (1, 1), (2, 0), (3, 0)]


def test_converters_against_reference(gold_dir):
    """oracle restatement of recalibrate_l1 / l1_to_l4_converter (pyrecode/utils/converters.py:15-123) against
    the outputs the live reference produced (oracle/make_golden.py converters)"""
    z = np.load(os.path.join(gold_dir, 'gold_e_converters.npz'))
    fr = z['frames']
    for i in range(fr.shape[0]):
        assert np.array_equal(orc.recalibrate_l1_frame(fr[i], z['orig'], z['new'], float(z['eps'])), z['recalibrated'][i])
        assert np.array_equal(orc.l1_to_l4_frame(fr[i], 0, True), z['l4'][i])
    assert not z['l4'][3].any() and z['l4'][:3].any()


def test_calibration_median_std_against_reference(gold_dir):
    """oracle restatement of _median_std_nb (pyrecode/utils/calibration.py:48-57) against the live reference's output"""
    z = np.load(os.path.join(gold_dir, 'gold_f_calibration.npz'))
    n = int(z['n_odd'])
    for tag, k in (('odd', n), ('even', n + 1)):
        m, s = orc.median_std(z['stack'][:k])
        assert np.array_equal(m, z['med_' + tag])
        assert np.allclose(s, z['std_' + tag], rtol=1e-6, atol=0)
    assert z['thresholds'].shape[0] == 4 and np.array_equal(z['thresholds'][0], np.floor(z['med_odd']).astype(np.uint16))
    m2 = orc.median_std(z['stack2'])[0]
    with np.errstate(all='ignore'):
        for k in (2, 5):
            assert np.array_equal(orc.pixel_thresholds(z['stack2'], m2, k), z['acc_k%d' % k])
