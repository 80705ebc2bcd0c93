"""GPU: the drop-in Python API end to end -- ReCoDeWriter -> part files -> ReCoDeReader / merge_parts, the way the
reference's own test drives it (tests/minimal_read_write_test.py:15-124), for every reduction level and both
operation modes, checked against the oracle and against the golden files the reference wrote."""
import os

import numpy as np
import pytest

from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def make_params(ny, nx, nz, level=1, mode=1, b=12, threads=1, eps=0, l2=0, l4=0, clevel=1):
    from pyrecode_b200.params import InputParams
    ip = InputParams()
    vals = dict(l4_centroiding=l4, source_file_type=0, num_frames=nz, source_header_length=0,
                calibration_frame_offset=0, compression_scheme=0, calibration_file_type=0, compression_level=clevel,
                l2_statistics=l2, calibration_threshold_epsilon=eps, frame_offset=0, num_threads=threads,
                rc_operation_mode=mode, num_calibration_frames=1, reduction_level=level, keep_calibration_data=1,
                source_bit_depth=b, target_bit_depth=b, keep_part_files=0, num_rows=ny, num_cols=nx,
                source_data_type=0, target_data_type=0)
    for k, v in vals.items():
        ip._param_map[k] = v
    assert ip.validate()
    return ip


def reference_test_data(rng, nz=9, ny=512, nx=512):
    """the data model of tests/minimal_read_write_test.py:15-25 (seeded here)"""
    d = rng.integers(0, 4096, size=(nz, ny, nx)).astype(np.int64) - 3500
    d[d < 0] = 0
    return d.astype(np.uint16)


def write_parts(tmp, name, data, dark, ip, n_nodes, **kw):
    from pyrecode_b200.recode_writer import ReCoDeWriter
    metrics = []
    for node in range(n_nodes):
        w = ReCoDeWriter(name, dark_data=dark, output_directory=str(tmp), input_params=ip, mode='batch', node_id=node,
                         **kw)
        w.start()
        metrics.append(w.run(data))
        w.close()
    return metrics


def test_minimal_read_write(tmp_path):
    """tests/minimal_read_write_test.py: L1 / zlib 1 / 12 bit / 3 workers; part file, merged file, random access"""
    from pyrecode_b200.recode_reader import ReCoDeReader, merge_parts
    rng = np.random.default_rng(42)
    data = reference_test_data(rng)
    dark = np.zeros((1, 512, 512), np.uint16)
    ip = make_params(512, 512, 9, threads=3)
    metrics = write_parts(tmp_path, 'test_data', data, dark, ip, 3)
    assert [m['run_frames'] for m in metrics] == [3, 3, 3]
    assert 'frame_binary_image_compression_time' in metrics[0] and 'run_time' in metrics[0]

    r = ReCoDeReader(str(tmp_path / 'test_data.rc1_part000'), is_intermediate=True)
    r.open(print_header=False)
    for i in range(3):
        fd = r.get_next_frame()
        (fid, fr), = fd.items()
        assert fid == i
        assert np.array_equal(np.asarray(fr['data'].todense()), data[fid])
        assert fr['data'].dtype == np.uint16
    assert r.get_next_frame() is None
    r.close()

    merge_parts(str(tmp_path), 'test_data.rc1', 3)
    r = ReCoDeReader(str(tmp_path / 'test_data.rc1'), is_intermediate=False)
    r.open(print_header=False)
    assert r.get_shape() == (9, 512, 512)
    for i in range(9):
        (fid, fr), = r.get_next_frame().items()
        assert fid == i and np.array_equal(np.asarray(fr['data'].todense()), data[i])
    for z in (7, 0, 4):
        (fid, fr), = r.get_frame(z).items()
        assert fid == z and np.array_equal(fr['data'].toarray(), data[z])
    with pytest.raises(ValueError):
        r.get_frame(9)
    r.close()


@pytest.mark.parametrize('level', [1, 2, 3, 4])
@pytest.mark.parametrize('mode', [1, 0])
def test_levels_and_modes_roundtrip(tmp_path, level, mode):
    """every (level, mode): record streams == oracle, and the reader returns what the format promises"""
    from pyrecode_b200.recode_reader import ReCoDeReader
    rng = np.random.default_rng(level * 10 + mode)
    nz, ny, nx, b, eps = 5, 96, 160, 12, 3
    dark = rng.integers(0, 6, size=(ny, nx)).astype(np.uint16)
    frames = orc.synth_frames('l2' if level != 1 else 'l1', nz, ny, nx, dark + 20, seed=5 + level, bit_depth=b)
    ip = make_params(ny, nx, nz, level=level, mode=mode, b=b, eps=eps)
    write_parts(tmp_path, 'x', frames, dark[None], ip, 1)
    path = str(tmp_path / ('x.rc%d_part000' % level))
    thr = orc.make_threshold(dark, eps)
    hdr, recs = orc.parse_part_file(path)
    assert hdr['nz'] == nz and hdr['reduction_level'] == level and hdr['rc_operation_mode'] == mode
    r = ReCoDeReader(path, is_intermediate=True)
    r.open(print_header=False)
    for f in range(nz):
        m_ref, v_ref, n_ref = orc.reduce_frame(frames[f], thr, level, b)
        rec = recs[f]
        assert rec['frame_id'] == f
        assert rec['map'] == m_ref                          # parse_part_file inflates with stock zlib
        if level <= 2:
            assert rec['vals'] == v_ref
            assert rec['metadata'][orc.metadata_fields(level, mode)[-1]] == len(v_ref)
        (fid, fr), = r.get_next_frame().items()
        assert fid == f
        dense = fr['data'].toarray()
        binary = np.unpackbits(np.frombuffer(m_ref, np.uint8), bitorder='little')[:ny * nx].reshape(ny, nx)
        if level == 1:
            want = np.where(frames[f] > thr, frames[f] - thr, 0)
            assert np.array_equal(dense, want)
        else:
            assert np.array_equal(dense, binary)            # value 1 at every map pixel (reader.h:39-41)
        if level == 2:
            assert np.array_equal(fr['summary_stats'], orc.bit_unpack(np.frombuffer(v_ref, np.uint8), n_ref, b))
    assert r.get_next_frame() is None
    r.close()


def test_multi_batch_run_matches_single_frames(tmp_path):
    """more frames than one launch batch, odd geometry, two batches in flight: records in order and identical
    payloads to frame-at-a-time processing"""
    from pyrecode_b200.recode_writer import ReCoDeWriter
    rng = np.random.default_rng(3)
    nz, ny, nx = 23, 37, 53
    dark = rng.integers(0, 4, size=(ny, nx)).astype(np.uint16)
    frames = (dark[None] + (rng.random((nz, ny, nx)) < 0.1) * rng.integers(1, 4000, size=(nz, ny, nx))).astype(np.uint16)
    ip = make_params(ny, nx, nz, level=2, eps=1)
    for tag, bf in (('a', 4), ('b', 1)):
        w = ReCoDeWriter(tag, dark_data=dark[None], output_directory=str(tmp_path), input_params=ip, batch_frames=bf)
        w.start()
        w.run(frames)
        w.close()
    ha, ra = orc.parse_part_file(str(tmp_path / 'a.rc2_part000'))
    hb, rb = orc.parse_part_file(str(tmp_path / 'b.rc2_part000'))
    assert ha['nz'] == hb['nz'] == nz
    for f in range(nz):
        assert ra[f]['frame_id'] == rb[f]['frame_id'] == f
        assert ra[f]['map'] == rb[f]['map'] and ra[f]['vals'] == rb[f]['vals']


def test_cuda_tensor_input_and_stream_chunks(tmp_path):
    """run() accepts a CUDA tensor (no host round trip) and successive run() calls continue the frame ids
    (stream mode, recode_writer.py:385)"""
    import torch
    from pyrecode_b200.recode_reader import ReCoDeReader
    from pyrecode_b200.recode_writer import ReCoDeWriter
    rng = np.random.default_rng(8)
    ny, nx = 64, 96
    dark = np.zeros((1, ny, nx), np.uint16)
    chunks = [reference_test_data(rng, 4, ny, nx) for _ in range(3)]
    ip = make_params(ny, nx, 4)
    w = ReCoDeWriter('', dark_data=dark, output_directory=str(tmp_path), input_params=ip, mode='stream', run_name='live')
    w.start()
    for c in chunks:
        w.run(torch.from_numpy(c).cuda())
    w.close()
    r = ReCoDeReader(str(tmp_path / 'live.rc1_part000'), is_intermediate=True)
    r.open(print_header=False)
    assert r.get_shape()[0] == 12
    allf = np.concatenate(chunks)
    for i in range(12):
        (fid, fr), = r.get_next_frame().items()
        assert fid == i and np.array_equal(fr['data'].toarray(), allf[i])
    r.close()


def test_reads_reference_written_files(gold_dir):
    """files written by the reference writer / merge_parts (tests/golden) decode to the reference's input"""
    from pyrecode_b200.recode_reader import ReCoDeReader
    z = np.load(os.path.join(gold_dir, 'gold_a_input.npz'))
    data, dark, eps = z['data'], z['dark'], int(z['eps'])
    thr = orc.make_threshold(dark, eps)
    r = ReCoDeReader(os.path.join(gold_dir, 'gold_a.rc1'))
    r.open(print_header=False)
    nz = r.get_shape()[0]
    for i in list(range(nz)) + [nz - 1, 0]:
        fd = r.get_frame(i) if i in (nz - 1, 0) else r.get_next_frame()
        (fid, fr), = fd.items()
        want = np.where(data[fid] > thr, data[fid] - thr, 0)
        assert np.array_equal(fr['data'].toarray(), want)
    r.close()


def test_dense_and_live_view_sum(tmp_path):
    """batched read extras: dense frames on the device and the summed live-view image
    (examples/ReCoDe_Live_View_MT.ipynb cell 1) equal the sum of the reconstructed frames"""
    from pyrecode_b200.recode_reader import ReCoDeReader
    rng = np.random.default_rng(11)
    nz, ny, nx = 20, 128, 256
    data = reference_test_data(rng, nz, ny, nx)
    ip = make_params(ny, nx, nz)
    write_parts(tmp_path, 'lv', data, np.zeros((1, ny, nx), np.uint16), ip, 1)
    r = ReCoDeReader(str(tmp_path / 'lv.rc1_part000'), is_intermediate=True, batch_frames=8)
    r.open(print_header=False)
    ids, dense = r.read_frames_dense(7)
    assert ids == list(range(7)) and np.array_equal(dense.cpu().numpy(), data[:7])
    ids, total = r.sum_frames(100)
    assert ids == list(range(7, nz))
    assert np.array_equal(total.cpu().numpy().astype(np.int64).reshape(ny, nx), data[7:].astype(np.int64).sum(0))
    r.close()


@pytest.mark.parametrize('level', [1, 2, 3])
def test_bulk_read_pipeline(tmp_path, level):
    """read_frames_dense / sum_frames with several small batches in flight (records staged through the engines'
    pinned blocks), part file and merged file, equal the frame-by-frame reader"""
    from pyrecode_b200.recode_reader import ReCoDeReader, merge_parts
    rng = np.random.default_rng(5 + level)
    nz, ny, nx = 23, 96, 160
    data = reference_test_data(rng, nz, ny, nx)
    ip = make_params(ny, nx, nz, level=level, threads=2)
    write_parts(tmp_path, 'bulk', data, np.zeros((1, ny, nx), np.uint16), ip, 2)
    merge_parts(str(tmp_path), 'bulk.rc%d' % level, 2)
    r = ReCoDeReader(str(tmp_path / ('bulk.rc%d' % level)))
    r.open(print_header=False)
    want = np.stack([r.get_next_frame().popitem()[1]['data'].toarray() for _ in range(nz)])
    r.close()
    for name, inter, n_expect in (('bulk.rc%d' % level, False, nz), ('bulk.rc%d_part001' % level, True, nz - 12)):
        r = ReCoDeReader(str(tmp_path / name), is_intermediate=inter, bulk_frames=3)
        r.open(print_header=False)
        ids, dense = r.read_frames_dense(8)
        first = 12 if inter else 0
        assert ids == list(range(first, first + 8))
        assert np.array_equal(dense.cpu().numpy(), want[first:first + 8])
        ids2, total = r.sum_frames(1000)
        assert ids2 == list(range(first + 8, first + n_expect))
        assert np.array_equal(total.cpu().numpy().astype(np.int64).reshape(ny, nx),
                              want[first + 8:first + n_expect].astype(np.int64).sum(0))
        r.rewind()                                     # the same engines again, from the first frame
        ids3, total3 = r.sum_frames(5)
        assert ids3 == list(range(first, first + 5))
        assert np.array_equal(total3.cpu().numpy().astype(np.int64).reshape(ny, nx),
                              want[first:first + 5].astype(np.int64).sum(0))
        # frame-by-frame reads continue where the bulk call stopped
        (fid, fr), = r.get_next_frame().items()
        assert fid == first + 5 and np.array_equal(fr['data'].toarray(), want[first + 5])
        r.close()


def test_converters_golden(gold_dir):
    """utils.converters.recalibrate_l1 / l1_to_l4_converter on the GPU == the live reference's outputs"""
    from scipy.sparse import coo_matrix
    from pyrecode_b200.utils.converters import recalibrate_l1, l1_to_l4_converter
    z = np.load(os.path.join(gold_dir, 'gold_e_converters.npz'))
    fr, ids = z['frames'], [int(i) for i in z['ids']]
    frames = {k: {'metadata': {'n': i}, 'data': coo_matrix(fr[i])} for i, k in enumerate(ids)}
    rec = recalibrate_l1(frames, original_calibration_frame=z['orig'], new_calibration_frame=z['new'],
                         epsilon=float(z['eps']), batch_frames=3)
    l4 = l1_to_l4_converter(frames, fr.shape[1:], batch_frames=3)
    assert list(rec) == ids and list(l4) == ids
    for i, k in enumerate(ids):
        assert rec[k]['metadata'] == {'n': i} and rec[k]['data'].dtype == np.uint16
        assert np.array_equal(np.asarray(rec[k]['data'].todense()), z['recalibrated'][i])
        assert l4[k]['data'].dtype == bool and l4[k]['data'].shape == fr.shape[1:]
        assert np.array_equal(np.asarray(l4[k]['data'].todense()), z['l4'][i])
    assert frames[ids[0]]['data'].dtype == np.uint16                # inputs untouched (in_place=False)


@pytest.mark.parametrize('method,mode', [('weighted_average', 0), ('max', 2), ('unweighted', 3)])
def test_converters_random(method, mode):
    """larger, non-square frames (transpose=False) and every centroiding method against the oracle; the device-
    resident recalibration; uint8 frames"""
    import torch
    from scipy.sparse import coo_matrix
    from pyrecode_b200.utils.converters import recalibrate_l1, recalibrate_dense, l1_to_l4_converter
    rng = np.random.default_rng(99 + mode)
    nz, ny, nx = 5, 200, 328
    fr = np.where(rng.random((nz, ny, nx)) < 0.08, rng.integers(1, 4096, (nz, ny, nx)), 0).astype(np.uint16)
    frames = {i: {'data': coo_matrix(fr[i])} for i in range(nz)}
    l4 = l1_to_l4_converter(frames, (ny, nx), method=method, transpose=False)
    for i in range(nz):
        assert np.array_equal(np.asarray(l4[i]['data'].todense()), orc.l1_to_l4_frame(fr[i], mode, False))
    with pytest.raises(NotImplementedError):
        l1_to_l4_converter(frames, (ny, nx), area_threshold=2)
    orig = rng.integers(0, 300, (ny, nx)).astype(np.uint16)
    new = rng.integers(0, 300, (ny, nx)).astype(np.uint16)
    out = recalibrate_dense(torch.from_numpy(fr).cuda(), orig, new, epsilon=-0.75)
    want = np.stack([orc.recalibrate_l1_frame(fr[i], orig, new, -0.75) for i in range(nz)])
    assert np.array_equal(out.cpu().numpy(), want)
    # saturation at both ends, uint8
    f8 = rng.integers(0, 256, (2, 64, 64)).astype(np.uint8)
    o8 = rng.integers(0, 256, (64, 64)).astype(np.uint8)
    n8 = rng.integers(0, 256, (64, 64)).astype(np.uint8)
    r8 = recalibrate_l1({i: {'data': coo_matrix(f8[i])} for i in range(2)}, original_calibration_frame=o8,
                        new_calibration_frame=n8, epsilon=0.5)
    for i in range(2):
        assert np.array_equal(np.asarray(r8[i]['data'].todense()), orc.recalibrate_l1_frame(f8[i], o8, n8, 0.5))


@pytest.mark.parametrize('level', [1, 2, 4])
def test_pipelined_contexts_match_single(level):
    """three batches in flight (rc_set_pipelined: post-streaming work on the contexts' high-priority streams, persistent
    labelling grid) produce byte-identical records to one context running one batch at a time"""
    from pyrecode_b200.engine import WriteEngine
    ny, nx, F = 256, 512, 6
    dark = orc.synth_dark(ny, nx)
    frames = np.stack(orc.synth_frames({1: 'l1', 2: 'l2', 4: 'l4'}[level], 3 * F, ny, nx, dark, seed=77))
    piped = WriteEngine(ny, nx, 2, 12, level, 1, 0, 0, 1, max_frames=F, n_slots=3)
    piped.set_threshold(dark, 20)
    want = []
    for b in range(3):
        # a fresh context per batch, like the slot that will see this batch: the Huffman code a context keeps across
        # calls is built from the first batch it sees
        single = WriteEngine(ny, nx, 2, 12, level, 1, 0, 0, 1, max_frames=F, n_slots=1)
        single.set_threshold(dark, 20)
        rec, offs, counts, _, _ = single.reduce_compress(frames[b * F:(b + 1) * F], first_frame_id=b * F)
        want.append((bytes(rec[:int(offs[F])]), offs.copy(), counts.copy()))
        del single
    for rep in range(3):                               # slots are reused: the joins between the streams must hold
        ks = [piped.submit(frames[b * F:(b + 1) * F], first_frame_id=b * F) for b in range(3)]
        for b, k in enumerate(ks):
            rec, offs, counts, _, _ = piped.collect(k)
            assert np.array_equal(offs, want[b][1]) and np.array_equal(counts, want[b][2])
            assert bytes(rec[:int(offs[F])]) == want[b][0]


def test_read_ahead_is_invisible(tmp_path):
    """get_next_frame decodes batch_frames frames per GPU round trip; frame order, the observable file position,
    get_next_frame_raw / get_frame in between and the end of file behave as with one frame at a time"""
    from pyrecode_b200.recode_reader import ReCoDeReader, merge_parts
    rng = np.random.default_rng(21)
    nz, ny, nx = 11, 64, 96
    data = reference_test_data(rng, nz, ny, nx)
    ip = make_params(ny, nx, nz, threads=1)
    write_parts(tmp_path, 'ra', data, np.zeros((1, ny, nx), np.uint16), ip, 1)
    merge_parts(str(tmp_path), 'ra.rc1', 1)
    for name, inter in (('ra.rc1_part000', True), ('ra.rc1', False)):
        one = ReCoDeReader(str(tmp_path / name), is_intermediate=inter, batch_frames=1)
        many = ReCoDeReader(str(tmp_path / name), is_intermediate=inter, batch_frames=4)
        one.open(print_header=False)
        many.open(print_header=False)
        for i in range(nz):
            if i == 5:                                  # a raw read in the middle of a read-ahead batch
                a, b = one.get_next_frame_raw(), many.get_next_frame_raw()
                assert list(a) == list(b) == [5] and a[5]['data']['binary_map'] == b[5]['data']['binary_map']
                continue
            a, b = one.get_next_frame(), many.get_next_frame()
            assert list(a) == list(b) == [i]
            assert np.array_equal(a[i]['data'].toarray(), data[i]) and np.array_equal(b[i]['data'].toarray(), data[i])
            assert one.get_file_position() == many.get_file_position()
        assert one.get_next_frame() is None and many.get_next_frame() is None      # end of file (recode_reader.py:229-230)
        if not inter:
            fr = many.get_frame(3)
            assert np.array_equal(fr[3]['data'].toarray(), data[3])
            assert list(many.get_next_frame()) == [4]
        one.close()
        many.close()


@pytest.mark.parametrize('level,mode', [(1, 1), (2, 1), (4, 1), (1, 0), (3, 0)])
def test_merged_writer_equals_merge_parts(tmp_path, level, mode):
    """ReCoDeWriter(merged=True) emits byte for byte the file merge_parts makes from the part file of the same run
    (header, metadata table, payloads: pyrecode/recode_reader.py:513-591)"""
    from pyrecode_b200.recode_reader import ReCoDeReader, merge_parts
    from pyrecode_b200.recode_writer import ReCoDeWriter
    rng = np.random.default_rng(40 + level)
    nz, ny, nx = 13, 72, 104
    data = reference_test_data(rng, nz, ny, nx)
    dark = np.zeros((1, ny, nx), np.uint16)
    ip = make_params(ny, nx, nz, level=level, mode=mode)
    (tmp_path / 'a').mkdir()
    (tmp_path / 'b').mkdir()
    write_parts(tmp_path / 'a', 'm', data, dark, ip, 1, batch_frames=4)
    merge_parts(str(tmp_path / 'a'), 'm.rc%d' % level, 1)
    w = ReCoDeWriter('m', dark_data=dark, output_directory=str(tmp_path / 'b'), input_params=ip, mode='batch',
                     batch_frames=4, merged=True)
    w.start()
    w.run(data)
    w.close()
    a = (tmp_path / 'a' / ('m.rc%d' % level)).read_bytes()
    b = (tmp_path / 'b' / ('m.rc%d' % level)).read_bytes()
    assert a == b
    r = ReCoDeReader(str(tmp_path / 'b' / ('m.rc%d' % level)))
    r.open(print_header=False)
    assert r.get_true_shape()[0] == nz
    fr = r.get_frame(nz - 2)
    if level == 1:
        assert np.array_equal(fr[nz - 2]['data'].toarray(), data[nz - 2])
    r.close()
    with pytest.raises(ValueError):
        ReCoDeWriter('m', dark_data=dark, output_directory=str(tmp_path / 'b'), input_params=ip, mode='batch', node_id=1,
                     merged=True)


def test_calibration_golden(gold_dir, tmp_path):
    """utils.calibration on the GPU == the live reference: per-pixel median (exact) and std (float32, 1e-6 relative:
    the reference sums float64 deviations, the kernel exact integers), and the threshold frames
    make_calibration_frames writes (pyrecode/utils/calibration.py:87-138)"""
    from pyrecode_b200.utils.calibration import median_std, make_calibration_frames
    z = np.load(os.path.join(gold_dir, 'gold_f_calibration.npz'))
    n = int(z['n_odd'])
    for tag, k in (('odd', n), ('even', n + 1)):
        m, s = median_std(z['stack'][:k])
        assert m.dtype == np.float32 and np.array_equal(m, z['med_' + tag])
        assert np.allclose(s, z['std_' + tag], rtol=1e-6, atol=0)
    res = make_calibration_frames('unused.seq', np.uint16, n, 10, 4, savepath=str(tmp_path), filename_prefix='c',
                                  data=z['stack'][:n])
    for i in range(4):
        t = np.fromfile(str(tmp_path / ('c__dark_ref_%d.bin' % i)), dtype=np.uint16).reshape(z['thresholds'][i].shape)
        assert np.array_equal(t, z['thresholds'][i]), i
        assert np.array_equal(res['thresholds'][i], t)
    # the per-pixel "accurate" thresholds: as the live reference computes them (as_run) and as written (oracle)
    from pyrecode_b200.utils.calibration import pixel_thresholds
    d2 = z['stack2']
    m2, _ = median_std(d2)
    for k in (2, 5):
        a = pixel_thresholds(d2, m2, k)
        assert a.dtype == np.float32 and np.array_equal(a, z['acc_k%d' % k])
        with np.errstate(all='ignore'):
            assert np.array_equal(pixel_thresholds(d2, m2, k, as_run=False), orc.pixel_thresholds(d2, m2, k, as_run=False))
    make_calibration_frames('unused.seq', np.uint16, n, 10, 4, savepath=str(tmp_path), filename_prefix='d', data=d2,
                            use_acc=True, sigma_acc=2)
    acc = np.fromfile(str(tmp_path / 'd__dark_ref_2A.bin'), dtype=np.uint16).reshape(z['acc_file'].shape)
    assert np.array_equal(acc, z['acc_file'])
    for i in range(4):
        t = np.fromfile(str(tmp_path / ('d__dark_ref_%d.bin' % i)), dtype=np.uint16).reshape(z['acc_file'].shape)
        assert np.array_equal(t, z['thresholds2'][i]), i


@pytest.mark.parametrize('dtype,n', [(np.uint16, 64), (np.uint16, 257), (np.uint8, 33)])
def test_calibration_median_std_random(dtype, n):
    """wider distributions (most medians outside the 64-count window: the bisection kernel), uint8 stacks, ragged
    pixel counts, against the oracle"""
    from pyrecode_b200.utils.calibration import median_std
    rng = np.random.default_rng(n)
    ny, nx = 37, 53
    hi = 250 if dtype == np.uint8 else 60000
    stack = rng.integers(0, hi, (n, ny, nx)).astype(dtype)
    stack[:, :10] = (100 + rng.integers(0, 6, (n, 10, nx))).astype(dtype)      # narrow rows: the window path
    m, s = median_std(stack)
    om, os_ = orc.median_std(stack)
    assert np.array_equal(m, om)
    assert np.allclose(s, os_, rtol=1e-6, atol=0)


def test_c_recode_shim(gold_dir):
    """c_recode.Reader signatures (pyrecode.cpp:57-141) on the GPU, against the reference's recorded triples"""
    from pyrecode_b200 import c_recode
    z = np.load(os.path.join(gold_dir, 'gold_d_unpack.npz'))
    ny, nx, b = int(z['ny']), int(z['nx']), int(z['b'])
    rd = c_recode.Reader()
    assert rd.create_buffers(ny, nx, b) == 1
    out = np.zeros(ny * nx * 3, dtype=np.uint64)
    n = rd.get_frame_sparse(1, z['map'].tobytes(), z['packed'].tobytes(), out)
    assert n == z['triples_l1'].shape[0] and np.array_equal(out[:n * 3].reshape(n, 3), z['triples_l1'])
    vals = z['triples_l1'][:, 2].astype(np.uint16)
    packed = np.zeros((n * b + 7) // 8 + 8, dtype=np.uint8)
    rd.bit_pack_pixel_intensities(len(packed), n, b, vals, packed)
    assert packed[:(n * b + 7) // 8].tobytes() == z['packed'].tobytes()[:(n * b + 7) // 8]
    back = np.zeros(n, dtype=np.uint64)
    assert rd.bit_unpack_pixel_intensities(n, packed, back) == n
    assert np.array_equal(back, vals)


def test_writer_errors(tmp_path):
    from pyrecode_b200.recode_writer import ReCoDeWriter
    dark = np.zeros((1, 32, 32), np.uint16)
    with pytest.raises(RuntimeError):                       # calibration frame shape mismatch (recode_writer.py:123-124)
        ReCoDeWriter('x', dark_data=np.zeros((1, 16, 32), np.uint16), output_directory=str(tmp_path),
                     input_params=make_params(32, 32, 2))
    w = ReCoDeWriter('x', dark_data=dark, output_directory=str(tmp_path), input_params=make_params(32, 32, 2))
    w.start()
    with pytest.raises(RuntimeError):                       # frame shape mismatch (recode_writer.py:274-278)
        w.run(np.zeros((2, 32, 48), np.uint16))
    with pytest.raises(RuntimeError):                       # more frames requested than available (:283-284)
        w.run(np.zeros((1, 32, 32), np.uint16))
    w.close()
    ip = make_params(32, 32, 2)
    ip._param_map['compression_scheme'] = 1
    with pytest.raises(NotImplementedError):                # no CPU fallback for other codecs
        ReCoDeWriter('x', dark_data=dark, output_directory=str(tmp_path), input_params=ip)


@pytest.mark.parametrize('level', [1, 2, 3, 4])
def test_degenerate_frames(tmp_path, level):
    """empty frames, fully foreground frames (one puddle covering everything, far beyond the shared-memory
    labelling capacity) and single-pixel events in the corners, through the whole writer / reader path"""
    from pyrecode_b200.recode_reader import ReCoDeReader
    ny, nx, b = 130, 192, 12
    dark = np.zeros((ny, nx), np.uint16)
    frames = np.zeros((5, ny, nx), np.uint16)
    frames[1] = 4095
    frames[2, 0, 0] = 7; frames[2, 0, nx - 1] = 9; frames[2, ny - 1, 0] = 11; frames[2, ny - 1, nx - 1] = 13
    frames[3, ::2, :] = 100                                  # every other row: ny/2 long puddles
    frames[4, :, ::2] = 200                                  # every other column: puddles crossing every tile
    ip = make_params(ny, nx, 5, level=level, b=b)
    write_parts(tmp_path, 'deg', frames, dark[None], ip, 1, batch_frames=2)
    path = str(tmp_path / ('deg.rc%d_part000' % level))
    hdr, recs = orc.parse_part_file(path)
    assert hdr['nz'] == 5
    r = ReCoDeReader(path, is_intermediate=True)
    r.open(print_header=False)
    for f in range(5):
        m_ref, v_ref, n_ref = orc.reduce_frame(frames[f], dark, level, b)
        assert recs[f]['map'] == m_ref, 'frame %d' % f
        if level <= 2:
            assert recs[f]['vals'] == v_ref, 'frame %d' % f
        (fid, fr), = r.get_next_frame().items()
        assert fid == f
        if level == 1:
            assert np.array_equal(fr['data'].toarray(), frames[f])
        else:
            bits = np.unpackbits(np.frombuffer(m_ref, np.uint8), bitorder='little')[:ny * nx].reshape(ny, nx)
            assert np.array_equal(fr['data'].toarray(), bits)
    r.close()


def test_compression_levels_and_table_reuse(tmp_path):
    """levels 0 (stored), 1 (shared sampled code, kept across batches), 9 (per-stream codes): same payloads.
    The second half of the run has very different statistics from the batch the kept code was built from."""
    rng = np.random.default_rng(21)
    nz, ny, nx = 12, 256, 256
    data = reference_test_data(rng, nz, ny, nx)
    data[nz // 2:] = rng.integers(0, 4096, size=(nz - nz // 2, ny, nx)).astype(np.uint16)     # dense noise
    dark = np.zeros((1, ny, nx), np.uint16)
    ref = None
    sizes = {}
    for cl in (0, 1, 9):
        ip = make_params(ny, nx, nz, clevel=cl)
        write_parts(tmp_path, 'lv%d' % cl, data, dark, ip, 1, batch_frames=2)
        p = str(tmp_path / ('lv%d.rc1_part000' % cl))
        h, recs = orc.parse_part_file(p)
        got = [(rr['map'], rr['vals']) for rr in recs]
        if ref is None:
            ref = got
            for f in range(nz):
                m, v, _ = orc.reduce_frame(data[f], dark[0], 1, 12)
                assert got[f] == (m, v)
        assert got == ref
        sizes[cl] = os.path.getsize(p)
    assert sizes[1] < sizes[0] and sizes[9] < sizes[0]


def test_write_sharded_single_rank(tmp_path):
    """pyrecode_b200.distributed on one rank (no process group): part file + merged file + live-view sum"""
    from pyrecode_b200 import distributed as rd
    rng = np.random.default_rng(5)
    nz, ny, nx = 7, 64, 128
    data = reference_test_data(rng, nz, ny, nx)
    ip = make_params(ny, nx, nz, threads=1)
    m = rd.write_sharded('one', data, np.zeros((1, ny, nx), np.uint16), str(tmp_path), ip, 0, 1)
    assert m['run_frames'] == nz
    h, recs = orc.parse_merged_file(str(tmp_path / 'one.rc1'))
    assert h['nz'] == nz
    ids, total = rd.live_view_sum(str(tmp_path / 'one.rc1_part000'), nz)
    assert ids == list(range(nz))
    assert np.array_equal(total.cpu().numpy().astype(np.int64), data.astype(np.int64).sum(0))
