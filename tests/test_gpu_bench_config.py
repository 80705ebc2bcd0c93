"""Parity at the BENCHMARKED configuration itself (bench.py's defaults): 4096 x 4096 uint16 frames, 32 frames per
launch, 3 pipelined slots (one context + CUDA stream each, priority streams, persistent labelling grid), several
rounds of slot reuse.  Every record of every slot is inflated with stock zlib and compared with the CPU oracle:
binary maps, packed statistics / intensities / centroid maps, counts and frame ids."""
import zlib

import numpy as np
import pytest

from oracle import oracle as orc

pytestmark = pytest.mark.gpu

NY = NX = 4096
F = 32
SLOTS = 3
EPS = 20
B = 12
KIND = {1: 'l1', 2: 'l2', 4: 'l4'}


def _check_batch(rec, offs, counts, level, first_id, expect, tag):
    rec = memoryview(rec)
    for i in range(F):
        r = bytes(rec[int(offs[i]):int(offs[i + 1])])
        m_ref, v_ref, n_ref = expect[i % len(expect)]
        hdr = np.frombuffer(r[:16 if level <= 2 else 8], dtype='<u4')
        assert hdr[0] == first_id + i, '%s frame %d: id %d' % (tag, i, hdr[0])
        if level <= 2:
            assert len(r) == 16 + hdr[1] + hdr[2], '%s frame %d: record length' % (tag, i)
            assert zlib.decompress(r[16:16 + hdr[1]]) == m_ref, '%s frame %d: map' % (tag, i)
            assert zlib.decompress(r[16 + hdr[1]:]) == v_ref and hdr[3] == len(v_ref), '%s frame %d: values' % (tag, i)
        else:
            assert len(r) == 8 + hdr[1], '%s frame %d: record length' % (tag, i)
            assert zlib.decompress(r[8:]) == m_ref, '%s frame %d: centroid map' % (tag, i)
        assert int(counts[i]) == n_ref, '%s frame %d: count %d != %d' % (tag, i, counts[i], n_ref)


@pytest.mark.parametrize('level', [2, 4, 1])
def test_bench_configuration_parity(level):
    import torch
    from pyrecode_b200.engine import WriteEngine
    from pyrecode_b200.synth import synth_dark, synth_frames
    dark = synth_dark(NY, NX)
    frames = synth_frames(KIND[level], 4, NY, NX, dark, seed=1234, bit_depth=B)
    thr = orc.make_threshold(dark, EPS)
    expect = [orc.reduce_frame(f, thr, level, B) for f in frames]
    eng = WriteEngine(NY, NX, 2, B, level, 1, 0, 0, 1, max_frames=F, records_capacity=F * (NY * NX * 2 // 4),
                      n_slots=SLOTS)
    eng.set_threshold(dark, EPS)
    host = torch.empty((F, NY, NX), dtype=torch.uint16).pin_memory()
    hv = host.numpy()
    for i in range(F):
        hv[i] = frames[i % len(frames)]
    d_frames = host.to(eng.dev)
    torch.cuda.synchronize()

    # ---- device-resident steps issued round-robin on the slots' streams, exactly like bench.py's run_steps
    cur = torch.cuda.current_stream()
    rounds = 4
    for sl in eng.slots:
        sl.stream.wait_stream(cur)
    for s in range(rounds * SLOTS):
        sl = eng.slots[s % SLOTS]
        with torch.cuda.stream(sl.stream):
            eng.launch(d_frames, F, 1000 + s * F, s % SLOTS)
    for sl in eng.slots:
        cur.wait_stream(sl.stream)
    torch.cuda.synchronize()
    for k, sl in enumerate(eng.slots):
        assert int(sl.status.cpu()[0]) == 0
        offs = sl.offsets.cpu().numpy()
        counts = sl.counts.cpu().numpy()
        rec = sl.records[:int(offs[F])].cpu().numpy()
        last_step = (rounds - 1) * SLOTS + k
        _check_batch(rec, offs, counts, level, 1000 + last_step * F, expect, 'resident slot %d' % k)

    # ---- the host-buffer path (bench.py's e2e leg): pinned frames -> H2D -> kernels -> D2H, 3 batches in flight
    pending = []
    for s in range(2 * SLOTS + 1):
        pending.append((eng.submit(host, first_frame_id=50000 + s * F), 50000 + s * F))
        if len(pending) == SLOTS:
            k, fid = pending.pop(0)
            rec, offs, counts, _, _ = eng.collect(k)
            _check_batch(rec, offs, counts, level, fid, expect, 'e2e step id %d' % fid)
    while pending:
        k, fid = pending.pop(0)
        rec, offs, counts, _, _ = eng.collect(k)
        _check_batch(rec, offs, counts, level, fid, expect, 'e2e step id %d' % fid)


@pytest.mark.parametrize('level', [1, 2])
def test_validation_frames_do_not_disturb_batches_in_flight(tmp_path, level):
    """validation_frame_gap > 0 with more than two batches: the validation frame is reduced on its own engine while
    the next batch is in flight on the writer's slots; the part file must equal the one written without validation"""
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from test_gpu_api import make_params
    from pyrecode_b200.recode_writer import ReCoDeWriter
    nz, ny, nx = 14, 256, 512
    dark = orc.synth_dark(ny, nx)
    frames = np.stack(orc.synth_frames('l2', nz, ny, nx, dark, seed=5, bit_depth=B))
    out = {}
    for gap in (-1, 1):
        d = tmp_path / ('gap%d' % gap)
        d.mkdir()
        ip = make_params(ny, nx, nz, level=level, b=B, eps=EPS)
        w = ReCoDeWriter('val', dark_data=dark[None], output_directory=str(d), input_params=ip, mode='batch',
                         validation_frame_gap=gap, batch_frames=3)
        w.start()
        m = w.run(frames)
        w.close()
        out[gap] = open(os.path.join(str(d), 'val.rc%d_part000' % level), 'rb').read()
        if gap > 0:
            assert len(m['run_dose_rates']) == nz
            vf = open(os.path.join(str(d), 'val_part000_validation_frames.bin'), 'rb').read()
            assert vf == frames.tobytes()
            # dose rate = puddles of the central 128 x 128 ROI / ROI area (recode_writer.py:402-415)
            thr = orc.make_threshold(dark, EPS)
            y0, x0 = (ny - 128) // 2, (nx - 128) // 2
            for i in (0, nz - 1):
                _, k = orc.label8((frames[i] > thr)[y0:y0 + 128, x0:x0 + 128])
                assert abs(m['run_dose_rates'][i] - k / (128 * 128)) < 1e-12
    assert out[-1] == out[1]
