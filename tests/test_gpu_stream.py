"""GPU: streaming ingest -- synthetic `.seq` chunks dropped into a watched directory are reduced by
a ReCoDeWriter in mode='stream' straight from the SEQ reader's pinned buffers; the part file must hold the same records
as a batch-mode run over the same frames, and open in ReCoDeReader."""
import os
import sys
import threading
import time

import numpy as np
import pytest

from oracle import oracle as orc

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize('level,b', [(1, 12), (2, 12), (4, 12), (1, 8)])
def test_stream_ingest_equals_batch(tmp_path, level, b):
    from test_gpu_api import make_params
    from pyrecode_b200 import em_reader, stream
    from pyrecode_b200.recode_reader import ReCoDeReader
    from pyrecode_b200.recode_writer import ReCoDeWriter
    ny, nx, eps = 128, 256, 20 if b > 8 else 2
    dt = np.uint16 if b > 8 else np.uint8
    dark16 = orc.synth_dark(ny, nx)
    sizes = [4, 5, 7, 3, 6]                                   # chunk 0 is dropped by the queue rules
    allf = np.stack(orc.synth_frames('l2', sum(sizes), ny, nx, dark16, seed=11, bit_depth=12))
    if b <= 8:
        dark = (dark16 // 16).astype(np.uint8)
        allf = np.clip(allf.astype(np.int64) - 94, 0, 255).astype(np.uint8)
    else:
        dark = dark16
    chunks, o = [], 0
    for s in sizes:
        chunks.append(allf[o:o + s])
        o += s
    ram = tmp_path / 'ram'
    ram.mkdir()
    out = tmp_path / 'out'
    out.mkdir()

    def producer():
        time.sleep(0.2)
        for i, c in enumerate(chunks):
            em_reader.write_seq(str(ram / ('chunk_%03d.seq' % i)), c, allocated_frames=len(c) + (3 if i == 2 else 0))
            time.sleep(0.05)

    ip = make_params(ny, nx, 1, level=level, b=b, eps=eps)
    ip._param_map['source_file_type'] = 2          # SEQ chunks; the frame count comes from every chunk
    ip._param_map['num_frames'] = -1
    assert ip.validate()
    w = ReCoDeWriter(str(ram / stream.NEXT_STREAM), dark_data=dark[None], output_directory=str(out), input_params=ip,
                     mode='stream', run_name='acq7', batch_frames=3, validation_frame_gap=4)
    w.start()
    th = threading.Thread(target=producer)
    th.start()
    metrics = stream.run_stream(str(ram), w, max_count=4, chunk_time_in_sec=-1, poll_s=0.01, idle_timeout_s=20)
    th.join()
    w.close()
    assert [m['run_frames'] for m in metrics] == sizes[1:]
    part = str(out / ('acq7.rc%d_part000' % level))
    h, recs = orc.parse_part_file(part)
    assert h['nz'] == sum(sizes[1:]) and h['source_header_length'] == 1024
    assert open(part, 'rb').read()[512:512 + 1024] == bytes(1024)
    thr = orc.make_threshold(dark, eps, dtype=dt).astype(np.uint16)
    want = np.concatenate(chunks[1:])
    assert [r['frame_id'] for r in recs] == list(range(len(want)))
    for i, r in enumerate(recs):
        m, v, _ = orc.reduce_frame(want[i].astype(np.uint16), thr, level, b)
        assert r['map'] == m and r['vals'] == (v if level <= 2 else None), 'frame %d' % i
    # validation frames: ids 0, 4, 8, ... of the stream
    vf = open(str(out / 'acq7_part000_validation_frames.bin'), 'rb').read()
    assert vf == want[::4].tobytes()
    # and the file opens in the reader
    r = ReCoDeReader(part, is_intermediate=True)
    r.open(print_header=False)
    fr = r.get_next_frame()
    assert list(fr.keys()) == [0]
    d = fr[0]['data'].toarray()
    if level == 1:
        assert np.array_equal(d, np.where(want[0] > thr, want[0] - thr, 0))
    r.close()
    assert sorted(os.listdir(str(ram))) == ['chunk_000.seq']
