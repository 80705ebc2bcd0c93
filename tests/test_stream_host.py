"""CPU: the streaming ingest's host logic -- the SEQ chunk reader / writer pair and the watched-directory queue rules of
ReCoDeServer._recode_queue_manager (pyrecode/recode_server.py:463-564)."""
import os
import threading
import time

import numpy as np
import pytest

from pyrecode_b200 import em_reader, stream


@pytest.mark.parametrize('dtype,version,pad', [(np.uint16, 3, None), (np.uint8, 3, None), (np.uint16, 5, None),
                                               (np.uint16, 3, 37 * 53 * 2 + 24)])
def test_seq_roundtrip(tmp_path, dtype, version, pad):
    rng = np.random.default_rng(3)
    frames = rng.integers(0, np.iinfo(dtype).max, size=(7, 37, 53)).astype(dtype)
    p = str(tmp_path / 'a.seq')
    em_reader.write_seq(p, frames, version=version, pad_to=pad)
    with em_reader.SEQReader(p) as f:
        assert f.shape == (7, 37, 53) and f.dtype == dtype
        assert np.array_equal(f[:], frames)
        assert np.array_equal(f[2:5], frames[2:5])
        assert np.array_equal(f[6], frames[6:7])
        out = np.zeros((4, 37, 53), dtype)
        assert f.read_into(out, 5, 9) == 2 and np.array_equal(out[:2], frames[5:7])
        with pytest.raises(IndexError):
            f[7]


def test_seq_header_ahead_of_file(tmp_path):
    """a chunk still being written: allocated_frames says 10, 4 frames are on disk"""
    frames = np.arange(4 * 6 * 8, dtype=np.uint16).reshape(4, 6, 8)
    p = str(tmp_path / 'a.seq')
    em_reader.write_seq(p, frames, allocated_frames=10)
    with em_reader.SEQReader(p) as f:
        assert f.shape[0] == 4 and np.array_equal(f[:], frames)
    with open(p, 'r+b') as fp:                          # half a frame more
        fp.seek(0, 2)
        fp.write(bytes(40))
    with em_reader.SEQReader(p) as f:
        assert f.shape[0] == 4


def test_seq_rejects_other_files(tmp_path):
    p = str(tmp_path / 'x.seq')
    open(p, 'wb').write(bytes(2048))
    with pytest.raises(ValueError):
        em_reader.SEQReader(p)


def test_queue_manager_rules(tmp_path):
    """the first chunk is dropped, a chunk is taken only once a newer one is queued, the last after the wait, max_count
    chunks in all, every processed chunk is renamed to Next_Stream.seq and deleted"""
    d = tmp_path / 'ram'
    d.mkdir()
    (d / 'stale.seq').write_bytes(b'old')               # cleared at session start
    seen = []

    def producer():
        time.sleep(0.1)
        for i in range(6):
            tmp = d / ('c%02d.tmp' % i)
            tmp.write_bytes(b'chunk %d' % i)
            os.rename(tmp, d / ('c%02d.seq' % i))
            time.sleep(0.05)

    def process(path, name):
        assert os.path.basename(path) == stream.NEXT_STREAM
        seen.append((name, open(path, 'rb').read()))

    th = threading.Thread(target=producer)
    th.start()
    done = stream.recode_queue_manager(str(d), 4, 0, process, poll_s=0.01, idle_timeout_s=5)
    th.join()
    assert done == ['c01.seq', 'c02.seq', 'c03.seq', 'c04.seq']
    assert seen == [('c%02d.seq' % i, b'chunk %d' % i) for i in (1, 2, 3, 4)]
    left = sorted(os.listdir(str(d)))
    assert 'c00.seq' in left and stream.NEXT_STREAM not in left and 'stale.seq' not in left


def test_queue_manager_single_chunk_and_timeout(tmp_path):
    d = tmp_path / 'ram'
    d.mkdir()
    got = []
    threading.Timer(0.1, lambda: (d / 'only.seq').write_bytes(b'x')).start()
    done = stream.recode_queue_manager(str(d), 1, -1, lambda p, n: got.append(n), poll_s=0.01, idle_timeout_s=5)
    assert done == ['only.seq'] == got
    # nothing arrives: the bounded wait returns what was done
    assert stream.recode_queue_manager(str(d), 3, -1, lambda p, n: got.append(n), poll_s=0.01, idle_timeout_s=0.2) == []
    with pytest.raises(ValueError):
        stream.recode_queue_manager(str(d / 'missing'), 1, 0, lambda p, n: None)
