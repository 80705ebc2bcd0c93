"""Soak run of the read path against the CPU oracle (test infrastructure): random geometries, levels, bit depths,
occupancies and batch sizes; files written by ReCoDeWriter (part files and a merged file), read back with
read_frames_dense / sum_frames (several batches in flight, one engine + stream each), with get_next_frame (scipy COO +
L2 statistics) and with get_frame (random access), every frame compared with the oracle's reduce -> unpack of the input.
usage: python tools/soak_read.py [iterations] [seed]   (needs a GPU)"""
import os
import shutil
import sys
import tempfile
import time

sys.path.insert(0, '.')
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests'))
import numpy as np

from oracle import oracle as orc
from soak import frames_of


def main(iters=8, seed=77):
    from test_gpu_api import make_params
    from pyrecode_b200.recode_reader import ReCoDeReader, merge_parts
    from pyrecode_b200.recode_writer import ReCoDeWriter
    rng = np.random.default_rng(seed)
    geoms = [(2048, 4096), (1000, 1200), (1536, 2048), (777, 4100), (512, 512)]
    t00 = time.time()
    for it in range(iters):
        ny, nx = geoms[int(rng.integers(len(geoms)))]
        level = int(rng.choice([1, 2, 3, 4]))
        b = int(rng.choice([12, 10, 14]))
        p = float(rng.choice([0.001, 0.0075, 0.02]))
        grow = float(rng.choice([0.2, 0.6]))
        eps = int(rng.choice([8, 20]))
        nz = int(rng.integers(9, 41))
        nodes = int(rng.choice([1, 2, 3]))
        bulk = int(rng.choice([2, 5, 16]))
        inflight = int(rng.choice([2, 4, 6]))
        dark = orc.synth_dark(ny, nx)
        distinct = frames_of(rng, 4, ny, nx, dark, p, grow, b)
        frames = np.stack([distinct[i % 4] for i in range(nz)])
        thr = orc.make_threshold(dark, eps)
        ref = []
        for f in distinct:
            m, v, n = orc.reduce_frame(f, thr, level, b)
            dense = orc.unpack_dense(ny, nx, b, m, v, level)
            stats = orc.bit_unpack(np.frombuffer(v + b'\0' * 8, dtype=np.uint8), n, b) if level == 2 else None
            ref.append((dense, stats))
        tmp = tempfile.mkdtemp(prefix='soak_read_', dir='/dev/shm' if os.path.isdir('/dev/shm') else None)
        try:
            ip = make_params(ny, nx, nz, level=level, b=b, eps=eps, threads=nodes)
            for node in range(nodes):
                w = ReCoDeWriter('s', dark_data=dark[None], output_directory=tmp, input_params=ip, mode='batch',
                                 node_id=node, batch_frames=int(rng.choice([3, 8, 16])))
                w.start()
                w.run(frames)
                w.close()
            merge_parts(tmp, 's.rc%d' % level, nodes)
            r = ReCoDeReader(os.path.join(tmp, 's.rc%d' % level), bulk_frames=bulk, bulk_inflight=inflight)
            r.open(print_header=False)
            k = int(rng.integers(1, nz))
            ids, dense = r.read_frames_dense(k)
            assert ids == list(range(k)), (it, 'ids')
            dn = dense.cpu().numpy()
            for i in range(k):
                assert np.array_equal(dn[i], ref[i % 4][0]), (it, 'dense frame', i)
            ids2, total = r.sum_frames(10 ** 6)
            assert ids2 == list(range(k, nz)), (it, 'ids of the sum')
            want = np.zeros((ny, nx), dtype=np.int64)
            for i in range(k, nz):
                want += ref[i % 4][0]
            assert np.array_equal(total.cpu().numpy().astype(np.int64).reshape(ny, nx), want), (it, 'live-view sum')
            r.rewind()
            for i in range(min(nz, 7)):
                (fid, fr), = r.get_next_frame().items()
                assert fid == i and np.array_equal(fr['data'].toarray(), ref[i % 4][0]), (it, 'get_next_frame', i)
                if level == 2:
                    assert np.array_equal(np.asarray(fr['summary_stats']).astype(np.int64),
                                          np.asarray(ref[i % 4][1]).astype(np.int64)), (it, 'summary_stats', i)
            z = int(rng.integers(nz))
            (fid, fr), = r.get_frame(z).items()
            assert fid == z and np.array_equal(fr['data'].toarray(), ref[z % 4][0]), (it, 'get_frame', z)
            r.close()
        finally:
            shutil.rmtree(tmp, ignore_errors=True)
        print('it %2d ok: %4dx%4d L%d b=%d p=%.4f eps=%d nz=%d parts=%d bulk=%d x %d in flight'
              % (it, ny, nx, level, b, p, eps, nz, nodes, bulk, inflight), flush=True)
    print('SOAK_READ_OK %d iterations in %.0f s' % (iters, time.time() - t00))


if __name__ == '__main__':
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 8, int(sys.argv[2]) if len(sys.argv) > 2 else 77)
