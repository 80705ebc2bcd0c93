"""Worker of tests/test_gpu_multi.py (launched by torchrun, one rank per GPU): frame-sharded write, merge, and the
NCCL all-reduced live-view sum, checked against the oracle on every rank."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

from oracle import oracle as orc                      # noqa: E402
from pyrecode_b200 import distributed as rd           # noqa: E402
from test_gpu_api import make_params                  # noqa: E402


def main():
    out_dir, level = sys.argv[1], int(sys.argv[2])
    rank, world, local = rd.init_from_env('nccl')
    nz, ny, nx, b, eps = 13, 128, 256, 12, 4
    dark = orc.synth_dark(ny, nx)
    frames = orc.synth_frames('l2', nz, ny, nx, dark + 16, seed=31, bit_depth=b)
    ip = make_params(ny, nx, nz, level=level, b=b, eps=eps, threads=world)
    m = rd.write_sharded('multi', frames, dark[None], out_dir, ip, rank, world, device=local, batch_frames=3)
    off, cnt = rd.shard_frames(nz, world, rank)
    assert m['run_frames'] == cnt
    thr = orc.make_threshold(dark, eps)
    if rank == 0:
        h, recs = orc.parse_merged_file(os.path.join(out_dir, 'multi.rc%d' % level))
        assert h['nz'] == nz and [r['frame_id'] for r in recs] == list(range(nz))
        for i, r in enumerate(recs):
            mm, vv, _ = orc.reduce_frame(frames[i], thr, level, b)
            assert r['map'] == mm and r['vals'] == (vv if level <= 2 else None), 'frame %d' % i
    ids, total = rd.live_view_sum(os.path.join(out_dir, 'multi.rc%d_part%03d' % (level, rank)), nz, device=local,
                                  batch_frames=4)
    assert ids == list(range(off, off + cnt))
    want = np.zeros((ny, nx), dtype=np.int64)
    for i in range(nz):
        mm, vv, _ = orc.reduce_frame(frames[i], thr, level, b)
        want += orc.unpack_dense(ny, nx, b, mm, vv, level).astype(np.int64)
    assert np.array_equal(total.cpu().numpy().astype(np.int64), want), 'rank %d live view' % rank
    rd.barrier()
    if rank == 0:
        print('MULTI_GPU_OK world=%d level=%d' % (world, level))
    torch.distributed.destroy_process_group()


if __name__ == '__main__':
    main()
