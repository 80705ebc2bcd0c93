"""CPU, world_size 2 over gloo: the N > 1 host logic of pyrecode_b200.distributed -- the reference's partition rule
(recode_writer.py:320-322), one part file per rank, rank-0 merge after a barrier (recode_reader.py:495-595) and the
all-reduced live-view sum (examples/ReCoDe_Live_View_MT.ipynb cell 1).  The per-frame arithmetic, which only exists
on the GPU in the product, is supplied here by the CPU oracle so that the multi-rank plumbing runs without a GPU."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from oracle import oracle as orc

NZ, NY, NX, B, EPS = 11, 48, 80, 12, 3


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _inputs():
    dark = orc.synth_dark(NY, NX)
    frames = orc.synth_frames('l1', NZ, NY, NX, dark + 15, seed=77, bit_depth=B)
    return dark, frames


def _worker(rank, world, port, tmp):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR='127.0.0.1',
                      MASTER_PORT=str(port))
    from pyrecode_b200 import distributed as rd
    from pyrecode_b200.recode_reader import ReCoDeReader, merge_parts
    r, w, _ = rd.init_from_env('gloo')
    assert (r, w) == (rank, world)
    dark, frames = _inputs()
    thr = orc.make_threshold(dark, EPS)
    off, cnt = rd.shard_frames(NZ, world, rank)
    assert (off, cnt) == orc.partition(NZ, world, rank)
    # this rank's part file (container bytes from the oracle: header + records of its frame range)
    hdr = dict(uid=orc.UID, version_major=0, version_minor=2, is_intermediate=1, reduction_level=1, rc_operation_mode=1,
               is_bit_packed=1, target_bit_depth=B, nx=NX, ny=NY, nz=cnt, compression_scheme=0, compression_level=1,
               source_file_name='dist', calibration_file_name='', calibration_threshold_epsilon=EPS,
               source_bit_depth=B, source_dtype=0, target_dtype=0)
    part = os.path.join(tmp, 'dist.rc1_part%03d' % rank)
    local = np.zeros((NY, NX), dtype=np.int64)
    with open(part, 'wb') as f:
        f.write(orc.build_header(hdr))
        for i in range(off, off + cnt):
            m, v, _ = orc.reduce_frame(frames[i], thr, 1, B)
            f.write(orc.build_record(i, 1, 1, m, v))
            local += orc.unpack_dense(NY, NX, B, m, v, 1).astype(np.int64)
    rd.barrier()
    if rank == 0:
        merge_parts(tmp, 'dist.rc1', world)
    rd.barrier()
    # every rank can open the merged file: host-side metadata / seek table only (no decode without a GPU)
    rr = ReCoDeReader(os.path.join(tmp, 'dist.rc1'))
    rr.open(print_header=False)
    assert rr.get_shape() == (NZ, NY, NX)
    rr.close()
    total = torch.from_numpy(local.astype(np.int32).ravel().copy())
    rd.allreduce_view(total)
    want = np.where(frames > thr, frames - thr, 0).astype(np.int64).sum(0)
    assert np.array_equal(total.numpy().reshape(NY, NX), want), 'rank %d: live-view sum differs' % rank
    torch.distributed.destroy_process_group()


def test_two_ranks_gloo(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    # the merged file holds every frame once, in order, with the payloads of the parts
    h, recs = orc.parse_merged_file(str(tmp_path / 'dist.rc1'))
    assert h['nz'] == NZ and [r['frame_id'] for r in recs] == list(range(NZ))
    dark, frames = _inputs()
    thr = orc.make_threshold(dark, EPS)
    for i, r in enumerate(recs):
        m, v, _ = orc.reduce_frame(frames[i], thr, 1, B)
        assert r['map'] == m and r['vals'] == v


@pytest.mark.parametrize('n,world', [(9, 3), (10, 4), (1, 8), (0, 2), (17, 8)])
def test_shard_rule_covers_every_frame_once(n, world):
    from pyrecode_b200.distributed import shard_frames
    seen = []
    for r in range(world):
        off, cnt = shard_frames(n, world, r)
        assert (off, cnt) == orc.partition(n, world, r)
        seen += list(range(off, off + cnt))
    assert seen == list(range(n))
