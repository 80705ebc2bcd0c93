"""Multi-GPU driver of the hot path: one process per GPU (torchrun), frames sharded by contiguous range.

The reference scales by running N ReCoDeWriter workers, worker i taking the i-th contiguous share of every chunk
of frames and writing its own part file, merged afterwards (pyrecode/recode_writer.py:320-322,
pyrecode/recode_server.py:350-363, pyrecode/recode_reader.py:495-595).  Here a worker is a GPU rank.  Frames are
independent, so the write path needs NO collective: ranks only meet at a barrier before rank 0 merges the parts.
The read / live-view path sums each rank's frames on its GPU and all-reduces the uint32 image once per view
(NCCL over NVLink; the reference does `sums = np.add(sums, r['sum'])` over worker processes,
examples/ReCoDe_Live_View_MT.ipynb cell 1).

Everything here also runs on the gloo backend with CPU tensors (host logic tests without a GPU); the frame
arithmetic itself only exists on the GPU.
"""
import math
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """torchrun environment -> (rank, world_size, local_rank).  A single process without RANK is (0, 1, 0)."""
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        kw = {}
        if backend == 'nccl':
            torch.cuda.set_device(local_rank)
            kw['device_id'] = torch.device('cuda', local_rank)
        dist.init_process_group(backend, **kw)
    return rank, world, local_rank


def bind_to_gpu_cpus(local_rank):
    """Pin this process to the host CPUs the driver reports as local to its GPU (NVML cpu affinity), so that the
    rank's pinned staging buffers are first touched -- and the copies driven -- on the GPU's own NUMA node.  With 8
    ranks on one box the alternative is every rank allocating on node 0 and half of the host-to-device traffic
    crossing the socket interconnect.  Returns the number of CPUs bound to, or None when NVML is not usable."""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get('CUDA_VISIBLE_DEVICES')
        idx = int(vis.split(',')[local_rank]) if vis and all(x.strip().isdigit() for x in vis.split(',')) else local_rank
        h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        n_words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


def shard_frames(n_frames, world, rank):
    """(first frame, number of frames) of `rank`: the reference's partition rule (recode_writer.py:320-322)."""
    per = int(math.ceil(n_frames / float(world)))
    off = rank * per
    return off, min(per, max(n_frames - off, 0))


def barrier():
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()


def write_sharded(image_filename, data, dark_data, output_directory, input_params, rank, world, device=None,
                  merge=True, **writer_kw):
    """Every rank writes `<stem>.rc<level>_part<rank>` for its share of `data` ([nz, ny, nx], the same array on
    every rank, as in ReCoDeServer.run where every worker receives the whole chunk); rank 0 then merges.
    Returns this rank's run_metrics."""
    from pathlib import Path

    from .recode_reader import merge_parts
    from .recode_writer import ReCoDeWriter
    if input_params.num_threads != world:
        raise ValueError('input_params.num_threads (%d) must equal the number of ranks (%d): it is the divisor of '
                         'the partition rule' % (input_params.num_threads, world))
    w = ReCoDeWriter(image_filename, dark_data=dark_data, output_directory=output_directory, input_params=input_params,
                     mode='batch', node_id=rank, device=device, **writer_kw)
    w.start()
    metrics = w.run(data)
    w.close()
    barrier()
    if merge and rank == 0:
        base = Path(image_filename).stem + '.rc' + str(input_params.reduction_level)
        merge_parts(output_directory, base, world)
    barrier()
    return metrics


def allreduce_view(total):
    """Sum the per-rank live-view images in place (uint32 counts held as int32 / int64 tensors; CUDA -> NCCL,
    CPU -> gloo).  Exact integer sum, unlike the notebook's wrapping uint16 accumulator."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(total, op=dist.ReduceOp.SUM)
    return total


def live_view_sum(part_filename, n_frames, device=None, batch_frames=16):
    """This rank's part file -> summed image of its next `n_frames` frames, all-reduced over the ranks.
    Returns (frame ids read by this rank, uint32 image [ny, nx] as an int32 CUDA tensor, identical on all ranks)."""
    from .recode_reader import ReCoDeReader
    r = ReCoDeReader(part_filename, is_intermediate=True, device=device, batch_frames=batch_frames)
    r.open(print_header=False)
    ids, total = r.sum_frames(n_frames)
    _, ny, nx = r.get_shape()
    r.close()
    allreduce_view(total)
    return ids, total.view(ny, nx)
