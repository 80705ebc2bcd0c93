"""GPU, >= 2 devices: one process per GPU under torchrun -- frame-sharded ReCoDeWriter ranks, rank-0 merge_parts and
the NCCL all-reduced live-view image (skipped on a single-GPU box; `gpurun --gpus 2` runs it)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize('world', [2, 4])
@pytest.mark.parametrize('level', [1, 2, 4])
def test_multi_gpu_sharded_write_and_live_view(tmp_path, level, world):
    """13 frames over 2 ranks (7 + 6) and over 4 ranks (4 + 4 + 4 + 1: an uneven last share)."""
    import torch
    n = torch.cuda.device_count()
    if n < world:
        pytest.skip('needs %d GPUs' % world)
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=%d' % world,
           '--master-addr', '127.0.0.1', '--master-port', str(29500 + 10 * world + level),
           os.path.join(HERE, 'multi_gpu_worker.py'), str(tmp_path), str(level)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and 'MULTI_GPU_OK world=%d' % world in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
    print(r.stdout[-300:])
