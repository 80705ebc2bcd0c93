"""GPU, >= 2 devices: one process per GPU under torchrun -- frame-sharded ReCoDeWriter ranks, rank-0 merge_parts and
the NCCL all-reduced live-view image (skipped on a single-GPU box; `gpurun --gpus 2` runs it)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize('level', [1, 2])
def test_two_gpu_sharded_write_and_live_view(tmp_path, level):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip('needs 2 GPUs')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=2', '--master-addr', '127.0.0.1',
           '--master-port', str(29500 + level), os.path.join(HERE, 'multi_gpu_worker.py'), str(tmp_path), str(level)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and 'MULTI_GPU_OK world=2' in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
