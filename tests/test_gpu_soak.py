"""A short soak of the pipelined write path (tools/soak.py): random geometries (full-tile and ragged), levels, bit depths,
occupancies up to the tile-overflow path, thresholds down to two sigma of the noise, 3-4 batches in flight with slot
reuse; every record of every slot compared with the CPU oracle.  compute-sanitizer is not available on the GPU pool:
this and tests/test_gpu_bench_config.py are the race checks that run."""
import os
import sys

import pytest

pytestmark = pytest.mark.gpu


def test_soak_pipelined_write_path():
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tools'))
    import soak
    soak.main(iters=8, seed=11)


def test_soak_read_path():
    """tools/soak_read.py: random files written by ReCoDeWriter (several parts, merged), read back through the bulk
    pipeline, get_next_frame and get_frame; every frame against the oracle"""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tools'))
    import soak_read
    soak_read.main(iters=5, seed=5)
