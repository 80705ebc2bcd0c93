python -m pytest tests -m gpu -x -q 2>&1 | tail -8
RB_FRAMES=64 python tests/read_bench.py 2>&1 | tail -14
