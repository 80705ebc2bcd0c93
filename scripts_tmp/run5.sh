python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python tests/read_bench.py 2>&1 | tail -6
RB_FRAMES=64 python tests/read_bench.py 2>&1 | tail -6
RB_FRAMES=128 python tests/read_bench.py 2>&1 | tail -6
