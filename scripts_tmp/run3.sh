python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python tests/read_bench.py 2>&1 | tail -6
python bench.py --no-cpu --no-e2e > gpurun_out/bench_l2_v46.json 2> gpurun_out/err46.txt
python bench.py --no-cpu --no-e2e --level 1 > gpurun_out/bench_l1_v46.json 2>> gpurun_out/err46.txt
python bench.py --no-cpu --no-e2e --level 4 > gpurun_out/bench_l4_v46.json 2>> gpurun_out/err46.txt
