ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_read_v47.csv python tests/read_bench.py > gpurun_out/ncu_read.log 2>&1
