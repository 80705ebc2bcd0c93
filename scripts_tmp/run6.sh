RB_FRAMES=64 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_read_v48.csv python tests/read_bench.py > gpurun_out/ncu_read.log 2>&1
RB_FRAMES=64 ncu --set full --clock-control none --import-source on -k regex:k_inflate_lanes -c 1 -o gpurun_out/prof_v48_lanes -f python tests/read_bench.py > gpurun_out/ncu_read2.log 2>&1
