for lv in 1 4; do
for cfg in "0 0 2" "1 0 2" "1 0 3" "1 2 2" "1 2 3" "1 3 3"; do set -- $cfg
RECODE_B200_PRIORITY=$1 RECODE_B200_CCL_CTAS=$2 python bench.py --level $lv --no-cpu --no-e2e --slots $3 > gpurun_out/bench_l${lv}_v45_p$1_c$2_s$3.json 2>> gpurun_out/err45.txt
done; done
RECODE_B200_PRIORITY=1 RECODE_B200_CCL_CTAS=2 python bench.py --no-cpu --no-e2e --slots 4 > gpurun_out/bench_l2_v45_p1_c2_s4.json 2>> gpurun_out/err45.txt
RECODE_B200_PRIORITY=1 RECODE_B200_CCL_CTAS=3 python bench.py --no-cpu --no-e2e --slots 3 > gpurun_out/bench_l2_v45_p1_c3_s3.json 2>> gpurun_out/err45.txt
RECODE_B200_PRIORITY=1 RECODE_B200_CCL_CTAS=2 python bench.py --no-cpu --no-e2e --slots 3 --frames 64 > gpurun_out/bench_l2_v45_p1_c2_s3_f64.json 2>> gpurun_out/err45.txt
