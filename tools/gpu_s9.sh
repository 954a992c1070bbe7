#!/bin/bash
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 400 --csv --log-file gpurun_out/bench_c4_launches_v3.csv python bench.py --steps 2 --warmup 3 --no-others --no-ess --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -1 gpurun_out/ncu_bench.log | cut -c1-200
ncu --set full --clock-control none --import-source on -k regex:tc_gemm -s 4 -c 2 -o gpurun_out/r01_tc_gemm_c4_v3_f16 -f python tools/glm_eval_bench.py --n 100000 --d 1000 --chains 4096 --path tc16 --check 0 --reps 3 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
