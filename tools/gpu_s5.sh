#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; echo "bench rc=$?"; tail -6 gpurun_out/bench_c4.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_c4.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['avg_launch_ms'], d['roofline']['executed_tf32_tflops'], d['roofline']['peak'], d['roofline']['gemm_share_of_step'])
print({k:(v['value'],v['roofline']['frac']) for k,v in d['other_workloads'].items()})
print(d['ess']); print(d['cpu_baseline']['value'])
PY
# ncu: launch list of the bench command, then full capture of the two GEMM kernels
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/bench_c4_launches.csv python bench.py --steps 2 --warmup 3 --no-others --no-ess --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log | cut -c1-300
ncu --set full --clock-control none --import-source on -k regex:tc_gemm -s 4 -c 2 -o gpurun_out/r01_tc_gemm_c4_v2 -f python tools/glm_eval_bench.py --n 100000 --d 1000 --chains 4096 --path tc --check 0 --reps 3 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
