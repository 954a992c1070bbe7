#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"; tail -4 gpurun_out/bench_default.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_default.json').read().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], d['roofline']['avg_launch_ms'], d['roofline']['executed_tf32_tflops'], d['roofline']['gemm_share_of_step'])
print({k:(v['value'],v['e2e'],v['roofline']['frac']) for k,v in d['other_workloads'].items()})
print(d['ess']); print(d['cpu_baseline']['value'], d['clocks'])
PY
python bench.py --workload c2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2.json 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:hmc_kernel -s 4 -c 1 -o gpurun_out/r01_hmc_kernel_c2_v2 -f python bench.py --workload c2 --steps 3 --warmup 3 --no-cpu-baseline --no-ess > gpurun_out/ncu_c2.log 2>&1
tail -2 gpurun_out/ncu_c2.log | cut -c1-200
