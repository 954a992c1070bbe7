#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/tc_accuracy_probe.py 2>&1 | tail -12
timeout 300 python tools/glm_eval_bench.py --n 100000 --d 1000 --chains 4096 --path tc --check 32 2>&1 | tail -2
B2M_TC_PAIR=0 timeout 300 python tools/glm_eval_bench.py --n 100000 --d 1000 --chains 4096 --path tc --check 32 2>&1 | tail -1
timeout 300 python tools/glm_eval_bench.py --n 10000 --d 100 --chains 1024 --path tc --check 32 2>&1 | tail -1
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
