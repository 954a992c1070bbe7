#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/glm_eval2.jsonl
for w in near zero mixed; do
python tools/glm_eval_bench.py --n 10000 --d 100 --chains 1024 --path simt --where $w >> gpurun_out/glm_eval2.jsonl 2>gpurun_out/glm_eval2.err
python tools/glm_eval_bench.py --n 10000 --d 100 --chains 1024 --path tc --where $w >> gpurun_out/glm_eval2.jsonl 2>>gpurun_out/glm_eval2.err
python tools/glm_eval_bench.py --n 100000 --d 1000 --chains 4096 --path tc --check 32 --where $w >> gpurun_out/glm_eval2.jsonl 2>>gpurun_out/glm_eval2.err
done
B2M_TC_CHUNK_RESID=32 python tools/glm_eval_bench.py --n 100000 --d 1000 --chains 4096 --path tc --check 32 --where near >> gpurun_out/glm_eval2.jsonl 2>>gpurun_out/glm_eval2.err
cat gpurun_out/glm_eval2.jsonl
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/pytest_gpu.log
ncu --set full --clock-control none --import-source on -k regex:tc_gemm -s 2 -c 2 -o gpurun_out/r01_tc_gemm_c4_v1 -f python tools/glm_eval_bench.py --n 100000 --d 1000 --chains 4096 --path tc --check 0 --reps 2 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
