#!/bin/bash
# scratch GPU session: full gpu test-suite + GLM evaluation timings (C3, C4) with promotion-interval sweep
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" 
tail -3 gpurun_out/pytest_gpu.log
: > gpurun_out/glm_eval.jsonl
python tools/glm_eval_bench.py --n 10000 --d 100 --chains 1024 --path simt >> gpurun_out/glm_eval.jsonl 2>gpurun_out/glm_eval.err
python tools/glm_eval_bench.py --n 10000 --d 100 --chains 1024 --path tc >> gpurun_out/glm_eval.jsonl 2>>gpurun_out/glm_eval.err
for cr in 1 2 4; do for cg in 1 2 4; do
  echo "{\"chunk_resid\": $cr, \"chunk_grad\": $cg}" >> gpurun_out/glm_eval.jsonl
  B2M_TC_CHUNK_RESID=$cr B2M_TC_CHUNK_GRAD=$cg python tools/glm_eval_bench.py --n 100000 --d 1000 --chains 4096 --path tc --check 32 >> gpurun_out/glm_eval.jsonl 2>>gpurun_out/glm_eval.err
done; done
cat gpurun_out/glm_eval.jsonl
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/glm_c4_launches.csv python tools/glm_eval_bench.py --n 100000 --d 1000 --chains 4096 --path tc --check 0 --reps 2 > gpurun_out/ncu_glm.log 2>&1
tail -20 gpurun_out/glm_c4_launches.csv
