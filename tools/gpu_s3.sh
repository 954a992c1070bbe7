#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
( time python bench.py --steps 10 --warmup 3 ) > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; echo "bench rc=$?"
tail -5 gpurun_out/bench_c4.err
cat gpurun_out/bench_c4.json
