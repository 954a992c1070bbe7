cd $GRAFT_REPO_ROOT
nvidia-smi -L
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tests/mgpu_check.py > gpurun_out/r02_mgpu_check_n2_v2.log 2>&1; echo "mgpu_check rc=$?"
tail -12 gpurun_out/r02_mgpu_check_n2_v2.log
timeout 300 python -m pytest tests -m gpu -x -q -k "multi_gpu or two_gpus or mgpu" > gpurun_out/r02_gputest_n2.log 2>&1; echo "pytest n2 rc=$?"; tail -3 gpurun_out/r02_gputest_n2.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_bench_n2_v5.json 2> gpurun_out/r02_bench_n2_v5.log; echo "bench n2 rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_n2_v5.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['scaling'], d.get('parity'), d['e2e']['value'], d['clocks'])
PY
