cd $GRAFT_REPO_ROOT
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r02_bench_n4_v5.json 2> gpurun_out/r02_bench_n4_v5.log; echo "bench n4 rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_n4_v5.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['scaling'], d.get('parity'), d['e2e']['value'], d['clocks'], d.get('issue'))
PY
