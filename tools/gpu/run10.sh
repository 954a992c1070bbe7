cd $GRAFT_REPO_ROOT
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_gemm_kernel -s 4 -c 2 -o gpurun_out/r02_c4_gemms_v3 -f python tools/glm_eval_bench.py --reps 2 --check 0 > gpurun_out/ncu_gemms_v3.log 2>&1; echo "ncu full rc=$?"
timeout 600 python bench.py > gpurun_out/r02_bench_default_v6.json 2> gpurun_out/r02_bench_default_v6.log; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_default_v6.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['avg_launch_ms'], d['e2e']['value'], d['clocks'], d['issue'])
PY
