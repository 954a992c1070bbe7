cd $GRAFT_REPO_ROOT
timeout 500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputest_8.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_gputest_8.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/r02_bench_default_v5.json 2> gpurun_out/r02_bench_default_v5.log; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_default_v5.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['avg_launch_ms'], d['e2e']['value'], d['clocks'], d['issue'])
print({k:(v.get('value'), v.get('e2e',{}).get('value') if isinstance(v.get('e2e'),dict) else None) for k,v in d.get('other_workloads',{}).items()})
print(d.get('ess'))
PY
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 1200 --csv --log-file gpurun_out/r02_bench_c4_v2_launches.csv python bench.py --steps 2 --warmup 3 --no-others --no-ess --no-cpu-baseline > gpurun_out/ncu_launches_v2.log 2>&1; echo "ncu launches rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_gemm_kernel -s 4 -c 2 -o gpurun_out/r02_c4_gemms_v2 -f python tools/glm_eval_bench.py --reps 2 --check 0 > gpurun_out/ncu_gemms_v2.log 2>&1; echo "ncu full rc=$?"
