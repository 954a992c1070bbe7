set -x
cd $GRAFT_REPO_ROOT
timeout 500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputest_7.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r02_gputest_7.log
run() { env "$@" timeout 200 python tools/glm_eval_bench.py --reps 9 --check 64 2>/dev/null >> gpurun_out/r02_fused_ab3.jsonl; }
rm -f gpurun_out/r02_fused_ab3.jsonl
run B2M_TC_FUSE=0
run B2M_TC_FUSE=1
run B2M_TC_FUSE=1 B2M_TC_FUSE_GROUPS5=33
run B2M_TC_FUSE=1 B2M_TC_FUSE_GROUPS5=37
run B2M_TC_FUSE=1 B2M_TC_FUSE_RING=3
run B2M_TC_FUSE=1 B2M_TC_FUSE_SLAB=3
run B2M_TC_FUSE=0
python - <<'PY'
import json
for l in open('gpurun_out/r02_fused_ab3.jsonl'):
    d=json.loads(l); print(d['knobs'], d['gemm_ms'], round(d['ms_per_eval'],3), d['grad_normwise_err'])
PY
B2M_TC_FUSE=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_gemm -s 2 -c 1 -o gpurun_out/r02_c4_fused3 -f python tools/glm_eval_bench.py --reps 1 --check 0 > gpurun_out/ncu_fused3.log 2>&1; echo "ncu rc=$?"
B2M_TC_FUSE=1 timeout 400 python bench.py --steps 10 --warmup 3 --no-others --no-ess --no-cpu-baseline > gpurun_out/r02_bench_fused_v1.json 2> gpurun_out/r02_bench_fused_v1.log; echo "bench fused rc=$?"
timeout 400 python bench.py --steps 10 --warmup 3 --no-others --no-ess --no-cpu-baseline > gpurun_out/r02_bench_sep_v1.json 2> gpurun_out/r02_bench_sep_v1.log; echo "bench sep rc=$?"
python - <<'PY'
import json
for f in ('fused','sep'):
    try:
        d=json.loads(open(f'gpurun_out/r02_bench_{f}_v1.json').read().strip().splitlines()[-1])
        print(f, d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['avg_launch_ms'], d['clocks'])
    except Exception as e: print(f, 'ERR', e)
PY
