set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_glm_tc.py -k "concurrent" -x -q -s > gpurun_out/r02_fused_tests_2.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/r02_fused_tests_2.log
run() { env "$@" timeout 200 python tools/glm_eval_bench.py --reps 9 --check 64 2>/dev/null >> gpurun_out/r02_fused_ab2.jsonl; }
rm -f gpurun_out/r02_fused_ab2.jsonl
run B2M_TC_FUSE=0
run B2M_TC_FUSE=1
run B2M_TC_FUSE=1 B2M_TC_FUSE_SLAB=2
run B2M_TC_FUSE=1 B2M_TC_FUSE_SLAB=6 B2M_TC_FUSE_RING=2
run B2M_TC_FUSE=1 B2M_TC_FUSE_SLAB=8 B2M_TC_FUSE_RING=2
run B2M_TC_FUSE=1 B2M_TC_FUSE_RING=4
run B2M_TC_FUSE=1 B2M_TC_FUSE_GROUPS5=36
run B2M_TC_FUSE=1 B2M_TC_FUSE_GROUPS5=38
run B2M_TC_FUSE=0
python - <<'PY'
import json
for l in open('gpurun_out/r02_fused_ab2.jsonl'):
    d=json.loads(l); print(d['knobs'], d['gemm_ms'], round(d['ms_per_eval'],3), d['grad_normwise_err'])
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_gemm -s 2 -c 1 -o gpurun_out/r02_c4_fused2 -f python tools/glm_eval_bench.py --reps 1 --check 0 > gpurun_out/ncu_fused2.log 2>&1; echo "ncu rc=$?"
