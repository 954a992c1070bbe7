set -x
cd $GRAFT_REPO_ROOT
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputest_6.log 2>&1; echo "pytest rc=$?"
for h in 1 0 1 0; do B2M_TC_L2_HINTS=$h timeout 200 python tools/glm_eval_bench.py --reps 9 --check 64 2>/dev/null | sed "s/^/hints=$h /" >> gpurun_out/r02_l2hints_ab.jsonl; done
cat gpurun_out/r02_l2hints_ab.jsonl
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_gemm_kernel -s 2 -c 2 -o gpurun_out/r02_c4_gemms_l2hints -f python tools/glm_eval_bench.py --reps 2 --check 0 > gpurun_out/ncu_l2hints.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/r02_gputest_6.log
