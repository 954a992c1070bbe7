cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_glm_tc.py -k "concurrent" -x -q -s > gpurun_out/r02_fused_tests_3.log 2>&1; echo "pytest rc=$?"
tail -7 gpurun_out/r02_fused_tests_3.log
run() { env "$@" timeout 200 python tools/glm_eval_bench.py --reps 9 --check 64 2>/dev/null >> gpurun_out/r02_fused_ab4.jsonl; }
rm -f gpurun_out/r02_fused_ab4.jsonl
run B2M_TC_FUSE=0
run B2M_TC_FUSE=1
run B2M_TC_FUSE=1 B2M_TC_FUSE_GROUPS5=36
run B2M_TC_FUSE=1 B2M_TC_FUSE_GROUPS5=37
run B2M_TC_FUSE=1 B2M_TC_FUSE_RING=3
run B2M_TC_FUSE=1 B2M_TC_FUSE_SLAB=3 B2M_TC_FUSE_GROUPS5=36
run B2M_TC_FUSE=1 B2M_TC_FUSE_SLAB=1 B2M_TC_FUSE_RING=4
run B2M_TC_FUSE=0
python - <<'PY'
import json
for l in open('gpurun_out/r02_fused_ab4.jsonl'):
    d=json.loads(l); print(d['knobs'], {k:(round(v,3) if isinstance(v,float) else v) for k,v in d['gemm_ms'].items()}, round(d['ms_per_eval'],3), d['grad_normwise_err'])
PY
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__cycles_elapsed.avg.per_second
nc() { tag=$1; shift; env "$@" timeout 200 ncu --metrics $M --clock-control none -k regex:tc_gemm -s 2 -c 2 --csv --log-file gpurun_out/r02_fz2_$tag.csv python tools/glm_eval_bench.py --reps 1 --check 0 > /dev/null 2>&1; echo "$tag $@"; grep -E "dram__bytes|gpu__time|hit_rate|tensor|per_second" gpurun_out/r02_fz2_$tag.csv | awk -F'","' '{print "   ", $5, $(NF-2), $(NF-1), $NF}' | cut -c1-150; }
nc s2r2 B2M_TC_FUSE=1
nc s3r2 B2M_TC_FUSE=1 B2M_TC_FUSE_SLAB=3
nc sep B2M_TC_FUSE=0
B2M_TC_FUSE=1 timeout 400 python bench.py --steps 10 --warmup 3 --no-others --no-ess --no-cpu-baseline > gpurun_out/r02_bench_fused_v2.json 2> gpurun_out/r02_bench_fused_v2.log; echo "bench fused rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_fused_v2.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['avg_launch_ms'], d['clocks'], d['issue'])
PY
