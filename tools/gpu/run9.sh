cd $GRAFT_REPO_ROOT
run() { env "$@" timeout 200 python tools/glm_eval_bench.py --reps 15 --check 64 2>/dev/null >> gpurun_out/r02_k6order_ab.jsonl; }
rm -f gpurun_out/r02_k6order_ab.jsonl
run B2M_TC_K6_ORDER=0
run B2M_TC_K6_ORDER=1
run B2M_TC_K6_ORDER=0
run B2M_TC_K6_ORDER=1
python - <<'PY'
import json
for l in open('gpurun_out/r02_k6order_ab.jsonl'):
    d=json.loads(l); print(d['knobs'], {k:(round(v,3) if isinstance(v,float) else v) for k,v in d['gemm_ms'].items()}, round(d['ms_per_eval'],3), d['grad_normwise_err'])
PY
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__cycles_elapsed.avg.per_second
for o in 0 1; do B2M_TC_K6_ORDER=$o timeout 200 ncu --metrics $M --clock-control none -k regex:tc_gemm_kernel -s 4 -c 4 --csv --log-file gpurun_out/r02_k6order_$o.csv python tools/glm_eval_bench.py --reps 3 --check 0 > /dev/null 2>&1; done
python - <<'PY'
import csv
for tag in ('0','1'):
    rows=list(csv.reader(open(f'gpurun_out/r02_k6order_{tag}.csv')))
    h=[i for i,r in enumerate(rows) if 'Metric Name' in r][0]
    hdr=rows[h]; ik=hdr.index('Kernel Name'); im=hdr.index('Metric Name'); iv=hdr.index('Metric Value'); iid=hdr.index('ID')
    out={}
    for r in rows[h+1:]:
        out.setdefault((r[iid], r[ik][22:52]),{})[r[im].split('.')[0][-24:]]=r[iv]
    print('order',tag)
    for k,v in out.items(): print('  ',k, v)
PY
timeout 300 python -m pytest tests/test_gpu_glm_tc.py tests/test_gpu_parity.py -x -q 2>&1 | tail -2
