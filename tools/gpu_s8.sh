#!/bin/bash
for cr in 2 4 8; do for cg in 2 4; do
echo "tc16 chunk_resid=$cr chunk_grad=$cg"
B2M_TC_CHUNK_RESID=$cr B2M_TC_CHUNK_GRAD=$cg python tools/glm_eval_bench.py --path tc16 --n 100000 --d 1000 --chains 4096 --check 32 --where near --reps 9 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_eval'], d['logp_rel_err'], d['grad_normwise_err'], d['grad_err_per_chain_max'])"
done; done
