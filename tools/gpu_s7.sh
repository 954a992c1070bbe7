#!/bin/bash
for w in near zero mixed; do
python tools/glm_eval_bench.py --n 10000 --d 100 --chains 1024 --path tc16 --where $w 2>&1 | tail -1 | cut -c1-700
python tools/glm_eval_bench.py --n 100000 --d 1000 --chains 4096 --path tc16 --check 32 --where $w 2>&1 | tail -1 | cut -c1-700
done
python tools/glm_eval_bench.py --n 100000 --d 1000 --chains 4096 --path tc --check 32 --where near 2>&1 | tail -1 | cut -c1-400
