#!/bin/bash
run() { timeout 300 python bench.py --chains $1 --steps 3 --warmup 3 --no-cpu-baseline --no-roofline --batch $BATCH 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('chains=$1', 'batch', $BATCH, round(d['value'],1), 'patches/s', round(d['ms_per_step'],1), 'ms')"; }
BATCH=64 run 1
BATCH=64 run 2
BATCH=128 run 1
BATCH=128 run 2
