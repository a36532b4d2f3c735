#!/bin/bash
run() { env "$@" timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-roofline --batch $BATCH 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$*', 'batch', $BATCH, round(d['value'],1), 'patches/s', round(d['ms_per_step'],1), 'ms')"; }
BATCH=64 run P2PB_CHAINS=1
BATCH=64 run P2PB_CHAINS=2
BATCH=128 run P2PB_CHAINS=1
BATCH=128 run P2PB_CHAINS=2
