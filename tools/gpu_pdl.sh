#!/bin/bash
for pdl in 1 0 1 0; do
P2PB_PDL=$pdl timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-roofline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('PDL=$pdl', round(d['value'],1), 'patches/s', round(d['ms_per_step'],1), 'ms', d['clocks']['sm_mhz'])"
done
