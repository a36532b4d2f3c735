#!/bin/bash
# A/B of an environment switch: usage gpu_ab.sh VAR
for v in 1 0 1 0; do
env $1=$v timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-roofline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1=$v', round(d['value'],1), 'patches/s', round(d['ms_per_step'],1), 'ms', d['clocks']['sm_mhz'])"
done
