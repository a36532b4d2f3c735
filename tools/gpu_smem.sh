#!/bin/bash
mkdir -p gpurun_out
for kb in 227 200 180 160; do
  for dual in 1 0; do
    P2PB_SMEM_KB=$kb P2PB_DUAL=$dual timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-roofline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('smem_kb=$kb dual=$dual', round(d['value'],1), 'patches/s')"
  done
done
