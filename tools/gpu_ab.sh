#!/bin/bash
# A/B of two builds of the library on the SAME box (box-to-box variance of bench.py is ~1 %): alternates
#   A = p2pb_b200/libp2pb_b200.so   B = p2pb_b200/libp2pb_b200_prev.so (build of an older tree, see below)
# usage: gpurun -- 'bash tools/gpu_ab.sh [rounds]'
#   git stash; python -m p2pb_b200.build --force; cp p2pb_b200/libp2pb_b200.so p2pb_b200/libp2pb_b200_prev.so; git stash pop; python -m p2pb_b200.build --force
mkdir -p gpurun_out
for i in $(seq 1 ${1:-3}); do
  for v in A B; do
    if [ $v = A ]; then unset P2PB_LIB; else export P2PB_LIB=$PWD/p2pb_b200/libp2pb_b200_prev.so; fi
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extra --no-roofline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$v', round(d['value'],1), round(d['e2e']['value'],1), d['clocks']['sm_mhz'])"
  done
done | tee gpurun_out/ab.txt
