#!/bin/bash
# A/B of two builds of the library on the same box, PVDL config 3 (see tools/gpu_ab.sh for how libp2pb_b200_prev.so is made)
for i in $(seq 1 ${1:-2}); do
  for v in A B; do
    if [ $v = A ]; then unset P2PB_LIB; else export P2PB_LIB=$PWD/p2pb_b200/libp2pb_b200_prev.so; fi
    echo -n "$v "; python tools/bench_pvdl.py 2>/dev/null | tail -1
  done
done | tee gpurun_out/ab_pvdl.txt
