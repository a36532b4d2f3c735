#!/bin/bash
mkdir -p gpurun_out
for G in ${1:-0}; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_halo -s 2 -c 1 -f -o gpurun_out/prof_halo_r2_G$G python tools/prof_conv.py $G one > gpurun_out/prof_halo_r2_G$G.log 2>&1
echo "G=$G exit $?"
done
