#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_halo -s 3 -c 5 -o gpurun_out/prof_halo_v7_b64 -f python tools/prof_conv.py > gpurun_out/prof_halo_v7_b64.log 2>&1
tail -5 gpurun_out/prof_halo_v7_b64.log; ls -la gpurun_out/*.ncu-rep
