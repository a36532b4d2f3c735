#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/bench_pvdl.py 16 30 0 2>&1 | tail -2
P2PB_NO_GRAPH=1 P2PB_DUAL=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 900 --csv --log-file gpurun_out/launches_pvdl.csv \
  python tools/bench_pvdl.py 16 3 0 > gpurun_out/ncu_pvdl.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_pvdl.csv bridge_update_kernel > gpurun_out/launches_pvdl.md 2>&1; head -36 gpurun_out/launches_pvdl.md
