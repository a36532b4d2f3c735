#!/bin/bash
mkdir -p gpurun_out
true
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 900 --csv --log-file gpurun_out/launches_pvdl.csv \
  python tools/bench_pvdl.py 16 3 0 nograph > gpurun_out/ncu_pvdl.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_pvdl.csv bridge_update_kernel > gpurun_out/launches_pvdl.md 2>&1; head -36 gpurun_out/launches_pvdl.md
