#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 639 -c 639 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 1 --warmup 0 --no-roofline --no-cpu-baseline --no-graph > gpurun_out/ncu_bench.log 2>&1
python tools/summarize_launches.py gpurun_out/launches.csv bridge_update_kernel > gpurun_out/launches.md 2>&1; head -45 gpurun_out/launches.md
