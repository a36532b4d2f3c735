#!/bin/bash
# launch list of the bench engine (no graph), evaluations 8.. of the first sample() call
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1600 -c 600 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 1 --warmup 0 --no-roofline --no-cpu-baseline --no-extra --no-graph > gpurun_out/ncu_bench.log 2>&1
python tools/summarize_launches.py gpurun_out/launches.csv head_bridge_kernel > gpurun_out/launches.md 2>&1; head -50 gpurun_out/launches.md
