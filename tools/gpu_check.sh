#!/bin/bash
# One GPU-box visit: parity tests, bench line, ncu launch list of 3 network evaluations (shares only).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-graph > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -2 gpurun_out/bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 639 -c 639 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 1 --warmup 0 --no-roofline --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
python tools/summarize_launches.py gpurun_out/launches.csv bridge_update_kernel > gpurun_out/launches.md 2>&1; head -40 gpurun_out/launches.md
