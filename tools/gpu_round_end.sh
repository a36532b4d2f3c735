#!/bin/bash
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python tools/bench_gemm.py > gpurun_out/bench_gemm_r2.md 2>&1; tail -22 gpurun_out/bench_gemm_r2.md
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2200 -c 800 --csv --log-file gpurun_out/launches_pvdl.csv \
  python tools/bench_pvdl.py 32 3 0 nograph > gpurun_out/ncu_pvdl.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_pvdl.csv head_bridge_kernel > gpurun_out/launches_pvdl.md 2>&1; head -44 gpurun_out/launches_pvdl.md
timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench exit $?"; tail -2 gpurun_out/bench_full.err; cut -c1-700 gpurun_out/bench_full.json
