#!/bin/bash
# full GPU suite + a short bench line (+ optional launch list: tools/gpu_tests.sh launches)
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -x -q -s 2>&1 | grep -v "Warning\|@custom\|@torch\|warnings.warn\|^$\|INFO\|SUCCESS\|WARNING\|^tests/test_dropin" | tail -60 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log | tail -${TAILN:-40}
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/bench_short.json 2>gpurun_out/bench_short.err; cat gpurun_out/bench_short.json | cut -c1-260; tail -2 gpurun_out/bench_short.err
if [ "$1" = "launches" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 620 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 0 --no-roofline --no-cpu-baseline --no-extra --no-graph > gpurun_out/ncu_bench.log 2>&1
  python tools/summarize_launches.py gpurun_out/launches.csv head_bridge_kernel > gpurun_out/launches.md 2>&1; head -50 gpurun_out/launches.md
fi
