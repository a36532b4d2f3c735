#!/bin/bash
# full GPU suite + a short bench line
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -x -q -s 2>&1 | grep -v "Warning\|@custom\|@torch\|warnings.warn\|^$\|INFO\|SUCCESS" | tail -${1:-60} > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log | tail -70
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/bench_short.json 2>gpurun_out/bench_short.err; cat gpurun_out/bench_short.json | cut -c1-300
