#!/bin/bash
# round-2 first GPU call: R-GPU (the unmodified reference on the B200) goldens + noise floor, then the state of the suite/bench
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi_r2.txt
timeout 1500 python -m oracle.gen_golden_rgpu > gpurun_out/rgpu.log 2>&1; echo "rgpu exit $?"
tail -8 gpurun_out/rgpu.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r2_start.json 2> gpurun_out/bench_r2_start.err; echo "bench exit $?"
cat gpurun_out/bench_r2_start.json
