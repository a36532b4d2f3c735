#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py -x -q -k "halo" > gpurun_out/pytest_halo.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_halo.log
tail -8 gpurun_out/pytest_halo.log
timeout 300 python tools/bench_halo.py > gpurun_out/bench_halo.log 2>&1; cat gpurun_out/bench_halo.log
