#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py -x -q > gpurun_out/pytest_gemm.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gemm.log
tail -15 gpurun_out/pytest_gemm.log
timeout 300 python tools/bench_gemm.py > gpurun_out/bench_gemm.log 2>&1; cat gpurun_out/bench_gemm.log
