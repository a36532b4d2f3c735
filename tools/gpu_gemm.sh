#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gemm_gpu.py -x -q -k "persistent_cluster" > gpurun_out/pytest_gemm.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gemm.log
tail -15 gpurun_out/pytest_gemm.log
timeout 300 python tools/bench_gemm.py > gpurun_out/bench_gemm.log 2>&1; grep -v experiments gpurun_out/bench_gemm.log | cut -c1-170
