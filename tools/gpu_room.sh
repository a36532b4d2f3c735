#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_room_gpu.py tests/test_dropin_gpu.py "tests/test_engine_gpu.py::test_denoise_room_entry_point_end_to_end" -m gpu -x -q -s 2>&1 | grep -v Warning | grep -v "@custom\|@torch\|warnings.warn" | tail -60
