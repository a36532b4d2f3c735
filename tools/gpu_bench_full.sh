#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_dropin_gpu.py -m gpu -x -q -s 2>&1 | grep -v Warning | grep -v "@custom\|@torch" | tail -15
timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench exit $?"; tail -3 gpurun_out/bench_full.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_full.json').read().strip().splitlines()[-1])
print(json.dumps({k:d[k] for k in ('value','e2e','parity','extra','cpu_baseline')}, indent=1))
PY
