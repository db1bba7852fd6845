#!/bin/bash
set -u
mkdir -p gpurun_out
echo "=== pytest gpu"
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_j.log 2>&1; tail -25 gpurun_out/pytest_j.log
echo "=== smoke"
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
echo "=== bench"
timeout 1200 python bench.py > gpurun_out/bench_j.json 2> gpurun_out/bench_j.err
echo "bench exit $?"; python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_j.json"))
print(json.dumps({k:d.get(k) for k in ("value","ms_per_step","gpu_launches","clocks","other_configs")}, indent=1))
print(d["roofline"], d["e2e"], d["cpu_baseline"])
PY
tail -5 gpurun_out/bench_j.err
