#!/bin/bash
set -u
mkdir -p gpurun_out
echo "=== pytest gpu"
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_h.log 2>&1; tail -12 gpurun_out/pytest_h.log
echo "=== bench"
timeout 1200 python bench.py > gpurun_out/bench_h.json 2> gpurun_out/bench_h.err
echo "bench exit $?"; python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_h.json"))
print(json.dumps({k:d[k] for k in ("value","ms_per_step","gpu_launches","clocks","other_configs")}, indent=1))
print(d["roofline"]["frac"], d["roofline"]["achieved"], d["e2e"]["value"], d["cpu_baseline"]["value"])
PY
tail -5 gpurun_out/bench_h.err
