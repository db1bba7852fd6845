#!/bin/bash
set -u
mkdir -p gpurun_out
echo "=== pytest gpu"
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_i.log 2>&1; tail -25 gpurun_out/pytest_i.log
echo "=== bench"
timeout 1200 python bench.py > gpurun_out/bench_i.json 2> gpurun_out/bench_i.err
echo "bench exit $?"; python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_i.json"))
print(json.dumps({k:d[k] for k in ("value","ms_per_step","gpu_launches","clocks","other_configs")}, indent=1))
print(d["roofline"]["frac"], d["roofline"]["achieved"], d["e2e"]["value"], d["cpu_baseline"]["value"])
PY
tail -5 gpurun_out/bench_i.err
