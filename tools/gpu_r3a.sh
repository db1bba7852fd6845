#!/bin/bash
# x window in shared memory: parity tests, then the variant probe on C2
set -u
mkdir -p gpurun_out
echo "=== pytest gpu all"
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_3a.log 2>&1; tail -5 gpurun_out/pytest_3a.log
echo "=== probe"
timeout 900 python tools/probe_xw.py > gpurun_out/probe_xw.log 2>&1; cat gpurun_out/probe_xw.log | cut -c1-250
