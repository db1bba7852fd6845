#!/bin/bash
set -u
mkdir -p gpurun_out
echo "=== pytest gpu"
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_c.log 2>&1; tail -5 gpurun_out/pytest_c.log
echo "=== probe"
timeout 900 python tools/probe_dot.py 2>&1 | tail -60
