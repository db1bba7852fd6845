#!/bin/bash
set -u
mkdir -p gpurun_out
echo "=== pytest gpu (tma + compression)"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "tma or compression" > gpurun_out/pytest_g.log 2>&1; tail -15 gpurun_out/pytest_g.log
echo "=== probe"
timeout 900 python tools/probe_tma.py 2>&1 | tail -60
