#!/bin/bash
set -u
mkdir -p gpurun_out
echo "=== pytest gpu (dictionary tests)"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "dictionary or window or pcg_forms" > gpurun_out/pytest_3b.log 2>&1; tail -5 gpurun_out/pytest_3b.log
echo "=== probe"
timeout 600 python tools/probe_cdict.py > gpurun_out/probe_cdict.log 2>&1; cut -c1-250 gpurun_out/probe_cdict.log
