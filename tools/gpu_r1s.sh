#!/bin/bash
set -u
mkdir -p gpurun_out
echo "=== pytest"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "block_order or value_dictionary or pcg_forms" > gpurun_out/pytest_s.log 2>&1; tail -12 gpurun_out/pytest_s.log
echo "=== probe"
timeout 1200 python tools/probe_pk4.py > gpurun_out/probe_pk4.log 2>&1; cat gpurun_out/probe_pk4.log | cut -c1-300
