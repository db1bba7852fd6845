#!/bin/bash
set -u
mkdir -p gpurun_out
echo "=== pytest"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "value_dictionary or pcg_forms" > gpurun_out/pytest_r.log 2>&1; tail -12 gpurun_out/pytest_r.log
echo "=== probe"
timeout 900 python tools/probe_pk3.py > gpurun_out/probe_pk3.log 2>&1; tail -5 gpurun_out/probe_pk3.log | cut -c1-300
