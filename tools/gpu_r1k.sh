#!/bin/bash
set -u
mkdir -p gpurun_out
echo "=== pytest gpu ext"
timeout 900 python -m pytest tests/test_gpu_ext.py -m gpu -q -x > gpurun_out/pytest_k.log 2>&1; tail -40 gpurun_out/pytest_k.log
echo "=== pytest gpu all"
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_k_all.log 2>&1; tail -8 gpurun_out/pytest_k_all.log
