#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_dist.py -m gpu -q > gpurun_out/pytest_2b.log 2>&1; tail -40 gpurun_out/pytest_2b.log | cut -c1-250
