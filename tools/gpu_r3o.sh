#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "dictionary or pcg_forms or minres or gmres" > gpurun_out/pytest_3o.log 2>&1; tail -2 gpurun_out/pytest_3o.log
timeout 600 python tools/probe_cdict.py > gpurun_out/probe_cdict2.log 2>&1; cut -c1-200 gpurun_out/probe_cdict2.log
