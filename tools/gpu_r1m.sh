#!/bin/bash
# session m: dictionary-packed SpMV stream -- parity tests, probe, ncu capture
set -u
mkdir -p gpurun_out
echo "=== pytest value dictionary"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "value_dictionary or layout_and_launch" > gpurun_out/pytest_m.log 2>&1; tail -15 gpurun_out/pytest_m.log
echo "=== probe"
timeout 900 python tools/probe_pk.py > gpurun_out/probe_pk.log 2>&1; cat gpurun_out/probe_pk.log | cut -c1-200
echo "=== ncu"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_spmv_sell_pk -s 2 -c 2 -o gpurun_out/ncu_pk python tools/ncu_pk.py > gpurun_out/ncu_pk.log 2>&1; tail -3 gpurun_out/ncu_pk.log
ls -la gpurun_out/*.ncu-rep
