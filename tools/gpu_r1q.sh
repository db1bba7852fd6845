#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_spmv_sell_pk|k_pcg_update_r|k_pcg_direction_x" -s 9 -c 3 -o gpurun_out/ncu_pcg_ro python tools/ncu_pcg.py > gpurun_out/ncu_pcg_ro.log 2>&1; tail -3 gpurun_out/ncu_pcg_ro.log
