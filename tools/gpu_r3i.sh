#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_pcg_direction_x|k_pcg_update_r" -s 6 -c 2 -f -o gpurun_out/ncu_dir_idx python tools/ncu_pcg.py > gpurun_out/ncu_dir_idx.log 2>&1; tail -2 gpurun_out/ncu_dir_idx.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_pcg_direction_x|k_pcg_update_r" -s 6 -c 2 -f -o gpurun_out/ncu_dir_vec python tools/ncu_pcg.py pcg.diag_index=0 > gpurun_out/ncu_dir_vec.log 2>&1; tail -2 gpurun_out/ncu_dir_vec.log
