#!/bin/bash
# ncu evidence for the parameter-block dictionary kernel: full capture of one PCG iteration's kernels + launch list of the bench
set -u
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_spmv_sell_pk|k_pcg_update_r|k_pcg_direction_x" -s 9 -c 3 -f -o gpurun_out/ncu_pcg_cd python tools/ncu_pcg.py > gpurun_out/ncu_pcg_cd.log 2>&1; tail -3 gpurun_out/ncu_pcg_cd.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_3d.csv python bench.py --steps 1 --warmup 1 --iters 20 --no-cpu --no-host-matrix --no-small --no-general > gpurun_out/launches_3d.log 2>&1
tail -2 gpurun_out/launches_3d.log | cut -c1-300
ls -la gpurun_out/ncu_pcg_cd.ncu-rep gpurun_out/launches_3d.csv
