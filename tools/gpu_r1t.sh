#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_spmv_sell_pk" -s 4 -c 1 -o gpurun_out/ncu_pcg_blk python tools/ncu_pcg.py spmv.pkvariant=9 spmv.pkdotvariant=9 spmv.pk_threads=512 spmv.pk_slices_per_cta=128 > gpurun_out/ncu_pcg_blk.log 2>&1; tail -2 gpurun_out/ncu_pcg_blk.log
