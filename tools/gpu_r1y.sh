#!/bin/bash
# 2-GPU session: dist tests (local transport), NCCL worker incl. GMRES, launch list of the PCG kernels
set -u
mkdir -p gpurun_out
echo "=== dist tests"
timeout 900 python -m pytest tests/test_gpu_dist.py -m gpu -q -x > gpurun_out/pytest_y.log 2>&1; tail -15 gpurun_out/pytest_y.log
echo "=== nccl parity worker (2 ranks)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
    tests/dist_worker_gpu.py > gpurun_out/worker2_y.log 2>&1
echo "worker exit $?"; grep -E "WORKER_OK|iters|Error|error|assert" gpurun_out/worker2_y.log | head -20; tail -3 gpurun_out/worker2_y.log
echo "=== launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_spmv_sell_pk|k_pcg_|k_sum_warp_partials|k_gather_perm|k_scatter_perm|k_stream" -c 300 --csv --log-file gpurun_out/launches_y.csv python bench.py --steps 1 --warmup 1 --iters 20 --no-cpu --no-host-matrix --no-small --no-general > gpurun_out/launches_y.log 2>&1
tail -1 gpurun_out/launches_y.log | cut -c1-200
