#!/bin/bash
# 2-GPU session: dist tests on one GPU (local transport), NCCL parity worker, weak-scaling bench at N=2 and N=1 on the same box
set -u
mkdir -p gpurun_out
nvidia-smi -L
echo "=== dist tests"
timeout 900 python -m pytest tests/test_gpu_dist.py -m gpu -q -x > gpurun_out/pytest_v.log 2>&1; tail -5 gpurun_out/pytest_v.log
echo "=== nccl parity worker (2 ranks)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
    tests/dist_worker_gpu.py > gpurun_out/worker2.log 2>&1
echo "worker exit $?"; grep -E "WORKER_OK|iters|Error|error|assert" gpurun_out/worker2.log | head -20; tail -3 gpurun_out/worker2.log
echo "=== bench N=2"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 \
    bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_v_n2.json 2> gpurun_out/bench_v_n2.err
echo "bench2 exit $?"; cat gpurun_out/bench_v_n2.json | cut -c1-3000; tail -3 gpurun_out/bench_v_n2.err
echo "=== bench N=1"
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu --no-small > gpurun_out/bench_v_n1.json 2> gpurun_out/bench_v_n1.err
echo "bench1 exit $?"; cat gpurun_out/bench_v_n1.json | cut -c1-3000; tail -3 gpurun_out/bench_v_n1.err
