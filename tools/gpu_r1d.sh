#!/bin/bash
# 2-GPU session: NCCL parity worker + weak-scaling bench at N=2 (and N=1 on the same box).
set -u
mkdir -p gpurun_out
nvidia-smi -L
echo "=== nccl parity worker (2 ranks)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
    tests/dist_worker_gpu.py > gpurun_out/worker2.log 2>&1
echo "worker exit $?"; grep -E "WORKER_OK|iters|Error|error|assert" gpurun_out/worker2.log | head -20; tail -5 gpurun_out/worker2.log
echo "=== bench N=2"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 \
    bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
echo "bench2 exit $?"; cat gpurun_out/bench_n2.json; tail -5 gpurun_out/bench_n2.err
echo "=== bench N=1"
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo "bench1 exit $?"; cat gpurun_out/bench_n1.json; tail -5 gpurun_out/bench_n1.err
