#!/bin/bash
# 8-GPU session: NCCL parity worker at 8 ranks, weak-scaling bench at N=8 and N=4.
set -u
mkdir -p gpurun_out
nvidia-smi -L | wc -l
nvidia-smi topo -m > gpurun_out/topo8.txt 2>&1
echo "=== nccl parity worker (8 ranks)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29543 \
    tests/dist_worker_gpu.py > gpurun_out/worker8.log 2>&1
echo "worker exit $?"; grep -E "WORKER_OK|iters|Error|error|assert" gpurun_out/worker8.log | sort | uniq -c | head -20
echo "=== bench N=8"
NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 \
    bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err
echo "bench8 exit $?"; cat gpurun_out/bench_n8.json; tail -5 gpurun_out/bench_n8.err
echo "=== bench N=4"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29545 \
    bench.py --gpus 4 --steps 3 --warmup 3 > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err
echo "bench4 exit $?"; cat gpurun_out/bench_n4.json; tail -5 gpurun_out/bench_n4.err
