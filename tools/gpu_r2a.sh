#!/bin/bash
# 2-GPU session: dist tests, NCCL worker, bench N=2 / N=1, PCIe probe
set -u
mkdir -p gpurun_out
echo "=== all gpu tests"
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_2a.log 2>&1; tail -5 gpurun_out/pytest_2a.log
echo "=== nccl parity worker (2 ranks)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
    tests/dist_worker_gpu.py > gpurun_out/worker2_2a.log 2>&1
echo "worker exit $?"; grep -E "WORKER_OK|iters|Error|error|assert" gpurun_out/worker2_2a.log | head -20 | cut -c1-300; tail -3 gpurun_out/worker2_2a.log | cut -c1-300
echo "=== bench N=2"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 \
    bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_2a_n2.json 2> gpurun_out/bench_2a_n2.err
echo "bench2 exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_2a_n2.json')); print(d['value'], d['e2e']['value'], d['roofline']['avg_launch_ms'], d['roofline']['iteration']['kernel_ms_share_rank0'], d['config']['partition']['rank0_format'])"; tail -3 gpurun_out/bench_2a_n2.err
echo "=== bench N=2 dist.reorder=0"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 \
    bench.py --gpus 2 --steps 3 --warmup 3 --opt dist.reorder=0 --opt sell.block_order=0 > gpurun_out/bench_2a_n2b.json 2> gpurun_out/bench_2a_n2b.err
python -c "
import json; d=json.load(open('gpurun_out/bench_2a_n2b.json')); print(d['value'], d['e2e']['value'], d['roofline']['avg_launch_ms'], d['roofline']['iteration']['kernel_ms_share_rank0'])"
echo "=== bench N=1"
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu --no-small > gpurun_out/bench_2a_n1.json 2> gpurun_out/bench_2a_n1.err
python -c "
import json; d=json.load(open('gpurun_out/bench_2a_n1.json')); print(d['value'], d['e2e']['value'], d['roofline']['avg_launch_ms'])"
exit 0
timeout 300 python tools/probe_pcie.py
