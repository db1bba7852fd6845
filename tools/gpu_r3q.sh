#!/bin/bash
# last session of the round (2 GPUs): whole GPU suite, smoke, NCCL parity worker, N=2 bench
set -u
mkdir -p gpurun_out
echo "=== pytest gpu all"
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_3q.log 2>&1; tail -3 gpurun_out/pytest_3q.log
echo "=== smoke"
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
echo "=== nccl parity worker (2 ranks)"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
    tests/dist_worker_gpu.py > gpurun_out/worker_3q_n2.log 2>&1
echo "worker exit $?"; grep -E "iters|Error|error|assert" gpurun_out/worker_3q_n2.log | head -8 | cut -c1-200
echo "=== bench N=2"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 \
    bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_3q_n2.json 2> gpurun_out/bench_3q_n2.err
python - <<'PY'
import json
lines=[l for l in open('gpurun_out/bench_3q_n2.json') if l.startswith('{')]
if lines:
    d=json.loads(lines[-1]); print(round(d['value'],1), 'global it/s', round(d['config']['global_iterations_per_s'],1), 'e2e', round(d['e2e']['value'],1), 'spmv ms', round(d['roofline']['avg_launch_ms'],3), d['roofline']['iteration']['kernel_ms_share_rank0'], d['config']['partition'].get('rank0_format'), d['roofline']['iteration'].get('vector_passes'))
else:
    print(open('gpurun_out/bench_3q_n2.err').read()[-1500:])
PY
