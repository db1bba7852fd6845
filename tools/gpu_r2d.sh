#!/bin/bash
# N-GPU session (N = $1): NCCL worker + weak-scaling bench
set -u
N=${1:-4}
mkdir -p gpurun_out
nvidia-smi -L | head -8
echo "=== nccl parity worker ($N ranks)"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    tests/dist_worker_gpu.py > gpurun_out/worker_n$N.log 2>&1
echo "worker exit $?"; grep -E "iters|Error|error|assert" gpurun_out/worker_n$N.log | head -8 | cut -c1-330; grep -c WORKER_OK gpurun_out/worker_n$N.log
echo "=== bench N=$N"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 \
    bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_pk_n$N.json 2> gpurun_out/bench_pk_n$N.err
python - $N <<'PY'
import json,sys
N=sys.argv[1]
lines=[l for l in open(f'gpurun_out/bench_pk_n{N}.json') if l.startswith('{')]
if lines:
    d=json.loads(lines[-1]); print(round(d['value'],1), 'global it/s', round(d['config']['global_iterations_per_s'],1), 'e2e', round(d['e2e']['value'],1), 'spmv ms', round(d['roofline']['avg_launch_ms'],3), d['roofline']['iteration']['kernel_ms_share_rank0'], d['config']['partition'].get('rank0_format'), d['clocks'])
else:
    print(open(f'gpurun_out/bench_pk_n{N}.err').read()[-1500:])
PY
