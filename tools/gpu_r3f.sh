#!/bin/bash
# N-GPU session (N = $1) with the parameter-block dictionary kernels: NCCL parity worker, dist tests, weak-scaling bench (both arms)
set -u
N=${1:-2}
mkdir -p gpurun_out
echo "=== nccl parity worker ($N ranks)"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    tests/dist_worker_gpu.py > gpurun_out/worker_3f_n$N.log 2>&1
echo "worker exit $?"; grep -E "iters|Error|error|assert" gpurun_out/worker_3f_n$N.log | head -8 | cut -c1-330; grep -c WORKER_OK gpurun_out/worker_3f_n$N.log
if [ "$N" = "2" ]; then
echo "=== dist pytest"
timeout 400 python -m pytest tests/test_gpu_dist.py -m gpu -q > gpurun_out/pytest_3f.log 2>&1; tail -3 gpurun_out/pytest_3f.log
fi
echo "=== bench N=$N"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 \
    bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_3f_n$N.json 2> gpurun_out/bench_3f_n$N.err
python - $N <<'PY'
import json,sys
N=sys.argv[1]
lines=[l for l in open(f'gpurun_out/bench_3f_n{N}.json') if l.startswith('{')]
if lines:
    d=json.loads(lines[-1]); print(round(d['value'],1), 'global it/s', round(d['config']['global_iterations_per_s'],1), 'e2e', round(d['e2e']['value'],1), 'spmv ms', round(d['roofline']['avg_launch_ms'],3), d['roofline']['iteration']['kernel_ms_share_rank0'], d['config']['partition'].get('rank0_format'), d['clocks'])
else:
    print(open(f'gpurun_out/bench_3f_n{N}.err').read()[-1500:])
PY
