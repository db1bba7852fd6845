#!/bin/bash
# 2-GPU session: NCCL worker with the peer-to-peer halo, bench N=2 with / without it
set -u
mkdir -p gpurun_out
echo "=== nccl parity worker (2 ranks)"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
    tests/dist_worker_gpu.py > gpurun_out/worker2_2c.log 2>&1
echo "worker exit $?"; grep -E "WORKER_OK|iters|Error|error|assert" gpurun_out/worker2_2c.log | head -20 | cut -c1-330
echo "=== 2-rank pytest"
timeout 300 python -m pytest tests/test_gpu_dist.py -m gpu -q -k two_ranks > gpurun_out/pytest_2c.log 2>&1; tail -3 gpurun_out/pytest_2c.log
for o in "" "--opt dist.p2p=0 --opt dist.p2p_reduce=0" "--opt dist.p2p_reduce=0"; do
echo "=== bench N=2 $o"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 \
    bench.py --gpus 2 --steps 3 --warmup 3 $o > gpurun_out/bench_2c_n2.json 2> gpurun_out/bench_2c_n2.err
python - <<'PY'
import json
lines=[l for l in open('gpurun_out/bench_2c_n2.json') if l.startswith('{')]
if lines:
    d=json.loads(lines[-1]); print(round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'spmv ms', round(d['roofline']['avg_launch_ms'],3), d['roofline']['iteration']['kernel_ms_share_rank0'], d['config']['partition'].get('rank0_format'), d['clocks']['sm_mhz'])
else:
    print(open('gpurun_out/bench_2c_n2.err').read()[-1500:])
PY
done
