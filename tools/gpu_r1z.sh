#!/bin/bash
set -u
mkdir -p gpurun_out
echo "=== pytest gpu all"
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_z.log 2>&1; tail -12 gpurun_out/pytest_z.log
echo "=== bench variants"
for o in "" "--opt sell.pack_bits=0"; do
  echo "--- $o"
  timeout 600 python bench.py --no-cpu --no-small --no-host-matrix --no-general --steps 2 --warmup 2 $o 2> gpurun_out/bench_z.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(round(d['value'],1),'it/s  e2e',round(d['e2e']['value'],1),' spmv ms',round(d['roofline']['avg_launch_ms'],3),' frac',round(d['roofline']['frac'],3),' iter frac',round(d['roofline']['iteration']['frac'],3), d['roofline']['iteration']['kernel_ms_share'], d['clocks']['sm_mhz'], d['config']['spmv_kernel'])
"
  tail -2 gpurun_out/bench_z.err
done
