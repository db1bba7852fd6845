#!/bin/bash
set -u
mkdir -p gpurun_out
echo "=== pytest diag/pcg"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "diagonal or pcg_forms or dictionary" > gpurun_out/pytest_3h.log 2>&1; tail -3 gpurun_out/pytest_3h.log
for o in "--opt pcg.diag_index=0" "" "--opt vec.unroll2_update=1" "--opt vec.unroll2=0"; do
  echo "--- $o"
  timeout 600 python bench.py --no-cpu --no-small --no-host-matrix --no-general --no-solvers --steps 2 --warmup 2 $o 2> gpurun_out/bench_g.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())
ms=d['ms_per_step']; sh=d['roofline']['iteration']['kernel_ms_share']
print(round(d['value'],1),'it/s  e2e',round(d['e2e']['value'],1),' spmv ms',round(d['roofline']['avg_launch_ms'],3),' update ms', round(sh['pcg.update']*ms/200,4), ' direction ms', round(sh['pcg.direction']*ms/200,4), d['clocks']['sm_mhz'])
"
  tail -2 gpurun_out/bench_g.err
done
