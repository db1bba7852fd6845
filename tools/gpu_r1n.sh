#!/bin/bash
# session n: packed stream -- tests, probe of the dot forms, bench with the new PCG kernels
set -u
mkdir -p gpurun_out
echo "=== pytest gpu all"
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_n.log 2>&1; tail -8 gpurun_out/pytest_n.log
echo "=== probe"
timeout 900 python tools/probe_pk.py > gpurun_out/probe_pk.log 2>&1; grep -i "form\|unfused\|plain v0\"\|dot v[0169]\"\|error\|Trace" gpurun_out/probe_pk.log | cut -c1-200
echo "=== bench variants"
for o in "" "--opt pcg.fused_dot=0" "--opt pcg.defer_x=0" "--opt spmv.pkdotvariant=1 --opt spmv.pkdotform=0"; do
  echo "--- $o"
  timeout 600 python bench.py --no-cpu --no-small --no-host-matrix --steps 2 --warmup 2 $o 2> gpurun_out/bench_n.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(round(d['value'],1),'it/s  e2e',round(d['e2e']['value'],1),' spmv ms',round(d['roofline']['avg_launch_ms'],3),' frac',round(d['roofline']['frac'],3),' iter frac',round(d['roofline']['iteration']['frac'],3), d['roofline']['iteration']['kernel_ms_share'], d['clocks']['sm_mhz'], (d.get('general_12B_stream') or {}).get('iterations_per_s'))
"
done
