#!/bin/bash
# session o: row-ordered PCG -- tests, bench variants, ncu of the row-ordered kernel
set -u
mkdir -p gpurun_out
echo "=== pytest gpu all"
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_o.log 2>&1; tail -12 gpurun_out/pytest_o.log
echo "=== bench variants"
for o in "" "--opt pcg.reorder=0" "--opt pcg.fused_dot=0" "--opt spmv.pk_threads=256 --opt spmv.pk_slices_per_cta=32" "--opt spmv.pkdotvariant=9 --opt spmv.pkvariant=9 --opt spmv.pk_threads=512 --opt spmv.pk_slices_per_cta=128"; do
  echo "--- $o"
  timeout 600 python bench.py --no-cpu --no-small --no-host-matrix --no-general --steps 2 --warmup 2 $o 2> gpurun_out/bench_o.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(round(d['value'],1),'it/s  e2e',round(d['e2e']['value'],1),' spmv ms',round(d['roofline']['avg_launch_ms'],3),' frac',round(d['roofline']['frac'],3),' iter frac',round(d['roofline']['iteration']['frac'],3), d['roofline']['iteration']['kernel_ms_share'], d['clocks']['sm_mhz'])
"
  tail -2 gpurun_out/bench_o.err
done
