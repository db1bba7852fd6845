#!/bin/bash
set -u
mkdir -p gpurun_out
for o in "" "--opt spmv.pkvariant=15 --opt spmv.pkdotvariant=15" "--opt spmv.pktma=1 --opt spmv.pktma_geometry=3 --opt spmv.pktma_slices_per_cta=32" "--opt spmv.pktma=1 --opt spmv.pktma_geometry=8 --opt spmv.pktma_slices_per_cta=8" ""; do
  echo "--- $o"
  timeout 600 python bench.py --no-cpu --no-small --no-host-matrix --no-general --steps 2 --warmup 2 $o 2> gpurun_out/bench_w.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(round(d['value'],1),'it/s  e2e',round(d['e2e']['value'],1),' spmv ms',round(d['roofline']['avg_launch_ms'],3),' frac',round(d['roofline']['frac'],3),' iter frac',round(d['roofline']['iteration']['frac'],3), d['roofline']['iteration']['kernel_ms_share'], d['clocks']['sm_mhz'])
"
  tail -2 gpurun_out/bench_w.err
done
