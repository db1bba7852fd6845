#!/bin/bash
# final single-GPU record: all GPU tests, smoke, both bench arms with defaults
set -u
mkdir -p gpurun_out
echo "=== pytest gpu all"
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_3p.log 2>&1; tail -5 gpurun_out/pytest_3p.log
echo "=== smoke"
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
echo "=== bench ref"
timeout 600 python bench.py --impl reference > gpurun_out/bench_3p_ref.json 2> gpurun_out/bench_3p_ref.err; tail -c 300 gpurun_out/bench_3p_ref.json
echo "=== bench"
timeout 900 python bench.py > gpurun_out/bench_3p.json 2> gpurun_out/bench_3p.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_3p.json'))
print(round(d['value'],1),'it/s  e2e',round(d['e2e']['value'],1),' spmv ms',round(d['roofline']['avg_launch_ms'],3),' frac',round(d['roofline']['frac'],3),' iter frac',round(d['roofline']['iteration']['frac'],3), d['roofline']['iteration']['kernel_ms_share'], d['clocks'], d['general_12B_stream'], d['other_configs']['C1_p1_n64']['graph'])
PY
tail -3 gpurun_out/bench_3p.err
