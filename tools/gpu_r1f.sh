#!/bin/bash
set -u
mkdir -p gpurun_out
echo "=== pytest gpu"
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_f.log 2>&1; tail -5 gpurun_out/pytest_f.log
echo "=== probe (compressed)"
timeout 900 python tools/probe_dot.py 2>&1 | grep -E "plain v0\"|plain v[268]|mode2 v[0678]\"|unfused|mode2 v8" | head -30
cp gpurun_out/probe_dot.json gpurun_out/probe_dot_c16.json
echo "=== bench"
timeout 900 python bench.py --no-cpu > gpurun_out/bench_f.json 2> gpurun_out/bench_f.err
echo "bench exit $?"; cat gpurun_out/bench_f.json | cut -c1-1500; tail -5 gpurun_out/bench_f.err
