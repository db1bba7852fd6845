#!/bin/bash
# session l: re-validate the whole tree on a fresh box (all GPU tests, smoke, both bench arms)
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/l_smi.txt 2>&1
echo "=== pytest gpu all"
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_l_all.log 2>&1; tail -8 gpurun_out/pytest_l_all.log
echo "=== smoke"
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
echo "=== bench ref"
timeout 600 python bench.py --impl reference > gpurun_out/bench_l_ref.json 2> gpurun_out/bench_l_ref.err; tail -c 600 gpurun_out/bench_l_ref.json
echo "=== bench"
timeout 900 python bench.py > gpurun_out/bench_l.json 2> gpurun_out/bench_l.err; cat gpurun_out/bench_l.json; tail -5 gpurun_out/bench_l.err
