#!/bin/bash
# Second GPU session: full parity suite, extended SpMV sweep, ncu of the many-CTA configuration.
set -u
mkdir -p gpurun_out
echo "=== pytest gpu"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25

echo "=== sweep"
timeout 1200 python tools/sweep_spmv.py --out gpurun_out/sweep2.jsonl > gpurun_out/sweep2.log 2>&1
echo "sweep exit $?"; tail -150 gpurun_out/sweep2.log

echo "=== bench"
timeout 1200 python bench.py > gpurun_out/bench2.json 2> gpurun_out/bench2.err
echo "bench exit $?"; cat gpurun_out/bench2.json; tail -5 gpurun_out/bench2.err

echo "=== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_spmv_sell -s 3 -c 2 -f -o gpurun_out/prof2_spmv \
    python bench.py --steps 1 --warmup 1 --iters 5 --no-cpu --no-host-matrix > gpurun_out/ncu2_full.log 2>&1
echo "ncu full exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pcg_ -s 2 -c 2 -f -o gpurun_out/prof2_vec \
    python bench.py --steps 1 --warmup 1 --iters 5 --no-cpu --no-host-matrix > gpurun_out/ncu2_vec.log 2>&1
echo "ncu vec exit $?"
ls -la gpurun_out | tail -12
