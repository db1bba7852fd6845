#!/bin/bash
# GPU session B (round 1): parity suite incl. the partitioned path on the local transport,
# memcheck of the new kernels, bench with the two-level fused-dot reduction, ncu launch list + capture.
set -u
mkdir -p gpurun_out
echo "=== pytest gpu"
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -25

echo "=== memcheck (partitioned path + fused dot)"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_dist.py tests/test_gpu_parity.py -m gpu -x -q \
    -k "golden and 3 or empty_rank or fused_spmv_dot" > gpurun_out/memcheck_b.log 2>&1
echo "memcheck exit $?"; tail -5 gpurun_out/memcheck_b.log

echo "=== bench"
timeout 1200 python bench.py > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err
echo "bench exit $?"; cat gpurun_out/bench_b.json; tail -5 gpurun_out/bench_b.err

echo "=== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches_b.csv \
    python bench.py --steps 1 --warmup 1 --iters 20 --no-cpu --no-host-matrix > gpurun_out/ncu_list_b.log 2>&1
echo "ncu list exit $?"

echo "=== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_spmv_sell -s 3 -c 2 -f -o gpurun_out/prof_spmv_b \
    python bench.py --steps 1 --warmup 1 --iters 5 --no-cpu --no-host-matrix > gpurun_out/ncu_full_b.log 2>&1
echo "ncu full exit $?"
ncu -i gpurun_out/prof_spmv_b.ncu-rep --page raw --csv > gpurun_out/prof_spmv_b.raw.csv 2>/dev/null
ls -la gpurun_out | tail -12
