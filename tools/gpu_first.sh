#!/bin/bash
# First GPU session: parity tests, SpMV sweep, bench, ncu launch list + full capture of the SpMV kernel.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.max.mem,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/host.txt; free -g >> gpurun_out/host.txt

echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5

echo "=== pytest gpu (small)"
timeout 900 python -m pytest tests -m gpu -x -q -k "not full_size" 2>&1 | tail -25

echo "=== memcheck (subset)"
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q \
    -k "spmv_bit_identical or reference_kats or vector_kernels" > gpurun_out/memcheck.log 2>&1
echo "memcheck exit $?"; tail -5 gpurun_out/memcheck.log

echo "=== sweep"
timeout 900 python tools/sweep_spmv.py --out gpurun_out/sweep_spmv.jsonl > gpurun_out/sweep.log 2>&1
echo "sweep exit $?"; tail -70 gpurun_out/sweep.log

echo "=== full size test"
timeout 900 python -m pytest tests -m gpu -x -q -k "full_size" 2>&1 | tail -15

echo "=== bench"
timeout 1200 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err

echo "=== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --iters 20 --no-cpu --no-host-matrix > gpurun_out/ncu_list.log 2>&1
echo "ncu list exit $?"

echo "=== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_spmv_sell -s 3 -c 2 -f -o gpurun_out/prof_spmv \
    python bench.py --steps 1 --warmup 1 --iters 5 --no-cpu --no-host-matrix > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_stream -s 6 -c 2 -f -o gpurun_out/prof_vec \
    python bench.py --steps 1 --warmup 1 --iters 5 --no-cpu --no-host-matrix > gpurun_out/ncu_vec.log 2>&1
echo "ncu vec exit $?"
ls -la gpurun_out
