#!/bin/bash
# GPU session A (round 1): parity suite, bench, ncu launch list + full captures exported as CSV.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.max.mem,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/host.txt; free -g >> gpurun_out/host.txt

echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5

echo "=== pytest gpu"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25

echo "=== bench"
timeout 1200 python bench.py > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err
echo "bench exit $?"; cat gpurun_out/bench_a.json; tail -5 gpurun_out/bench_a.err

echo "=== bench reference arm"
timeout 600 python bench.py --impl reference > gpurun_out/bench_a_ref.json 2> gpurun_out/bench_a_ref.err
echo "ref exit $?"; cat gpurun_out/bench_a_ref.json; tail -3 gpurun_out/bench_a_ref.err

echo "=== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches.csv \
    python bench.py --steps 1 --warmup 1 --iters 20 --no-cpu --no-host-matrix > gpurun_out/ncu_list.log 2>&1
echo "ncu list exit $?"

echo "=== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_spmv_sell -s 3 -c 2 -f -o gpurun_out/prof_spmv \
    python bench.py --steps 1 --warmup 1 --iters 5 --no-cpu --no-host-matrix > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pcg_ -s 2 -c 2 -f -o gpurun_out/prof_vec \
    python bench.py --steps 1 --warmup 1 --iters 5 --no-cpu --no-host-matrix > gpurun_out/ncu_vec.log 2>&1
echo "ncu vec exit $?"
for f in prof_spmv prof_vec; do
  ncu -i gpurun_out/$f.ncu-rep --page raw --csv > gpurun_out/$f.raw.csv 2>/dev/null
done
ls -la gpurun_out | tail -20
