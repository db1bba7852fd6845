#!/bin/bash
# launch list of the final configuration + a clean default bench run (both arms)
set -u
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_x.csv python bench.py --steps 1 --warmup 1 --iters 20 --no-cpu --no-host-matrix --no-small --no-general > gpurun_out/launches_x.log 2>&1
tail -2 gpurun_out/launches_x.log | cut -c1-300
echo "=== bench ref"
timeout 600 python bench.py --impl reference > gpurun_out/bench_x_ref.json 2> gpurun_out/bench_x_ref.err; tail -c 400 gpurun_out/bench_x_ref.json
echo "=== bench"
timeout 900 python bench.py > gpurun_out/bench_x.json 2> gpurun_out/bench_x.err; cat gpurun_out/bench_x.json; tail -3 gpurun_out/bench_x.err
