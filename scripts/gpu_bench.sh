#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 30 --warmup 5 > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err; echo "bench rc=$?"; cat gpurun_out/bench_cfg2.json; tail -3 gpurun_out/bench_cfg2.err
python bench.py --workload cfg1 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_cfg1.json 2> gpurun_out/bench_cfg1.err; cat gpurun_out/bench_cfg1.json; tail -3 gpurun_out/bench_cfg1.err
python bench.py --workload cfg1_16 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_cfg1_16.json 2> gpurun_out/bench_cfg1_16.err; cat gpurun_out/bench_cfg1_16.json; tail -3 gpurun_out/bench_cfg1_16.err
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_cfg2.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu1.log 2>&1
echo "ncu rc=$?"
