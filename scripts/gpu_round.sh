#!/bin/bash
# One GPU-box visit: parity tests, bench lines, ncu launch list (run under gpurun from the repo root).
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
python bench.py --steps 30 --warmup 5 > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err; echo "bench rc=$?"; cat gpurun_out/bench_cfg2.json; tail -3 gpurun_out/bench_cfg2.err
python bench.py --workload cfg1 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_cfg1.json 2> gpurun_out/bench_cfg1.err; cat gpurun_out/bench_cfg1.json; tail -3 gpurun_out/bench_cfg1.err
python bench.py --workload cfg1_16 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_cfg1_16.json 2> gpurun_out/bench_cfg1_16.err; cat gpurun_out/bench_cfg1_16.json; tail -3 gpurun_out/bench_cfg1_16.err
