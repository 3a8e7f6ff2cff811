#!/bin/bash
# round 2, first GPU call: new parity tests + measured error histogram + baseline bench lines
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
nproc
python scripts/error_histogram.py > gpurun_out/r02_error_histogram.txt 2> gpurun_out/error_histogram.err; echo "hist rc=$?"
python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_pipeline.py tests/test_gpu_camera_isp.py -m gpu -q -s -k "vs_c_oracle or pipeline or lookahead or graphed" > gpurun_out/pytest_new.log 2>&1; echo "pytest rc=$?"
grep -E "fullsize|passed|failed|Error|assert" gpurun_out/pytest_new.log | tail -40
for w in cfg1 cfg3 cfg1_16; do
  python bench.py --workload $w --steps 40 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/r2a_bench_$w.json 2> gpurun_out/r2a_bench_$w.err; echo "bench $w rc=$?"
done
