#!/bin/bash
# stand-alone operator check: parity tests of the touched modules, then scripts/ops_bench.py
mkdir -p gpurun_out
python -m pytest ${TESTS:-tests/test_gpu_tonemap.py tests/test_gpu_color.py tests/test_gpu_golden.py tests/test_gpu_bayer.py} -m gpu -q -x > gpurun_out/pytest_ops.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_ops.log | cut -c1-250
python scripts/ops_bench.py ${OPS_FILTER} > gpurun_out/ops_bench.txt 2>gpurun_out/ops_bench.err; echo "ops rc=$?"; cat gpurun_out/ops_bench.txt | cut -c1-160
