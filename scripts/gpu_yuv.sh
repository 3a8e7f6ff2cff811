#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_color.py tests/test_gpu_camera_isp.py -m gpu -q -x > gpurun_out/pytest_yuv.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_yuv.log | cut -c1-250
python scripts/yuv_bench.py 2>&1 | tail -4
