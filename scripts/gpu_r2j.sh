#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_camera_isp.py tests/test_gpu_golden.py tests/test_gpu_fullsize.py tests/test_gpu_bilinear_isp.py tests/test_gpu_distributed.py tests/test_gpu_resize_isp.py -m gpu -q -x > gpurun_out/pytest_r2j.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_r2j.log
python scripts/ids_bench.py 2>&1 | tail -3
for i in 1 2; do python bench.py --steps 400 --no-cpu-baseline --no-e2e --configs 0 > gpurun_out/r2j_bench.json 2>/dev/null; python - <<'PY'
import json
d = json.load(open('gpurun_out/r2j_bench.json'))
print('cfg2 step %.1f Gpx/s  kernel alone %.4f ms = %.3f' % (d['value'], d['roofline']['kernel_ms'], d['roofline']['frac']))
PY
done
