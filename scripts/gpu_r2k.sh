#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_camera_isp.py tests/test_gpu_golden.py tests/test_gpu_fullsize.py tests/test_gpu_bilinear_isp.py tests/test_gpu_distributed.py tests/test_gpu_resize_isp.py tests/test_gpu_pipeline.py tests/test_gpu_rig.py tests/test_gpu_color.py -m gpu -q -x > gpurun_out/pytest_r2k.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_r2k.log
python scripts/ids_bench.py 2>&1 | tail -3
for w in cfg2 cfg3 cfg1_16; do python bench.py --workload $w --steps 200 --no-cpu-baseline --no-e2e --configs 0 > gpurun_out/r2k_bench.json 2>/dev/null; python - <<PY
import json
d = json.load(open('gpurun_out/r2k_bench.json'))
print('$w step %.1f Gpx/s  sustained %.1f  kernel alone %.4f ms = %.3f' % (d['value'], d['sustained']['value'], d['roofline']['kernel_ms'], d['roofline']['frac']))
PY
done
