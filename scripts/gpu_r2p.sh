#!/bin/bash
# default write policy for an L2-sized Reinhard map scratch (cfg1), single memset node; parity of the touched paths
mkdir -p gpurun_out
python -m pytest tests/test_gpu_reinhard_map16.py tests/test_gpu_camera_isp.py tests/test_gpu_fullsize.py tests/test_gpu_golden.py tests/test_gpu_resize_isp.py -m gpu -q > gpurun_out/pytest_r2p.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_r2p.log
for w in cfg1 cfg3 cfg1_16 cfg2; do python bench.py --workload $w --steps 200 --no-cpu-baseline --no-e2e --configs 0 > gpurun_out/r2p_bench_${w}.json 2>gpurun_out/r2p_bench.err; python - <<PY
import json
d = json.load(open('gpurun_out/r2p_bench_${w}.json'))
print('$w step %.1f Gpx/s (%.4f ms)  sustained %.1f  kernel alone %.4f ms = %.3f' % (d['value'], d['ms_per_step'], d['sustained']['value'], d['roofline']['kernel_ms'], d['roofline']['frac']))
PY
done
