#!/bin/bash
# bench cfg2 (and optionally others) with every experimental library under variants/ plus the product build
mkdir -p gpurun_out
if [ "$RUN_TESTS" = "1" ]; then python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log; fi
shopt -s nullglob
for lib in default variants/*.so; do
  for wl in ${WORKLOADS:-cfg2}; do
    if [ "$lib" = "default" ]; then unset B200ISP_LIB; else export B200ISP_LIB=$PWD/$lib; fi
    python bench.py --workload $wl --steps 40 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/v.json 2> gpurun_out/v.err || tail -3 gpurun_out/v.err
    python - <<PY
import json
d=json.load(open('gpurun_out/v.json'))
print('$lib', '$wl', 'value %.1f Gpx/s'%d['value'], 'ms/step %.4f'%d['ms_per_step'], 'kernel_ms %.4f'%d['roofline']['kernel_ms'], 'frac %.3f'%d['roofline']['frac'])
PY
  done
done
