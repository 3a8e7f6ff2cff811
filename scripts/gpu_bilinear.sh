#!/bin/bash
# bilinear-in-the-fused-sweep check: new parity tests first, then the whole GPU suite, then Malvar / bilinear bench lines
mkdir -p gpurun_out
python -m pytest tests/test_gpu_bilinear_isp.py -m gpu -q -x > gpurun_out/pytest_bilinear.log 2>&1; echo "bilinear pytest rc=$?"; tail -15 gpurun_out/pytest_bilinear.log | cut -c1-300
python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for wl in ${WORKLOADS:-cfg2 cfg1 cfg3}; do
  for dm in malvar bilinear; do
    python bench.py --workload $wl --demosaic $dm --steps 40 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/v.json 2> gpurun_out/v.err || tail -3 gpurun_out/v.err
    python - <<PY
import json
d=json.load(open('gpurun_out/v.json'))
print('$wl', '$dm', 'value %.1f Gpx/s'%d['value'], 'ms/step %.4f'%d['ms_per_step'], 'kernel_ms %.4f'%d['roofline']['kernel_ms'], 'frac %.3f'%d['roofline']['frac'])
PY
  done
done
