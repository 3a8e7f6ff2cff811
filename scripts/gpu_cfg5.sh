#!/bin/bash
# cfg5 (ISP + resize, staged kernels): frames per demosaic-sweep launch
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for g in ${GROUPS_:-1 2 4 8}; do
  B200ISP_RESIZE_GROUP=$g python bench.py --workload cfg5 --steps 30 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/v.json 2> gpurun_out/v.err || tail -3 gpurun_out/v.err
  python - <<PY
import json
d=json.load(open('gpurun_out/v.json'))
print('cfg5 group $g', 'value %.1f Gpx/s'%d['value'], 'ms/step %.4f'%d['ms_per_step'])
PY
done
