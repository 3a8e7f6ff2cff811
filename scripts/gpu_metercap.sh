#!/bin/bash
mkdir -p gpurun_out
for w in cfg5 cfg2 cfg3; do
  for cap in 0 74 148 296; do
    B200ISP_METER_GRID_CAP=$cap python bench.py --workload $w --steps 200 --no-cpu-baseline --no-e2e --configs 0 --min-seconds 0.3 > gpurun_out/mc.json 2>/dev/null
    python - <<PY
import json
d = json.load(open('gpurun_out/mc.json'))
print('$w cap=$cap step %.1f Gpx/s (%.4f ms) sustained %.1f' % (d['value'], d['ms_per_step'], d['sustained']['value']))
PY
  done
done
