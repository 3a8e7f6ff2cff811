#!/bin/bash
# compare library variants on the cfg2 step: VARIANTS="a.so b.so" (shipped library always first)
mkdir -p gpurun_out
for lib in "" $VARIANTS; do
  for i in 1 2; do
    B200ISP_LIB=$lib timeout 300 python bench.py --steps 400 --no-cpu-baseline --no-e2e --configs 0 ${BENCH_ARGS} > gpurun_out/var_bench.json 2> gpurun_out/var_bench.err
    python - <<PY
import json
d = json.load(open('gpurun_out/var_bench.json'))
r = d['roofline']
print('lib=${lib:-shipped}', 'step %.1f Gpx/s (%.4f ms)  sustained %.1f  kernel alone %.4f ms = %.3f of peak  after sustained %.4f ms' % (d['value'], d['ms_per_step'], d['sustained']['value'], r['kernel_ms'], r['frac'], r['after_sustained_window']['kernel_ms']))
PY
  done
done
