#!/bin/bash
# rows_per_task sweep for the small-launch workloads, then ONE ncu --set full capture (KERNEL / BENCH_ARGS as in gpu_ncu_kernel.sh)
mkdir -p gpurun_out
for wl in ${WORKLOADS:-cfg1 cfg3}; do
  for rpt in ${RPTS:-0 12 16 20 24 32 48}; do
    python bench.py --workload $wl --rows-per-task $rpt --steps 60 --warmup 10 --no-cpu-baseline --no-e2e > gpurun_out/v.json 2> gpurun_out/v.err || tail -3 gpurun_out/v.err
    python - <<PY
import json
d=json.load(open('gpurun_out/v.json'))
print('$wl', 'rpt $rpt', 'value %.1f Gpx/s'%d['value'], 'ms/step %.4f'%d['ms_per_step'], 'kernel_ms %.4f'%d['roofline']['kernel_ms'])
PY
  done
done
if [ -n "$KERNEL" ]; then bash scripts/gpu_ncu_kernel.sh; fi
