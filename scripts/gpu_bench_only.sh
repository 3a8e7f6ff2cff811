#!/bin/bash
mkdir -p gpurun_out
for wl in cfg2 cfg1 cfg1_16; do
python bench.py --workload $wl --steps 50 --warmup 5 --no-cpu-baseline --no-e2e $BENCH_ARGS > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err; python - <<PY
import json
d=json.load(open('gpurun_out/bench_$wl.json'))
print('$wl', 'value %.1f Gpx/s'%d['value'], 'ms/step %.4f'%d['ms_per_step'], 'kernel_ms %.4f'%d['roofline']['kernel_ms'], 'frac %.3f'%d['roofline']['frac'], d['clocks'])
PY
tail -2 gpurun_out/bench_$wl.err
done
