#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
for wl in cfg2 cfg1 cfg1_16; do
python bench.py --workload $wl --steps 30 --warmup 5 --no-cpu-baseline --no-e2e $BENCH_ARGS > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err; python - <<PY
import json
d=json.load(open('gpurun_out/bench_$wl.json'))
print('$wl', 'value %.1f Gpx/s'%d['value'], 'ms/step %.4f'%d['ms_per_step'], 'kernel_ms %.4f'%d['roofline']['kernel_ms'], 'frac %.3f'%d['roofline']['frac'], d['clocks'])
PY
tail -2 gpurun_out/bench_$wl.err
done
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_cfg2.csv $CMD > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:stream2_kernel -s 3 -c 1 -f -o gpurun_out/prof_stream $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu rc=$?"
