#!/bin/bash
# final tree of round 2 (turned normalise pass, transform_turn_kernel): full GPU suite, smoke, the default bench line,
# launch list + ncu --set full capture of the turned pass on cfg3_rot90
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
python bench.py > gpurun_out/r2r_bench_n1.json 2> gpurun_out/r2r_bench_n1.err; echo "bench rc=$?"
python - <<PY
import json
d = json.load(open('gpurun_out/r2r_bench_n1.json'))
print('cfg2 %.1f Gpx/s (%.4f ms) sustained %.1f kernel %.4f ms frac %.3f e2e %.2f' % (d['value'], d['ms_per_step'], d['sustained']['value'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['e2e']['value']))
for c in d.get('configs') or []:
    print('  %-60s %.1f Gpx/s  %.4f ms  kernel %.4f ms frac %.3f' % (c['workload'][:60], c['value'], c['ms_per_step'], c['roofline']['kernel_ms'], c['roofline']['frac']))
PY
CMD="python bench.py --workload cfg3_rot90 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --configs 0"
$CMD > gpurun_out/r2r_rot90_plain.json 2>gpurun_out/r2r_rot90_plain.err; echo "plain rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2r_launches_cfg3_rot90.csv $CMD > gpurun_out/r2r_ncu_launches.log 2>&1; echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:reinhard_out_transposed -s 3 -c 1 -f -o gpurun_out/prof_turned_pass $CMD > gpurun_out/r2r_ncu_turned.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/r2r_ncu_turned.log
