#!/bin/bash
# final tree of round 2: full GPU suite + smoke, the default bench line, the reference arm, the launch list of the default command
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_n1.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_bench_reference_n1.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
python - <<'PY'
import json
d = json.load(open('gpurun_out/r02_bench_n1.json'))
print('value %.1f sustained %.1f ms %.4f frac %.3f kernel_ms %.4f e2e %.2f cpu %.3f' % (d['value'], d['sustained']['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms'], d['e2e']['value'], d['cpu_baseline']['value']))
for c in d['configs']:
    print('  %-40s %.1f Gpx/s  kernel %.4f ms frac %.3f' % (c['workload'][:40], c['value'], c['roofline']['kernel_ms'], c['roofline']['frac']))
for v in d.get('e2e_variants') or []:
    print('  e2e %-60s %.2f' % (v['workload'][:60], v['value']))
r = json.load(open('gpurun_out/r02_bench_reference_n1.json'))
print('reference arm', r.get('value'), r.get('unit'), r.get('cpu_baseline', {}).get('cores'))
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_cfg2.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --configs 0 --min-seconds 0 > gpurun_out/ncu_launches.log 2>&1; echo "launch list rc=$?"
