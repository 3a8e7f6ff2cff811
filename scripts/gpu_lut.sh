#!/bin/bash
# full GPU suite + A/B of the table-based normalise pass (B200ISP_REINHARD_LUT=0: arithmetic pass)
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_full.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_full.log
for lut in 1 0; do
for w in cfg3 cfg1_16 cfg1; do B200ISP_REINHARD_LUT=$lut python bench.py --workload $w --steps 200 --no-cpu-baseline --no-e2e --configs 0 > gpurun_out/lut_bench_${w}_lut$lut.json 2>gpurun_out/lut_bench.err; python - <<PY
import json
d = json.load(open('gpurun_out/lut_bench_${w}_lut$lut.json'))
print('$w lut=$lut step %.1f Gpx/s (%.4f ms)  sustained %.1f  kernel alone %.4f ms = %.3f' % (d['value'], d['ms_per_step'], d['sustained']['value'], d['roofline']['kernel_ms'], d['roofline']['frac']))
PY
done
done
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:isp:: -s 40 -c 40 --csv --log-file gpurun_out/lut_launches.csv python bench.py --workload cfg3 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --configs 0 --min-seconds 0 > gpurun_out/lut_ncu.log 2>&1
python - <<'PY'
import csv, re
rows = [r for r in csv.reader(open('gpurun_out/lut_launches.csv')) if len(r) > 5]
for i, r in enumerate(rows):
    if 'Kernel Name' in r:
        hdr, rows = r, rows[i + 1:]
        break
ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
for r in rows[-10:]:
    print('%-110s %10.1f us' % (re.sub(r'\(bool\)|\(int\)|isp::', '', r[ki])[:110], float(r[vi].replace(',', '')) / 1000))
PY
