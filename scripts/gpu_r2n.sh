#!/bin/bash
# lean saturating front end of the Reinhard sweeps + fat normalise pass: parity, then cfg3 with 2 / 4 / 8 CTAs per SM in the pass
mkdir -p gpurun_out
python -m pytest tests/test_gpu_reinhard_map16.py tests/test_gpu_camera_isp.py tests/test_gpu_golden.py tests/test_gpu_fullsize.py tests/test_gpu_rig.py -m gpu -q > gpurun_out/pytest_r2n.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_r2n.log
for c in 4 2 8; do
for w in cfg3 cfg1; do B200ISP_MAP16_CTAS=$c python bench.py --workload $w --steps 200 --no-cpu-baseline --no-e2e --configs 0 > gpurun_out/r2n_bench_${w}_c$c.json 2>gpurun_out/r2n_bench.err; python - <<PY
import json
d = json.load(open('gpurun_out/r2n_bench_${w}_c$c.json'))
print('$w ctas/SM=$c step %.1f Gpx/s (%.4f ms)  sustained %.1f  kernel alone %.4f ms = %.3f' % (d['value'], d['ms_per_step'], d['sustained']['value'], d['roofline']['kernel_ms'], d['roofline']['frac']))
PY
done
done
B200ISP_REINHARD_EXACT=1 python bench.py --workload cfg3 --steps 200 --no-cpu-baseline --no-e2e --configs 0 > gpurun_out/r2n_bench_cfg3_exact.json 2>>gpurun_out/r2n_bench.err; python - <<PY
import json
d = json.load(open('gpurun_out/r2n_bench_cfg3_exact.json'))
print('cfg3 exact step %.1f Gpx/s (%.4f ms)  write sweep alone %.4f ms' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms']))
PY
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:isp:: -s 40 -c 24 --csv --log-file gpurun_out/r2n_launches.csv python bench.py --workload cfg3 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --configs 0 --min-seconds 0 > gpurun_out/r2n_ncu.log 2>&1
python - <<'PY'
import csv, re
rows = [r for r in csv.reader(open('gpurun_out/r2n_launches.csv')) if len(r) > 5]
for i, r in enumerate(rows):
    if 'Kernel Name' in r:
        hdr, rows = r, rows[i + 1:]
        break
ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
for r in rows[-7:]:
    print('%-110s %10.1f us' % (re.sub(r'\(bool\)|\(int\)|isp::', '', r[ki])[:110], float(r[vi].replace(',', '')) / 1000))
PY
