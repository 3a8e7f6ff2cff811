#!/bin/bash
# round 2, one-sweep Camera32 Reinhard -> u8 (u16 fixed-point map): parity tests, bench (new path vs exact two-sweep form), launch list
mkdir -p gpurun_out
python -m pytest tests/test_gpu_reinhard_map16.py tests/test_gpu_camera_isp.py tests/test_gpu_golden.py tests/test_gpu_fullsize.py tests/test_gpu_bilinear_isp.py tests/test_gpu_pipeline.py tests/test_gpu_rig.py tests/test_gpu_bayer.py -m gpu -q > gpurun_out/pytest_map16.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_map16.log
for mode in 0 1; do
for w in cfg3 cfg1 cfg2; do B200ISP_REINHARD_EXACT=$mode python bench.py --workload $w --steps 200 --no-cpu-baseline --no-e2e --configs 0 > gpurun_out/map16_bench_${w}_exact$mode.json 2>gpurun_out/map16_bench.err; python - <<PY
import json
d = json.load(open('gpurun_out/map16_bench_${w}_exact$mode.json'))
print('$w exact=$mode step %.1f Gpx/s (%.4f ms)  sustained %.1f  kernel alone %.4f ms = %.3f  launches %s' % (d['value'], d['ms_per_step'], d['sustained']['value'], d['roofline']['kernel_ms'], d['roofline']['frac'], d.get('gpu_launches')))
PY
done
done
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:isp:: -s 40 -c 60 --csv --log-file gpurun_out/map16_launches.csv python bench.py --workload cfg3 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --configs 0 --min-seconds 0 > gpurun_out/map16_ncu.log 2>&1
echo "launch list rc=$?"
python - <<'PY'
import csv, re
rows = [r for r in csv.reader(open('gpurun_out/map16_launches.csv')) if len(r) > 5]
for i, r in enumerate(rows):
    if 'Kernel Name' in r:
        hdr, rows = r, rows[i + 1:]
        break
ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
for r in rows[-14:]:
    n = r[ki]
    print('%-110s %10.1f us' % (re.sub(r'\(bool\)|\(int\)|isp::', '', n)[:110], float(r[vi].replace(',', '')) / 1000))
PY
