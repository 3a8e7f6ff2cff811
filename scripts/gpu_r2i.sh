#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_camera_isp.py tests/test_gpu_golden.py tests/test_gpu_bilinear_isp.py tests/test_gpu_fullsize.py tests/test_gpu_pipeline.py tests/test_gpu_rig.py -m gpu -q -x > gpurun_out/pytest_r2i.log 2>&1; echo "pytest rc=$?"; grep -E "fullsize\] cfg|passed|failed|^E " gpurun_out/pytest_r2i.log | tail -16
B="--steps 100 --warmup 5 --no-cpu-baseline --no-e2e --configs 0"
for w in cfg1 cfg3; do
  for two in 0 1; do
    B200ISP_CAM32_TWO_SWEEPS=$two python bench.py --workload $w $B > gpurun_out/r2i_${w}_two$two.json 2> gpurun_out/r2i_${w}_two$two.err; echo "bench $w two=$two rc=$?"
  done
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2i_*.json')):
    try:
        d = json.load(open(f))
        print(f.split('/')[-1], 'value %.1f sustained %.1f Gpx/s  ms/step %.4f  kernel_ms %.4f' % (d['value'], d['sustained']['value'], d['ms_per_step'], d['roofline']['kernel_ms']))
    except Exception as e:
        print(f, 'ERR', e)
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r2i_launches_cfg3.csv python bench.py --workload cfg3 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --configs 0 --min-seconds 0 --graph 0 > gpurun_out/ncu_l.log 2>&1
python - <<'PY'
import csv
rows = [r for r in csv.reader(open('gpurun_out/r2i_launches_cfg3.csv')) if len(r) > 5]
hdr = rows[0]
ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
for r in rows[-10:]:
    print(r[ki][:100], r[vi])
PY
