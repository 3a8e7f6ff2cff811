#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_resize_isp.py tests/test_gpu_golden.py tests/test_gpu_camera_isp.py -m gpu -q -x > gpurun_out/pytest_r2c.log 2>&1; echo "pytest rc=$?"; grep -E "fullsize|passed|failed|^E " gpurun_out/pytest_r2c.log | tail -20
python bench.py --workload cfg5 --steps 30 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/r2c_cfg5.json 2> gpurun_out/r2c_cfg5.err; echo "bench rc=$?"; cat gpurun_out/r2c_cfg5.json | head -c 600; echo
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2c_launches_cfg5.csv python bench.py --workload cfg5 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_l.log 2>&1
python - <<'PY'
import csv
rows = [r for r in csv.reader(open('gpurun_out/r2c_launches_cfg5.csv')) if len(r) > 5]
hdr = rows[0]
ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
for r in rows[-14:]:
    print(r[ki][:90], r[vi])
PY
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:resize_gather -s 2 -c 1 -f -o gpurun_out/r2c_gather python bench.py --workload cfg5 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_gather.log 2>&1; echo "ncu rc=$?"
