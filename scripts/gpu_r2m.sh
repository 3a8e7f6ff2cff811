#!/bin/bash
# reverse-order normalise pass, Camera16 table pass, cfg5 with the f16 / f32 ISP; Reinhard-path tests
mkdir -p gpurun_out
python -m pytest tests/test_gpu_reinhard_map16.py tests/test_gpu_camera_isp.py tests/test_gpu_golden.py tests/test_gpu_fullsize.py tests/test_gpu_resize_isp.py -m gpu -q > gpurun_out/pytest_r2m.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_r2m.log
for w in cfg3 cfg1 cfg1_16 cfg5 cfg5_32; do python bench.py --workload $w --steps 200 --no-cpu-baseline --no-e2e --configs 0 > gpurun_out/r2m_bench_${w}.json 2>gpurun_out/r2m_bench.err; python - <<PY
import json
d = json.load(open('gpurun_out/r2m_bench_${w}.json'))
print('$w step %.1f Gpx/s (%.4f ms)  sustained %.1f  kernel alone %.4f ms = %.3f' % (d['value'], d['ms_per_step'], d['sustained']['value'], d['roofline']['kernel_ms'], d['roofline']['frac']))
PY
done
python scripts/ops_bench.py > gpurun_out/ops_bench.txt 2>gpurun_out/ops_bench.err; echo "ops rc=$?"; cat gpurun_out/ops_bench.txt | cut -c1-170
