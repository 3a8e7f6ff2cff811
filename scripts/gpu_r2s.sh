#!/bin/bash
# gated fallback sweeps leave after one look at the per-frame flags: the Reinhard suites + the Reinhard -> RGB8 workloads
mkdir -p gpurun_out
python -m pytest tests/test_gpu_reinhard_map16.py tests/test_gpu_turned_pass.py tests/test_gpu_camera_isp.py tests/test_gpu_fullsize.py -m gpu -q > gpurun_out/r2s_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2s_pytest.log
for w in cfg1 cfg3 cfg3_rot90; do python bench.py --workload $w --steps 200 --no-cpu-baseline --no-e2e --configs 0 > gpurun_out/r2s_bench_${w}.json 2>gpurun_out/r2s_bench.err; python - <<PY
import json
d = json.load(open('gpurun_out/r2s_bench_${w}.json'))
print('$w step %.1f Gpx/s (%.4f ms)  sustained %.1f  kernel alone %.4f ms = %.3f' % (d['value'], d['ms_per_step'], d['sustained']['value'], d['roofline']['kernel_ms'], d['roofline']['frac']))
PY
done
