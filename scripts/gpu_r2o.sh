#!/bin/bash
# overlapped one-sweep Reinhard (normalise pass of group g under the map sweep of group g + 1): parity + A/B
mkdir -p gpurun_out
python -m pytest tests/test_gpu_reinhard_map16.py tests/test_gpu_camera_isp.py tests/test_gpu_fullsize.py tests/test_gpu_pipeline.py tests/test_gpu_rig.py tests/test_gpu_distributed.py -m gpu -q > gpurun_out/pytest_r2o.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_r2o.log
for cfg in "1 2" "2 2" "2 1" "2 3" "3 2" "3 1" "6 1"; do set -- $cfg
B200ISP_MAP16_OVERLAP=$1 B200ISP_MAP16_OVERLAP_CTAS=$2 python bench.py --workload cfg3 --steps 200 --no-cpu-baseline --no-e2e --configs 0 > gpurun_out/r2o_bench_g$1_c$2.json 2>gpurun_out/r2o_bench.err; python - <<PY
import json
d = json.load(open('gpurun_out/r2o_bench_g$1_c$2.json'))
print('cfg3 groups=$1 ctas=$2 step %.1f Gpx/s (%.4f ms)  sustained %.1f' % (d['value'], d['ms_per_step'], d['sustained']['value']))
PY
done
