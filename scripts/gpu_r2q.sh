#!/bin/bash
# predicated staged stores (st_out_lt): full GPU suite, the bench workloads, the stand-alone operators
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for w in cfg3 cfg1 cfg1_16 cfg5 cfg2; do python bench.py --workload $w --steps 200 --no-cpu-baseline --no-e2e --configs 0 > gpurun_out/r2q_bench_${w}.json 2>gpurun_out/r2q_bench.err; python - <<PY
import json
d = json.load(open('gpurun_out/r2q_bench_${w}.json'))
print('$w step %.1f Gpx/s (%.4f ms)  sustained %.1f  kernel alone %.4f ms = %.3f' % (d['value'], d['ms_per_step'], d['sustained']['value'], d['roofline']['kernel_ms'], d['roofline']['frac']))
PY
done
B200ISP_REINHARD_EXACT=1 python bench.py --workload cfg3 --steps 200 --no-cpu-baseline --no-e2e --configs 0 > gpurun_out/r2q_bench_cfg3_exact.json 2>>gpurun_out/r2q_bench.err; python - <<PY
import json
d = json.load(open('gpurun_out/r2q_bench_cfg3_exact.json'))
print('cfg3 exact step %.1f Gpx/s (%.4f ms)  write sweep alone %.4f ms' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms']))
PY
python scripts/ops_bench.py > gpurun_out/ops_bench.txt 2>gpurun_out/ops_bench.err; echo "ops rc=$?"; head -9 gpurun_out/ops_bench.txt | cut -c1-150
