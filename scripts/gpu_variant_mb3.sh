#!/bin/bash
# experiment: 3 CTAs of 168 registers per SM instead of 4 x 128 (build.py --out with B200ISP_NVCC_EXTRA=-DISP_S2_MINBLOCKS=3)
mkdir -p gpurun_out
for lib in "" variants/libb200isp_mb3.so; do
for w in cfg3 cfg1_16 cfg5 cfg2; do B200ISP_LIB=$lib python bench.py --workload $w --steps 200 --no-cpu-baseline --no-e2e --configs 0 > gpurun_out/mb3_bench.json 2>gpurun_out/mb3_bench.err; python - <<PY
import json
d = json.load(open('gpurun_out/mb3_bench.json'))
print('$w lib=${lib:-product} step %.1f Gpx/s (%.4f ms)  sustained %.1f  kernel alone %.4f ms = %.3f' % (d['value'], d['ms_per_step'], d['sustained']['value'], d['roofline']['kernel_ms'], d['roofline']['frac']))
PY
done
done
