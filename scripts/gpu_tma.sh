#!/bin/bash
# TMA row-ring experiment (ISP_S2_RING=2, variants/libb200isp_tma.so) against the shipped register-fetch sweep on cfg2
mkdir -p gpurun_out
V=variants/libb200isp_tma.so
B200ISP_LIB=$V timeout 300 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_camera_isp.py -m gpu -q -x -k "cfg2 or fused_linear or wide_frames" > gpurun_out/pytest_tma.log 2>&1; echo "pytest(tma) rc=$?"; tail -3 gpurun_out/pytest_tma.log
for lib in "" $V; do
  for i in 1 2; do
    B200ISP_LIB=$lib timeout 300 python bench.py --steps 400 --no-cpu-baseline --no-e2e --configs 0 > gpurun_out/tma_bench.json 2> gpurun_out/tma_bench.err
    python - <<PY
import json
d = json.load(open('gpurun_out/tma_bench.json'))
r = d['roofline']
print('lib=${lib:-shipped}', 'step %.1f Gpx/s (%.4f ms)  sustained %.1f  kernel alone %.4f ms = %.3f of peak  after sustained %.4f ms' % (d['value'], d['ms_per_step'], d['sustained']['value'], r['kernel_ms'], r['frac'], r['after_sustained_window']['kernel_ms']))
PY
  done
done
B200ISP_LIB=$V ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:stream2_kernel -s 6 -c 1 -f -o gpurun_out/r2g_tma python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --configs 0 --min-seconds 0 > gpurun_out/ncu_tma.log 2>&1; echo "ncu rc=$?"
