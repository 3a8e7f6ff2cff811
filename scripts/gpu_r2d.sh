#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_distributed.py tests/test_gpu_pipeline.py tests/test_gpu_camera_isp.py -m gpu -q -x > gpurun_out/pytest_r2d.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_r2d.log
python bench.py > gpurun_out/r2d_bench_default.json 2> gpurun_out/r2d_bench_default.err; echo "bench rc=$?"; tail -5 gpurun_out/r2d_bench_default.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2d_bench_default.json'))
print('value %.1f sustained %.1f ms/step %.4f frac %.3f step_frac %.3f' % (d['value'], d['sustained']['value'], d['ms_per_step'], d['roofline']['frac'], d['step_frac_of_peak']))
print('clocks', d['clocks'])
print('e2e', d['e2e']['value'], d['e2e']['host_gbs'], d['e2e']['host_ceiling'])
print('shared_n1', d['shared_exposure_at_n1'])
for c in d['configs'] or []:
    print(c['workload'][:40], 'value %.1f ms %.4f frac %.3f kern_ms %.4f' % (c['value'], c['ms_per_step'], c['roofline']['frac'], c['roofline']['kernel_ms']))
for e in d['e2e_variants'] or []:
    print('e2e', e['workload'][:60], e['value'], e['host_gbs'])
print('cpu', d['cpu_baseline'])
PY
