#!/bin/bash
# round 2: Reinhard MUFU diet -- parity subset, bench sweep over the group size, ncu of the two sweeps
mkdir -p gpurun_out
python -m pytest tests/test_gpu_camera_isp.py tests/test_gpu_golden.py tests/test_gpu_bilinear_isp.py tests/test_gpu_fullsize.py -m gpu -q -x -k "reinhard or golden or vs_c_oracle or wide" > gpurun_out/pytest_r2b.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_r2b.log
B="--steps 60 --warmup 5 --no-cpu-baseline --no-e2e"
for w in cfg1 cfg3 cfg1_16; do
  for g in 0 2; do
    B200ISP_REINHARD_GROUP=$g python bench.py --workload $w $B > gpurun_out/r2b_${w}_g$g.json 2> gpurun_out/r2b_${w}_g$g.err; echo "bench $w group $g rc=$?"
  done
done
B200ISP_CAM16_RECOMPUTE=1 python bench.py --workload cfg1_16 $B > gpurun_out/r2b_cfg1_16_recompute.json 2> gpurun_out/r2b_cfg1_16_recompute.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2b_*.json')):
    try:
        d = json.load(open(f))
        print(f.split('/')[-1], 'value %.1f Gpx/s  ms/step %.4f  kernel_ms %.4f' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms']))
    except Exception as e:
        print(f, 'ERR', e)
PY
for k in EpiReinhard2 EpiReinhardMax2; do
  CMD="python bench.py --workload cfg3 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --graph 0 --lookahead 0"
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:$k -s 3 -c 1 -f -o gpurun_out/r2b_$k $CMD > gpurun_out/ncu_$k.log 2>&1; echo "ncu $k rc=$?"
done
