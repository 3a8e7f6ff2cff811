#!/bin/bash
# ncu --set full of the sweep of one transformed workload: WL=cfg2 T=rotate_90 OUT=name
mkdir -p gpurun_out
python scripts/transform_once.py ${WL:-cfg2} ${T:-rotate_90} || exit 1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:stream2_kernel -s 2 -c 1 -f -o gpurun_out/${OUT:-xpose} \
  python scripts/transform_once.py ${WL:-cfg2} ${T:-rotate_90} > gpurun_out/ncu_${OUT:-xpose}.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_${OUT:-xpose}.log
