#!/bin/bash
# ncu --set full of one kernel: KERNEL=regex SKIP=n OUT=name BENCH_ARGS="..."
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e ${BENCH_ARGS}"
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:$KERNEL -s ${SKIP:-3} -c 1 -f -o gpurun_out/${OUT:-prof} $CMD > gpurun_out/ncu_${OUT:-prof}.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_${OUT:-prof}.log
