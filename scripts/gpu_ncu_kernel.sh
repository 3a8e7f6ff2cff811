#!/bin/bash
# ncu --set full of one kernel selected by a regex on the DEMANGLED name: KERNEL=regex  SKIP=n  BENCH_ARGS=...
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e ${BENCH_ARGS}"
$CMD > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:$KERNEL -s ${SKIP:-3} -c 1 -f -o gpurun_out/prof_kernel $CMD > gpurun_out/ncu_full.log 2>&1
echo "full rc=$?"; tail -3 gpurun_out/ncu_full.log
