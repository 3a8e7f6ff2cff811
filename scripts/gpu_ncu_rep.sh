#!/bin/bash
# one ncu --set full capture brought back as a report: KERNEL=regex on the demangled name, BENCH_ARGS, OUT=name
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --configs 0 --min-seconds 0 ${BENCH_ARGS}"
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; exit 1; }
ncu --set full --clock-control none --kernel-name-base demangled -k "regex:${KERNEL}" -s ${SKIP:-6} -c 1 -f -o gpurun_out/${OUT:-prof} $CMD > gpurun_out/ncu_rep.log 2>&1; echo "rc=$?"
ls -la gpurun_out/*.ncu-rep
