#!/bin/bash
# ncu --set full capture of the dominant kernel (run only after the same command exited 0 without ncu)
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e $BENCH_ARGS"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:${NCU_KERNEL:-stream_kernel} -s 3 -c 2 -f -o gpurun_out/${NCU_OUT:-prof_stream} $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu rc=$?"; tail -5 gpurun_out/ncu_full.log
