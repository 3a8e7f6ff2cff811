#!/bin/bash
# round artefacts for profiles/: default bench line, ncu launch list of the same command, one --set full capture
# of the dominant kernel (only after the plain command exited 0)
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"; cut -c1-600 gpurun_out/bench_default.json
CMD="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu1.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:stream2_kernel -s 4 -c 1 -f -o gpurun_out/prof_stream2 $CMD > gpurun_out/ncu_full.log 2>&1
echo "full rc=$?"
