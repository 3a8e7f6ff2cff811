#!/bin/bash
# round-2 evidence run: full GPU test suite, smoke, default bench line, reference arm, launch list + full capture of the cfg2 sweep, ops bench
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_full.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_full.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?"
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r02_bench_reference_n1.json 2> gpurun_out/r02_bench_reference_n1.err; echo "reference rc=$?"
CMD="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --configs 0 --min-seconds 0"
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:"isp::|mailbox" -c 400 --csv --log-file gpurun_out/r02_launches_cfg2.csv $CMD > gpurun_out/ncu_launches.log 2>&1; echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:stream2_kernel -s 6 -c 1 -f -o gpurun_out/r02_stream2_cfg2 $CMD > gpurun_out/ncu_full_cfg2.log 2>&1; echo "full rc=$?"
python scripts/ops_bench.py > gpurun_out/r02_ops_bench.txt 2> gpurun_out/ops_bench.err; echo "ops rc=$?"
