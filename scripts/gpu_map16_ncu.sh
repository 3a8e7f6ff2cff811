#!/bin/bash
# ncu --set full of the map sweep and the normalise pass of the one-sweep Camera32 Reinhard path (cfg3), plus the cfg2 sweep
mkdir -p gpurun_out
CMD="python bench.py --workload cfg3 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --configs 0 --min-seconds 0"
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain.log; exit 1; }
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:EpiReinhardMax2<.bool.0, .bool.1, .bool.1" -s 6 -c 1 -f -o gpurun_out/prof_map16_sweep $CMD > gpurun_out/ncu_a.log 2>&1; echo "sweep rc=$?"
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:reinhard_map16_out_kernel" -s 6 -c 1 -f -o gpurun_out/prof_map16_out $CMD > gpurun_out/ncu_b.log 2>&1; echo "out rc=$?"
CMD2="python bench.py --workload cfg2 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --configs 0 --min-seconds 0"
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:EpiLinear2" -s 6 -c 1 -f -o gpurun_out/prof_cfg2_sweep $CMD2 > gpurun_out/ncu_c.log 2>&1; echo "cfg2 rc=$?"
python scripts/ncu_summary.py gpurun_out/prof_map16_sweep.ncu-rep 73728000 > gpurun_out/r02_map16_sweep_ncu.txt 2>&1
python scripts/ncu_summary.py gpurun_out/prof_map16_out.ncu-rep 73728000 > gpurun_out/r02_map16_out_ncu.txt 2>&1
python scripts/ncu_summary.py gpurun_out/prof_cfg2_sweep.ncu-rep 119771136 > gpurun_out/r02_cfg2_sweep_fadd2_ncu.txt 2>&1
rm -f gpurun_out/*.ncu-rep
head -30 gpurun_out/r02_map16_sweep_ncu.txt gpurun_out/r02_map16_out_ncu.txt gpurun_out/r02_cfg2_sweep_fadd2_ncu.txt
