#!/bin/bash
# final tree: full GPU suite + smoke, then the ncu launch list of the default bench command
bash scripts/gpu_tests_smoke.sh
COUNT=60 bash scripts/gpu_launches.sh > gpurun_out/launch_summary.txt 2>&1; cat gpurun_out/launch_summary.txt
