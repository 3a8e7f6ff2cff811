#!/bin/bash
# usage: gpu_multi.sh N [workload] [extra bench args]   -- every command under its own timeout (a hang costs N x GPU time)
N=${1:-2}; WL=${2:-cfg2}; shift 2
mkdir -p gpurun_out
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 30 --warmup 5 --workload $WL --no-e2e "$@" > gpurun_out/bench_n${N}_$WL.json 2> gpurun_out/bench_n${N}_$WL.err; echo "rc=$?"
cat gpurun_out/bench_n${N}_$WL.json; grep -v "^\*\*\*\|OMP_NUM\|^$" gpurun_out/bench_n${N}_$WL.err | tail -8
