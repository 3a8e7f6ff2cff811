#!/bin/bash
# usage: gpu_multi.sh N [workload]
N=${1:-2}; WL=${2:-cfg2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --workload $WL > gpurun_out/bench_n${N}_$WL.json 2> gpurun_out/bench_n${N}_$WL.err; echo "rc=$?"
cat gpurun_out/bench_n${N}_$WL.json; tail -5 gpurun_out/bench_n${N}_$WL.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 3 --warmup 1 --impl reference > gpurun_out/bench_ref_n${N}.json 2> gpurun_out/bench_ref_n${N}.err; echo "ref rc=$?"; cat gpurun_out/bench_ref_n${N}.json
