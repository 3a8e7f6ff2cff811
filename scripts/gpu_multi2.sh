#!/bin/bash
# N-GPU runs: weak scaling cfg2 with shared exposure, strong scaling cfg3 with 12 cameras; topology probe
N=${N:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_n$N.txt 2>&1
(numactl -H; cat /proc/self/status | grep -i cpus_allowed_list; nproc) >> gpurun_out/topo_n$N.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus $N --steps 200 --warmup 5 > gpurun_out/r2e_n${N}_cfg2.json 2> gpurun_out/r2e_n${N}_cfg2.err; echo "cfg2 rc=$?"; tail -3 gpurun_out/r2e_n${N}_cfg2.err
$TR bench.py --gpus $N --steps 100 --warmup 5 --workload cfg3 --cameras 12 --no-e2e > gpurun_out/r2e_n${N}_cfg3cam12.json 2> gpurun_out/r2e_n${N}_cfg3cam12.err; echo "cfg3 rc=$?"; tail -3 gpurun_out/r2e_n${N}_cfg3cam12.err
if [ "$N" = "8" ]; then
  $TR bench.py --gpus $N --steps 100 --warmup 5 --workload cfg5 --no-e2e > gpurun_out/r2e_n${N}_cfg5.json 2> gpurun_out/r2e_n${N}_cfg5.err; echo "cfg5 rc=$?"
fi
python - <<PY
import json
for f in ('gpurun_out/r2e_n${N}_cfg2.json', 'gpurun_out/r2e_n${N}_cfg3cam12.json', 'gpurun_out/r2e_n${N}_cfg5.json'):
    try:
        d = json.load(open(f))
        print(f, 'value %.1f sustained %.1f ms %.4f' % (d['value'], d['sustained']['value'], d['ms_per_step']), 'parity', d['parity_nranks'], 'e2e', (d['e2e'] or {}).get('value'), (d['e2e'] or {}).get('host_ceiling'))
    except Exception as e:
        print(f, 'ERR', e)
PY
