#!/bin/bash
# ncu launch list (+ optional full capture of one kernel: KERNEL=regex) of the default bench command
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e ${BENCH_ARGS}"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c ${COUNT:-40} --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu1.log 2>&1
echo "launch list rc=$?"
python - <<'PY'
import csv, collections
rows=list(csv.reader(open('gpurun_out/launches.csv')))
h=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
agg=collections.OrderedDict()
for r in rows[h+1:]:
    if len(r) < 15: continue
    name=r[4].split('(')[0][:90]
    agg.setdefault(name, []).append(float(r[-1].replace(',','')))
for k,v in agg.items(): print(f"{len(v):3d} x {sum(v)/len(v)/1000:9.2f} us  {k}")
PY
if [ -n "$KERNEL" ]; then
ncu --set full --clock-control none --import-source on -k regex:$KERNEL -s 3 -c 1 -f -o gpurun_out/prof_$KERNEL $CMD > gpurun_out/ncu_full.log 2>&1; echo "full rc=$?"
fi
