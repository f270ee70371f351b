#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
timeout 600 $CMD > gpurun_out/plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum,sm__inst_executed.sum,smsp__cycles_active.avg,sm__cycles_elapsed.max --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches rc=$?"
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/launches.csv')) if len(r)>14 and r[0].isdigit()]
cur=collections.OrderedDict()
for r in rows:
    k=(int(r[0]), r[4].split('(')[0][-36:], r[7], r[8])
    cur.setdefault(k,{})[r[12][:22]]=float(r[14])
for k,v in list(cur.items())[:24]: print(k, {a: round(b) for a,b in v.items()})
PY
