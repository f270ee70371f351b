#!/bin/bash
mkdir -p gpurun_out
export DPGICP_LIBRARY=$PWD/dpg_slam_b200/libdpgicp_tw24.so
CMD="python tools/gpu_probe2.py corridor 5000 3,0,0"
timeout 600 $CMD > gpurun_out/plain_p2.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum,sm__inst_executed.sum,smsp__cycles_active.avg,sm__cycles_elapsed.max --clock-control none -c 60 --csv --log-file gpurun_out/launches_st3.csv $CMD > gpurun_out/ncu_st3.log 2>&1
echo rc=$?
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_st3.csv')) if len(r)>14 and r[0].isdigit()]
cur={}
for r in rows:
    k=(r[0], r[4][:40], r[7], r[8])
    cur.setdefault(k,{})[r[12]]=r[14]
for k,v in cur.items(): print(k, v)
PY
