#!/bin/bash
mkdir -p gpurun_out
export DPGICP_CHAIN=4,8,32
CMD="python tools/gpu_probe2.py corridor 5000 4,8,32"
timeout 600 $CMD > gpurun_out/plain_p2.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum,sm__inst_executed.sum,smsp__cycles_active.avg,sm__cycles_elapsed.max --clock-control none -c 14 --csv --log-file gpurun_out/launches_chain.csv $CMD > gpurun_out/ncu_chain.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_chain.csv')) if len(r)>14 and r[0].isdigit()]
cur={}
for r in rows:
    k=(r[0], r[4][5:30], r[7], r[8])
    cur.setdefault(k,{})[r[12][:24]]=r[14]
for k,v in cur.items(): print(k, v)
PY
unset DPGICP_CHAIN
python - <<'PY'
import os, sys
sys.path.insert(0, 'tools'); sys.path.insert(0, '.')
import numpy as np
from dpg_slam_b200 import synth
from dpg_slam_b200._abi import Params, COV_CENSI_CORR
from dpg_slam_b200.scanmatch import ScanMatcher
from gpu_probe import time_run
wl = synth.config_corridor(n_pairs=5000, seed=2)
p = Params.defaults(downsample_divisor=1, cov_mode=COV_CENSI_CORR)
k, its = 353, 305
for chain in ("4", "7", "9", "12", "17", "32"):
    os.environ["DPGICP_CHAIN"] = chain
    with ScanMatcher(0) as sm:
        sm.upload_ranges(wl.ranges, wl.scanner)
        idx = np.full(1, k)
        sm.set_pairs(wl.src_idx[idx], wl.tgt_idx[idx], wl.guess[idx])
        best, med = time_run(sm, p, reps=3)
        print(f"chain={chain}: {best:.3f} ms -> {1e3 * best / (its + 1):.2f} us per pass", flush=True)
PY
