"""Development aid: per-pass latency of ONE pair (the longest of the corridor batch) for each CTA width."""
import os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from dpg_slam_b200 import synth
from dpg_slam_b200._abi import Params, COV_CENSI_CORR
from dpg_slam_b200.scanmatch import ScanMatcher
from gpu_probe import time_run

wl = synth.config_corridor(n_pairs=5000, seed=2)
p = Params.defaults(downsample_divisor=1, cov_mode=COV_CENSI_CORR)
with ScanMatcher(0) as sm:
    sm.upload_ranges(wl.ranges, wl.scanner)
    rec = sm.submit_pairs(wl.src_idx, wl.tgt_idx, wl.guess, p)
k = int(np.argmax(rec["iterations"]))
its = int(rec["iterations"][k])
print("longest pair", k, "iterations", its, flush=True)
for warps in (1, 2, 4, 8, 16):
    os.environ["DPGICP_STAGES"], os.environ["DPGICP_WARPS"] = "1", str(warps)
    with ScanMatcher(0) as sm:
        sm.upload_ranges(wl.ranges, wl.scanner)
        for n in (1, 148):
            idx = np.full(n, k)
            sm.set_pairs(wl.src_idx[idx], wl.tgt_idx[idx], wl.guess[idx])
            best, med = time_run(sm, p, reps=3)
            print(f"W={warps} copies={n}: {best:.3f} ms -> {1e3 * best / (its + 1):.2f} us per pass", flush=True)
