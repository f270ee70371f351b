"""Development aid: per-pass latency of ONE long pair for several stage chains (incl. cluster shapes)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dpg_slam_b200 import synth
from dpg_slam_b200._abi import Params, COV_CENSI_CORR
from dpg_slam_b200.scanmatch import ScanMatcher
from gpu_probe import time_run
wl = synth.config_corridor(n_pairs=5000, seed=2)
p = Params.defaults(downsample_divisor=1, cov_mode=COV_CENSI_CORR)
k, its = 353, 305
for chain in sys.argv[1:]:
    os.environ["DPGICP_CHAIN"] = chain
    with ScanMatcher(0) as sm:
        sm.upload_ranges(wl.ranges, wl.scanner)
        for n in (1, 37):
            idx = np.full(n, k)
            sm.set_pairs(wl.src_idx[idx], wl.tgt_idx[idx], wl.guess[idx])
            best, med = time_run(sm, p, reps=3)
            print(f"chain={chain} copies={n}: {best:.3f} ms -> {1e3 * best / (its + 1):.2f} us per pass", flush=True)
