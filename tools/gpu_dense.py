"""Development aid: dense (4096-beam, point-to-line) timing for several chains."""
import os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dpg_slam_b200 import synth
from dpg_slam_b200._abi import Params, COV_CENSI_CORR
from dpg_slam_b200.scanmatch import ScanMatcher
from gpu_probe import time_run
wl = synth.config_loop_closure(n_pairs=20000, n_scans=2000, n_beams=4096, seed=4)
for chain in sys.argv[1:]:
    if chain == "default": os.environ.pop("DPGICP_CHAIN", None)
    else: os.environ["DPGICP_CHAIN"] = chain
    with ScanMatcher(0) as sm:
        sm.upload_ranges(wl.ranges, wl.scanner)
        sm.set_pairs(wl.src_idx, wl.tgt_idx, wl.guess)
        for metric in (1, 0):
            p = Params.defaults(downsample_divisor=1, cov_mode=COV_CENSI_CORR, metric=metric)
            best, med = time_run(sm, p, reps=2)
            c = sm.last_run_counters()
            print(f"chain={chain} metric={metric}: {best:.1f} ms -> {wl.n_pairs/best*1e3:.0f} pairs/s evals {c['distance_evals']/1e9:.0f}G tests {c['box_tests']/1e9:.1f}G iters {c['iterations']}", flush=True)
