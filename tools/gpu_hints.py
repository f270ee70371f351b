"""Development aid: effect of per-pair cost hints (last alignment's iteration counts) on the step time."""
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
with ScanMatcher(0) as sm:
    sm.upload_ranges(wl.ranges, wl.scanner)
    rec = sm.submit_pairs(wl.src_idx, wl.tgt_idx, wl.guess, p)
    sm.set_pairs(wl.src_idx, wl.tgt_idx, wl.guess)
    print("no hints           : %.2f ms" % time_run(sm, p, reps=5)[0])
    sm.set_pair_cost_hints(rec["iterations"])
    print("exact hints        : %.2f ms" % time_run(sm, p, reps=5)[0])
    # a re-alignment of the same pairs from slightly different estimates: hints from the previous run are approximate
    g2 = wl.guess + np.random.default_rng(3).normal(0, 0.01, wl.guess.shape).astype(np.float32)
    sm.set_pairs(wl.src_idx, wl.tgt_idx, g2)
    print("new guesses, none  : %.2f ms" % time_run(sm, p, reps=5)[0])
    sm.set_pair_cost_hints(rec["iterations"])
    print("new guesses, stale : %.2f ms" % time_run(sm, p, reps=5)[0])
    rec2 = sm.fetch_results()
    print("corr(iterations old, new) = %.3f" % np.corrcoef(rec["iterations"], rec2["iterations"])[0, 1])
