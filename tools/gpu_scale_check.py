"""One-off: a 1M-pair batch (reference down-sampling) end to end: submit, factors, spot parity against the oracle."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dpg_slam_b200 import synth
from dpg_slam_b200._abi import Params, COV_CENSI_CORR
from dpg_slam_b200.scanmatch import ScanMatcher
from oracle import oracle_py as O
wl = synth.config_loop_closure(n_pairs=1_000_000, n_scans=4000, seed=3)
p = Params.defaults(cov_mode=COV_CENSI_CORR)                       # divisor 5: the reference's default
with ScanMatcher(0) as sm:
    sm.upload_ranges(wl.ranges, wl.scanner)
    t = time.time(); rec = sm.submit_pairs(wl.src_idx, wl.tgt_idx, wl.guess, p); dt = time.time() - t
    fac = sm.fetch_factors()
    pts, off = sm.download_store()
idx = np.arange(0, 1_000_000, 9973)
ref, _ = O.run_batch(pts, off, wl.src_idx[idx], wl.tgt_idx[idx], wl.guess[idx], p, fast=1, threads=0)
ok = all(np.array_equal(rec[idx][f], ref[f]) for f in ("tx", "ty", "iterations", "status", "n_correspondences", "mse"))
print(f"1M pairs in {dt:.2f} s ({1e6/dt:.0f} pairs/s incl. H2D/D2H), sample of {len(idx)} equals oracle: {ok}, "
      f"converged {(rec['status'] & 0x100 != 0).mean():.3f}, factors ok {(fac['status'] & 0x800 == 0).mean():.3f}")
