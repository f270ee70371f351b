"""One-off confidence run: thousands of pairs, GPU vs oracle, every record field (bit-exact where the contract says so)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from dpg_slam_b200 import synth
from dpg_slam_b200._abi import Params
from dpg_slam_b200.scanmatch import ScanMatcher
from oracle import oracle_py as O
EXACT = ("tx", "ty", "rot_c", "rot_s", "iterations", "status", "n_correspondences", "mse")
with ScanMatcher(0) as sm:
    for name, wl in (("loop", synth.config_loop_closure(n_pairs=4000, n_scans=600, seed=77)), ("corridor", synth.config_corridor(n_pairs=3000, seed=78)),
                     ("dense", synth.config_loop_closure(n_pairs=300, n_scans=100, n_beams=4096, seed=79))):
        sm.upload_ranges(wl.ranges, wl.scanner)
        pts, off = sm.download_store()
        for metric in (0, 1):
            for div, cov in ((1, 2), (5, 1)):
                p = Params.defaults(downsample_divisor=div, cov_mode=cov, metric=metric)
                t = time.time(); got = sm.submit_pairs(wl.src_idx, wl.tgt_idx, wl.guess, p); tg = time.time() - t
                t = time.time(); ref, _ = O.run_batch(pts, off, wl.src_idx, wl.tgt_idx, wl.guess, p, fast=1, threads=0); tc = time.time() - t
                bad = {f: int((got[f] != ref[f]).sum()) for f in EXACT}
                dth = np.abs(got["theta"] - ref["theta"]).max()
                scale = np.abs(ref["cov"]).max(axis=1)
                rel = (np.abs(got["cov"] - ref["cov"]).max(axis=1) / np.where(scale > 0, scale, 1)).max()
                print(f"{name} metric={metric} div={div} cov={cov}: mismatches {bad} max dtheta {dth:.2e} max cov rel {rel:.2e}  gpu {tg:.2f}s cpu {tc:.1f}s", flush=True)
