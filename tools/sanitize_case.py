"""Small GPU case for compute-sanitizer: every code path of the ICP kernel once (both metrics, all cov modes, staging,
down-sampling, ragged/empty scans), the factor kernel, scan-store builders and pair enumeration."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dpg_slam_b200 import synth
from dpg_slam_b200._abi import Params
from dpg_slam_b200.scanmatch import ScanMatcher

wl = synth.config_corridor(n_pairs=40, n_beams=361, seed=3)
wl.ranges[2, :] = 40.0
wl.ranges[4, 2:] = 40.0
with ScanMatcher(0) as sm:
    sm.upload_ranges(wl.ranges, wl.scanner)
    for metric in (0, 1):
        for cov in (0, 1, 2):
            for div in (1, 5):
                p = Params.defaults(downsample_divisor=div, cov_mode=cov, metric=metric, max_iterations=30)
                r = sm.submit_pairs(wl.src_idx, wl.tgt_idx, wl.guess, p)
    f = sm.fetch_factors()
    pts, off = sm.download_store()
    sm.upload_scans(pts, off)
    sm.submit_pairs(wl.src_idx[:3], wl.tgt_idx[:3], wl.guess[:3], Params.defaults(search=0, max_iterations=5))
    sm.enumerate_pairs(wl.poses_est[:, :2], wl.passes)
    sm.correspondences(pts[off[0]:off[1]], pts[off[1]:off[2]], [1, 0, 0, 0], Params.defaults())
    sm.calculate_icp_cov(pts[off[0]:off[1]], pts[off[1]:off[2]], np.eye(4), Params.defaults(cov_mode=1))
print("sanitize case done", int(r["iterations"].sum()), len(f))
