"""Small GPU case for compute-sanitizer (memcheck / racecheck / synccheck): every code path of the ICP kernel once — both
metrics, all covariance modes, the three searches, both outlier rejectors, staging with suspend/resume, the 4-CTA cluster
stage with its distributed-shared-memory exchange, down-sampling, ragged/empty scans — plus the factor kernel, the scan-store
builders, the device-resident pair enumeration (both callers, sharded), the seeded correspondence hook and the
two-context fused gather (peer stores from the epilogue)."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dpg_slam_b200 import synth  # noqa: E402
from dpg_slam_b200._abi import ENUM_ONLINE, ENUM_REOPTIMIZE, Params, load_library  # noqa: E402
from dpg_slam_b200.scanmatch import ScanMatcher  # noqa: E402

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
    for search in (0, 2):
        for mode, param in ((1, 0.7), (2, 1.5), (0, 0.0)):
            p = Params.defaults(downsample_divisor=1, cov_mode=2, search=search, outlier_mode=mode, outlier_param=param, max_iterations=8)
            sm.submit_pairs(wl.src_idx[:12], wl.tgt_idx[:12], wl.guess[:12], p)
    f = sm.fetch_factors()
    pts, off = sm.download_store()
    sm.upload_scans(pts, off)
    sm.submit_pairs(wl.src_idx[:3], wl.tgt_idx[:3], wl.guess[:3], Params.defaults(search=0, max_iterations=5))
    sm.enumerate_pairs(wl.poses_est[:, :2], wl.passes)
    S, T = pts[off[0]:off[1]], pts[off[1]:off[2]]
    sm.correspondences(S, T, [1, 0, 0, 0], Params.defaults())
    sm.correspondences_seeded(S, T, [1, 0, 0, 0], Params.defaults(outlier_mode=1, outlier_param=0.8), np.zeros(len(S), np.int32))
    sm.calculate_icp_cov(S, T, np.eye(4), Params.defaults(cov_mode=1))
    sm.run_icp(T, S, [0, 0, 0])
    # device-resident callers, sharded over two contexts of this process, records gathered by peer stores
    sm.upload_ranges(wl.ranges, wl.scanner)
    with ScanMatcher(0) as sm2:
        sm2.upload_ranges(wl.ranges, wl.scanner)
        for mode in (ENUM_REOPTIMIZE, ENUM_ONLINE):
            total = 0
            for rank, m in enumerate((sm, sm2)):
                m.set_nodes(wl.poses_est, wl.passes)
                total, _ = m.enumerate_pairs_device(mode, 5.0, 2.0, rank, 2)
            arr = (C.c_void_p * 2)(sm._h, sm2._h)
            assert load_library().dpgicp_gather_attach_local(arr, 2, total, 0) == 0
            for m in (sm, sm2):
                m.run(Params.defaults(cov_mode=2, max_iterations=10))
            for m in (sm, sm2):
                m.synchronize()
            g = sm.gather_fetch(total)
            for m in (sm, sm2):
                m.gather_detach()
# the staged chain with the cluster stage: 1081 beams, a chain that suspends after the first pass of the narrow stage
os.environ["DPGICP_CHAIN"] = "4,8,16x4"
wl2 = synth.config_corridor(n_pairs=12, seed=5)
with ScanMatcher(0) as sm:
    sm.upload_ranges(wl2.ranges, wl2.scanner)
    for metric, search, cov in ((0, 1, 2), (1, 1, 1), (0, 2, 2)):
        r2 = sm.submit_pairs(wl2.src_idx, wl2.tgt_idx, wl2.guess, Params.defaults(downsample_divisor=1, cov_mode=cov, metric=metric, search=search, max_iterations=6))
print("sanitize case done", int(r["iterations"].sum()), len(f), len(g), int(r2["iterations"].sum()))
