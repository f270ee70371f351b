"""Development aid: latency of the runIcp-shaped single call (dpgicp_single_pair) and how many kernels it launches."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dpg_slam_b200 import synth
from dpg_slam_b200._abi import Params, COV_CENSI_CORR
from dpg_slam_b200.scanmatch import ScanMatcher, relative_guess
from oracle import oracle_py as O

wl = synth.config_corridor(n_pairs=40, seed=5)
pts, off = O.clouds_from_ranges(wl.ranges, wl.scanner)
with ScanMatcher(0) as sm:
    for tag, p in (("div5 live", Params.defaults()), ("div1 censi", Params.defaults(downsample_divisor=1, cov_mode=COV_CENSI_CORR))):
        ms, launches, iters = [], [], []
        for k in range(40):
            s, t = int(wl.src_idx[k]), int(wl.tgt_idx[k])
            S, T = pts[off[s]:off[s + 1]], pts[off[t]:off[t + 1]]
            sm.run_icp(T, S, wl.guess[k], p)
            l0 = sm.last_run_counters()["kernel_launches"]
            t0 = time.perf_counter()
            _, _, _, r = sm.run_icp(T, S, wl.guess[k], p)
            ms.append(1e3 * (time.perf_counter() - t0))
            launches.append(sm.last_run_counters()["kernel_launches"] - l0)
            iters.append(r.iterations)
        print(f"{tag}: median {np.median(ms):.3f} ms  p10 {np.percentile(ms, 10):.3f}  launches/call {np.median(launches):.0f}  "
              f"mean iterations {np.mean(iters):.1f}  ms per iteration {np.median(np.array(ms) / np.maximum(iters, 1)) * 1e3:.2f} us")
