"""Development aid: per-phase cycles of a pass for ONE pair (thread 0's view), per CTA width. Needs libdpgicp_phase.so."""
import os, sys, ctypes as C
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["DPGICP_LIBRARY"] = os.path.join(ROOT, "dpg_slam_b200", "libdpgicp_phase.so")
from dpg_slam_b200 import synth
from dpg_slam_b200._abi import Params, COV_CENSI_CORR
from dpg_slam_b200.scanmatch import ScanMatcher
wl = synth.config_corridor(n_pairs=5000, seed=2)
p = Params.defaults(downsample_divisor=1, cov_mode=COV_CENSI_CORR)
k, its = 353, 305
names = ["tiles", "reduce+atomics", "wait A", "solve", "transform+boxes", "wait C"]
for chain in ("12", "17"):
    os.environ["DPGICP_CHAIN"] = chain
    with ScanMatcher(0) as sm:
        sm.upload_ranges(wl.ranges, wl.scanner)
        idx = np.full(1, k)
        sm.submit_pairs(wl.src_idx[idx], wl.tgt_idx[idx], wl.guess[idx], p)
        c = (C.c_uint64 * 8)()
        sm._lib.dpgicp_last_run_counters(sm._h, C.byref(c))
        # phase counters live right after the 8 public counters
        import ctypes
        buf = (C.c_uint64 * 8)()
        sm._lib.dpgicp_debug_phase_counters(sm._h, C.byref(buf))
        tot = sum(buf[i] for i in range(6))
        print(f"W={chain}: " + ", ".join(f"{n} {buf[i]/its:.0f}" for i, n in enumerate(names)) + f" | total {tot/its:.0f} cycles/pass", flush=True)
