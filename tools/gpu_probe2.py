"""Development aid: time the ICP kernel for the library variant in DPGICP_LIBRARY over a few env settings."""
import os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from dpg_slam_b200 import synth
from dpg_slam_b200._abi import Params, COV_CENSI_CORR
from dpg_slam_b200.scanmatch import ScanMatcher
from gpu_probe import time_run

which, n_pairs = sys.argv[1], int(sys.argv[2])
chains = [a for a in sys.argv[3:]] or ["4,8,32"]
wl = synth.config_corridor(n_pairs=n_pairs, seed=2) if which == "corridor" else synth.config_loop_closure(n_pairs=n_pairs, n_scans=2000, seed=3)
tag = os.path.basename(os.environ.get("DPGICP_LIBRARY", "default")).replace(".so", "")
out = {}
for chain in chains:
    if chain == "default":
        os.environ.pop("DPGICP_CHAIN", None)       # the library's own stage chain
    else:
        os.environ["DPGICP_CHAIN"] = chain
    stages, warps, ctas = chain.replace(",", "-"), 0, 0
    with ScanMatcher(0) as sm:
        sm.upload_ranges(wl.ranges, wl.scanner)
        sm.set_pairs(wl.src_idx, wl.tgt_idx, wl.guess)
        for div in (1, 5):
            p = Params.defaults(downsample_divisor=div, cov_mode=COV_CENSI_CORR, search=int(os.environ.get("DPGICP_PROBE_SEARCH", "1")))
            best, med = time_run(sm, p, reps=5)
            c = sm.last_run_counters()
            key = f"st{stages}_w{warps}_c{ctas}_d{div}"
            out[key] = dict(ms=best, med=med, pairs_per_s=wl.n_pairs / best * 1e3, evals=c["distance_evals"], tests=c["box_tests"],
                            cands=c["dev_candidates"], loose=c["dev_loose_searches"], searches=c["dev_searches"], iters=c["iterations"])
            print(tag, key, json.dumps(out[key]), flush=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"probe2_{which}_{n_pairs}_{tag}.json"), "w"), indent=1)
