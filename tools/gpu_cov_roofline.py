"""Development aid: HBM roofline of the standalone batched covariance kernel (dpgicp_cov_pairs) on a store larger than L2."""
import os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dpg_slam_b200 import synth
from dpg_slam_b200._abi import Params, COV_CENSI_INDEXPAIR
from dpg_slam_b200.scanmatch import ScanMatcher
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
wl = synth.config_corridor(n_pairs=40000, seed=2)                   # 40001 scans x 1081 beams: 346 MB of points > 126 MB L2
p = Params.defaults(cov_mode=COV_CENSI_INDEXPAIR)
with ScanMatcher(0) as sm:
    sm.upload_ranges(wl.ranges, wl.scanner)
    counts = (wl.ranges < wl.scanner.range_max).sum(axis=1)
    rng = np.random.default_rng(0)
    n = 400000
    a = rng.integers(0, wl.n_scans, n).astype(np.int32); b = rng.integers(0, wl.n_scans, n).astype(np.int32)
    T = np.tile(np.array([1, 0, 0.5, 0.1], np.float32), (n, 1))
    best = 1e9
    for rep in range(4):
        cov, st, ms = sm.calculate_icp_cov_pairs(a, b, T, p)
        best = min(best, ms)
    nh = np.minimum(counts[a], counts[b])
    bytes_alg = float((16 * nh).sum() + n * (40 + 76))
    print(json.dumps({"kernel": "cov_indexpair_kernel<1> (one warp per item)", "items": n, "ms": best, "algorithmic_bytes": bytes_alg,
                      "achieved_gbs": bytes_alg / (best * 1e-3) / 1e9, "peak_gbs": peak, "frac": bytes_alg / (best * 1e-3) / 1e9 / peak,
                      "items_per_s": n / (best * 1e-3)}))
