#!/usr/bin/env python
"""Oracle ICP loop vs the independent PCL emulation (tests/pcl_emulation.py) on larger samples of BASELINE configs 1-3;
writes profiles/r02_pcl_emulation_report.json.  CPU only (build container)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
from dpg_slam_b200 import synth  # noqa: E402
import test_pcl_emulation as T  # noqa: E402

cases = [
    ("config2 corridor, divisor 5 (reference default)", synth.config_corridor(n_pairs=400, seed=2), 5, 200),
    ("config2 corridor, divisor 1 (benchmark setting)", synth.config_corridor(n_pairs=400, seed=2), 1, 48),
    ("config3 loop closure, divisor 5", synth.config_loop_closure(n_pairs=400, n_scans=200, seed=3), 5, 200),
    ("config3 loop closure, divisor 1", synth.config_loop_closure(n_pairs=400, n_scans=200, seed=3), 1, 48),
]
out = {"what": "oracle (oracle/dpg_oracle.c) vs independent whole-loop PCL emulation (tests/pcl_emulation.py); PCL itself is absent",
       "cases": []}
engines = [("scipy", "lapack")]
try:
    import cv2  # noqa: F401
    engines.append(("flann", "opencv"))       # PCL's own neighbour library + a float32 Jacobi SVD (Eigen's family)
    engines.append(("flann", "eigen_jacobi")) # ... + Eigen 3.3's two-sided Jacobi algorithm restated in float32
except ImportError:
    pass
for (name0, wl, div, n), (nn, svd) in [(c, e) for c in cases for e in engines]:
    name = f"{name0} [nn={nn}, svd={svd}]"
    rows = T.compare_pairs(wl, div, n, nn=nn, svd=svd)
    s = T.summarize(rows)
    s["name"] = name
    s["iterations_oracle_hist"] = np.percentile([r["it_oracle"] for r in rows], [50, 90, 99, 100]).tolist()
    s["iterations_emulation_hist"] = np.percentile([r["it_emu"] for r in rows], [50, 90, 99, 100]).tolist()
    s["stop_emulation"] = {k: sum(r["stop_emu"] == k for r in rows) for k in ("transform", "abs_mse", "iterations", "no_correspondences")}
    out["cases"].append(s)
    print(json.dumps(s))
with open(os.path.join(ROOT, "profiles", "r02_pcl_emulation_report.json"), "w") as f:
    json.dump(out, f, indent=1)
