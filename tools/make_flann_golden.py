"""Generates tests/golden/flann_nn.json: what a REAL FLANN single kd-tree — OpenCV's bundled copy of the library
pcl::KdTreeFLANN wraps (flann::KDTreeSingleIndex, leaf size 15, L2 on float32 3-D points, checks = -1, eps = 0: PCL's exact
search) — returns for the forward and the reciprocal neighbour queries of PCL's determineReciprocalCorrespondences at several
iterates of sampled BASELINE config-2 / config-3 pairs (divisor 5: 217-point clouds, the reference's own setting).

The clouds and iterates are stored too (exact bit patterns), so that the fixture pins the oracle AND the CUDA path without
OpenCV, the generator or the workload code:   python tools/make_flann_golden.py     (needs cv2; any container of this image)
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import pcl_emulation as E  # noqa: E402
from dpg_slam_b200 import synth  # noqa: E402
from dpg_slam_b200._abi import Params  # noqa: E402
from oracle import oracle_py as O  # noqa: E402


def f32hex(a):
    return [format(int(v), "08x") for v in np.ascontiguousarray(a, np.float32).reshape(-1).view(np.uint32)]


def xyz(a):
    out = np.zeros((len(a), 3), np.float32)
    out[:, :2] = a
    return out


cases = []
for name, wl in (("corridor", synth.config_corridor(n_pairs=12, seed=51)),
                 ("loop_closure", synth.config_loop_closure(n_pairs=12, n_scans=30, seed=52))):
    pts, off = O.clouds_from_ranges(wl.ranges, wl.scanner)
    for k in (0, 5, 11):
        s, t = int(wl.src_idx[k]), int(wl.tgt_idx[k])
        S, T = pts[off[s]:off[s + 1]][::5], pts[off[t]:off[t + 1]][::5]
        _, iterates, _ = O.icp(S, T, wl.guess[k], Params.defaults(downsample_divisor=1), trace=True)
        tree_t = E.FlannTree(xyz(T))
        for it in sorted(set([0, len(iterates) // 2, len(iterates) - 1])):
            Tm = np.asarray(iterates[it], np.float32)
            cur = O.transform_points(Tm, S)
            d2f, jf = tree_t.query(xyz(cur))
            _, back = E.FlannTree(xyz(cur)).query(xyz(T)[jf])
            cases.append({"workload": name, "pair": k, "iterate": int(it), "T_hex": f32hex(Tm),
                          "source_hex": f32hex(S), "target_hex": f32hex(T),
                          "flann_forward_index": [int(v) for v in jf], "flann_forward_d2_hex": f32hex(d2f),
                          "flann_backward_index": [int(v) for v in back]})
out = {"what": "cv2.flann.Index(algorithm=FLANN_INDEX_KDTREE_SINGLE, leaf_max_size=15).knnSearch(k=1, checks=-1, eps=0) on float32 "
               "(x, y, 0) points: forward = nearest target of every transformed source point, backward = nearest transformed "
               "source point of that target point (PCL determineReciprocalCorrespondences, SURVEY App. A.3-2)",
       "generator": "tools/make_flann_golden.py", "opencv": __import__("cv2").__version__, "cases": cases}
with open(os.path.join(ROOT, "tests", "golden", "flann_nn.json"), "w") as f:
    json.dump(out, f)
print(len(cases), "cases,", os.path.getsize(os.path.join(ROOT, "tests", "golden", "flann_nn.json")), "bytes")
