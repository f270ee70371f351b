"""ctypes wrapper of the CPU oracle (oracle/dpg_oracle.c).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package (dpg_slam_b200) never imports this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))
from dpg_slam_b200._abi import FACTOR_DTYPE, Params, Result, RESULT_DTYPE  # noqa: E402  (shared POD types only)

_lib = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libdpgoracle.so")
    src = [os.path.join(_HERE, f) for f in ("dpg_oracle.c", "dpg_oracle.h")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src):
        subprocess.run(["make", "-C", _HERE, "libdpgoracle.so"], check=True, capture_output=True)
    return so


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        vp, i32, i64, f = C.c_void_p, C.c_int32, C.c_int64, C.c_float
        PP = C.POINTER(Params)
        L.orc_angle_mod.restype = f
        L.orc_angle_mod.argtypes = [f]
        L.orc_relative_guess.restype = None
        L.orc_relative_guess.argtypes = [vp, f, vp, f, vp]
        L.orc_guess_matrix.restype = None
        L.orc_guess_matrix.argtypes = [vp, vp]
        L.orc_ranges_to_cloud.restype = C.c_int
        L.orc_ranges_to_cloud.argtypes = [vp, C.c_int, f, f, f, f, f, f, vp]
        L.orc_downsample.restype = C.c_int
        L.orc_downsample.argtypes = [vp, C.c_int, C.c_int, vp]
        L.orc_transform_points.restype = None
        L.orc_transform_points.argtypes = [vp, vp, C.c_int, vp]
        L.orc_correspondences.restype = C.c_int
        L.orc_correspondences.argtypes = [vp, C.c_int, vp, C.c_int, PP, C.c_int, vp, vp]
        L.orc_correspondences_ex.restype = C.c_int
        L.orc_correspondences_ex.argtypes = [vp, C.c_int, vp, C.c_int, PP, C.c_int, vp, vp, vp, vp]
        L.orc_correspondences_seeded.restype = C.c_int
        L.orc_correspondences_seeded.argtypes = [vp, C.c_int, vp, C.c_int, PP, C.c_int, vp, vp, vp, vp, vp]
        L.orc_outlier_threshold.restype = f
        L.orc_outlier_threshold.argtypes = [vp, C.c_int, PP]
        L.orc_enumerate_online.restype = i64
        L.orc_enumerate_online.argtypes = [vp, vp, C.c_int, f, f, vp, vp, i64]
        L.orc_beam_key.restype = f
        L.orc_beam_key.argtypes = [f, f, f, f]
        L.orc_icp.restype = None
        L.orc_icp.argtypes = [vp, C.c_int, vp, C.c_int, vp, PP, C.c_int, C.POINTER(Result), vp]
        L.orc_cov_censi.restype = C.c_uint32
        L.orc_cov_censi.argtypes = [vp, vp, C.c_int, C.c_int, vp, C.c_double, vp, vp, vp]
        L.orc_run_pair.restype = None
        L.orc_run_pair.argtypes = [vp, C.c_int, vp, C.c_int, vp, PP, C.c_int, C.POINTER(Result)]
        L.orc_run_batch.restype = C.c_int
        L.orc_run_batch.argtypes = [vp, vp, C.c_int, vp, vp, vp, i64, PP, C.c_int, C.c_int, vp]
        L.orc_factor.restype = None
        L.orc_factor.argtypes = [vp, i32, i32, vp]
        L.orc_enumerate_pairs.restype = i64
        L.orc_enumerate_pairs.argtypes = [vp, vp, C.c_int, f, f, vp, vp, i64]
        _lib = L
    return _lib


class _Trace(C.Structure):
    _fields_ = [("capacity", C.c_int), ("count", C.c_int), ("T_iter", C.c_void_p), ("n_corr", C.c_void_p)]


def _f32(a):
    return np.ascontiguousarray(a, np.float32)


def angle_mod(a: float) -> float:
    return float(lib().orc_angle_mod(a))


def relative_guess(p1, th1, p2, th2) -> np.ndarray:
    out = np.zeros(3, np.float32)
    a, b = _f32(p1), _f32(p2)
    lib().orc_relative_guess(a.ctypes.data, th1, b.ctypes.data, th2, out.ctypes.data)
    return out


def guess_matrix(guess) -> np.ndarray:
    g = _f32(guess)
    T = np.zeros(4, np.float32)
    lib().orc_guess_matrix(g.ctypes.data, T.ctypes.data)
    return T


def ranges_to_cloud(ranges, scanner) -> np.ndarray:
    r = _f32(ranges)
    out = np.zeros((r.shape[0], 2), np.float32)
    n = lib().orc_ranges_to_cloud(r.ctypes.data, r.shape[0], scanner.angle_min, scanner.angle_max,
                                  scanner.range_max, scanner.laser_x, scanner.laser_y,
                                  scanner.laser_theta, out.ctypes.data)
    return np.ascontiguousarray(out[:n])


def clouds_from_ranges(ranges2d, scanner):
    """-> (points (N,2) float32, offsets (n_scans+1,) int64): the CSR scan store"""
    clouds = [ranges_to_cloud(r, scanner) for r in ranges2d]
    offsets = np.zeros(len(clouds) + 1, np.int64)
    offsets[1:] = np.cumsum([c.shape[0] for c in clouds])
    pts = np.concatenate(clouds) if clouds else np.zeros((0, 2), np.float32)
    return np.ascontiguousarray(pts, np.float32), offsets


def downsample(xy, divisor) -> np.ndarray:
    a = _f32(xy)
    out = np.zeros_like(a)
    n = lib().orc_downsample(a.ctypes.data, a.shape[0], divisor, out.ctypes.data)
    return np.ascontiguousarray(out[:n])


def transform_points(T, xy) -> np.ndarray:
    a, t = _f32(xy), _f32(T)
    out = np.empty_like(a)
    lib().orc_transform_points(t.ctypes.data, a.ctypes.data, a.shape[0], out.ctypes.data)
    return out


def correspondences(src_t, tgt, params: Params, fast=0, src_orig=None, T=None, prev_nn=None):
    """One correspondence pass.  ``src_orig`` / ``T`` (the untransformed source and the (c, s, tx, ty)
    that produced ``src_t``) are only read by SEARCH_PROJECTIVE.  ``prev_nn`` (int32 per source point, -1 = none) is the
    sticky tie preference of an ICP run in progress; when given, a 4th value is returned: this pass's gated forward
    neighbours (the next pass's ``prev_nn``)."""
    s, t = _f32(src_t), _f32(tgt)
    corr = np.full(s.shape[0], -1, np.int32)
    d2 = np.zeros(max(s.shape[0], 1), np.float32)
    so = _f32(src_orig) if src_orig is not None else None
    tt = _f32(T) if T is not None else None
    nn = np.ascontiguousarray(prev_nn, np.int32).copy() if prev_nn is not None else None
    k = lib().orc_correspondences_seeded(s.ctypes.data, s.shape[0], t.ctypes.data, t.shape[0],
                                         C.byref(params), fast, so.ctypes.data if so is not None else None,
                                         tt.ctypes.data if tt is not None else None,
                                         nn.ctypes.data if nn is not None else None, corr.ctypes.data, d2.ctypes.data)
    d2 = d2[:s.shape[0]]
    if nn is not None:
        return k, corr, d2, nn
    return k, corr, d2


def outlier_threshold(d2_accepted, params: Params) -> float:
    d = _f32(d2_accepted)
    return float(lib().orc_outlier_threshold(d.ctypes.data, d.shape[0], C.byref(params)))


def icp(src, tgt, guess, params: Params, fast=0, trace=False):
    s, t, g = _f32(src), _f32(tgt), _f32(guess)
    res = Result()
    tr = None
    if trace:
        cap = params.max_iterations + 2
        T_iter = np.zeros((cap, 4), np.float32)
        n_corr = np.zeros(cap, np.int32)
        tr = _Trace(cap, 0, T_iter.ctypes.data, n_corr.ctypes.data)
    lib().orc_icp(s.ctypes.data, s.shape[0], t.ctypes.data, t.shape[0], g.ctypes.data,
                  C.byref(params), fast, C.byref(res), C.byref(tr) if tr is not None else None)
    if trace:
        return res, T_iter[:tr.count].copy(), n_corr[:tr.count].copy()
    return res


def cov_censi(P, Q, n_h, n_d, T, sensor_var=0.01, live=(0.5, 0.5, 0.3)):
    p, q, t, lv = _f32(P), _f32(Q), _f32(T), _f32(live)
    cov = np.zeros(9, np.float64)
    H = np.zeros(9, np.float64)
    st = lib().orc_cov_censi(p.ctypes.data, q.ctypes.data, n_h, n_d, t.ctypes.data, sensor_var,
                             lv.ctypes.data, cov.ctypes.data, H.ctypes.data)
    return int(st), cov.reshape(3, 3), H.reshape(3, 3)


def run_pair(source_full, target_full, guess, params: Params, fast=0) -> Result:
    s, t, g = _f32(source_full), _f32(target_full), _f32(guess)
    res = Result()
    lib().orc_run_pair(s.ctypes.data, s.shape[0], t.ctypes.data, t.shape[0], g.ctypes.data,
                       C.byref(params), fast, C.byref(res))
    return res


def run_batch(points, offsets, src_idx, tgt_idx, guess, params: Params, fast=1, threads=1):
    """-> (records structured array, threads used)"""
    pts = _f32(points)
    off = np.ascontiguousarray(offsets, np.int64)
    s = np.ascontiguousarray(src_idx, np.int32)
    t = np.ascontiguousarray(tgt_idx, np.int32)
    g = _f32(guess)
    out = np.zeros(s.shape[0], RESULT_DTYPE)
    used = lib().orc_run_batch(pts.ctypes.data, off.ctypes.data, off.shape[0] - 1, s.ctypes.data,
                               t.ctypes.data, g.ctypes.data, s.shape[0], C.byref(params), fast,
                               threads, out.ctypes.data)
    return out, used


def enumerate_pairs(node_xy, node_pass, same_radius, other_radius):
    xy = _f32(node_xy)
    ps = np.ascontiguousarray(node_pass, np.int32)
    n = lib().orc_enumerate_pairs(xy.ctypes.data, ps.ctypes.data, xy.shape[0], same_radius,
                                  other_radius, None, None, 0)
    src = np.zeros(n, np.int32)
    tgt = np.zeros(n, np.int32)
    lib().orc_enumerate_pairs(xy.ctypes.data, ps.ctypes.data, xy.shape[0], same_radius, other_radius,
                              src.ctypes.data, tgt.ctypes.data, n)
    return src, tgt


def enumerate_online(node_xy, node_pass, same_radius, other_radius):
    """Pair list of one updatePoseGraphObsConstraints call for the newest node (dpg_slam.cc:255-300)."""
    xy = _f32(node_xy)
    ps = np.ascontiguousarray(node_pass, np.int32)
    n = lib().orc_enumerate_online(xy.ctypes.data, ps.ctypes.data, xy.shape[0], same_radius, other_radius, None, None, 0)
    src = np.zeros(n, np.int32)
    tgt = np.zeros(n, np.int32)
    lib().orc_enumerate_online(xy.ctypes.data, ps.ctypes.data, xy.shape[0], same_radius, other_radius,
                               src.ctypes.data, tgt.ctypes.data, n)
    return src, tgt


def factors(records, src_idx, tgt_idx):
    """records (RESULT_DTYPE array) -> FACTOR_DTYPE array (orc_factor per pair)"""
    rec = np.ascontiguousarray(records)
    out = np.zeros(rec.shape[0], FACTOR_DTYPE)
    for k in range(rec.shape[0]):
        lib().orc_factor(rec[k:k + 1].ctypes.data, int(src_idx[k]), int(tgt_idx[k]), out[k:k + 1].ctypes.data)
    return out
