"""Generates tests/golden/*.json from the REFERENCE'S OWN code compiled by oracle/build_ref.sh
(oracle/_ref/libdpgref.so = /root/reference/src/icp_cov/cov_func_point_to_point.h and
/root/reference/src/dpg_slam/math_utils.cc built against oracle/ref_stubs/).

Run in the build container only (needs /root/reference):   python tools/make_golden.py
The fixtures are what travels to the GPU box; /root/reference does not.
Floats are stored as hex bit patterns (exact) next to readable decimals.
"""
import ctypes as C
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def ref():
    so = os.path.join(ROOT, "oracle", "_ref", "libdpgref.so")
    subprocess.run(["bash", os.path.join(ROOT, "oracle", "build_ref.sh")], check=True, capture_output=True)
    L = C.CDLL(so)
    vp = C.c_void_p
    L.ref_cov_live.restype = C.c_int
    L.ref_cov_live.argtypes = [vp, C.c_int, vp, C.c_int, vp, C.c_float, C.c_float, C.c_float, vp, vp, vp, C.c_long, vp]
    L.ref_cov_intended.restype = C.c_int
    L.ref_cov_intended.argtypes = [vp, C.c_int, vp, C.c_int, vp, C.c_double, vp, vp]
    L.ref_angle_mod.restype = C.c_float
    L.ref_angle_mod.argtypes = [C.c_float]
    L.ref_inverse_transform_point.restype = None
    L.ref_inverse_transform_point.argtypes = [vp, C.c_float, vp, C.c_float, vp]
    L.ref_transform_point.restype = None
    L.ref_transform_point.argtypes = [vp, C.c_float, vp, C.c_float, vp]
    return L


def f32hex(a):
    return [format(int(v), "08x") for v in np.ascontiguousarray(a, np.float32).reshape(-1).view(np.uint32)]


def f64hex(a):
    return [format(int(v), "016x") for v in np.ascontiguousarray(a, np.float64).reshape(-1).view(np.uint64)]


def transform(theta, tx, ty):
    th = np.float32(theta)
    c, s = np.float32(np.cos(np.float64(th))), np.float32(np.sin(np.float64(th)))
    T = np.eye(4, dtype=np.float32)
    T[0, 0], T[0, 1], T[1, 0], T[1, 1] = c, -s, s, c
    T[0, 3], T[1, 3] = np.float32(tx), np.float32(ty)
    return T


def cov_case(L, name, P, Q, T, sensor_var=0.01, live=(0.5, 0.5, 0.3)):
    P = np.ascontiguousarray(P, np.float32)
    Q = np.ascontiguousarray(Q, np.float32)
    Tc = np.ascontiguousarray(T.T, np.float32).reshape(16)
    cov, H3 = np.zeros(9), np.zeros(9)
    n = L.ref_cov_intended(P.ctypes.data, len(P), Q.ctypes.data, len(Q), Tc.ctypes.data, sensor_var,
                           cov.ctypes.data, H3.ctypes.data)
    livec, H6 = np.zeros(9), np.zeros(36)
    z = C.c_int(0)
    n2 = L.ref_cov_live(P.ctypes.data, len(P), Q.ctypes.data, len(Q), Tc.ctypes.data, live[0], live[1], live[2],
                        livec.ctypes.data, H6.ctypes.data, None, 0, C.byref(z))
    assert n2 == min(len(P), 200), (n, n2)
    assert z.value == 6 * n2
    singular = n == -2               # the 6x6 d2J_dX2 is not invertible (e.g. a single point pair)
    assert singular or n == n2
    n = n2
    return dict(name=name, n_data=len(P), n_model=len(Q), n_d_used=n, singular=bool(singular), cov_z_dim=z.value, sensor_var=sensor_var,
                live_in=list(live), P_hex=f32hex(P), Q_hex=f32hex(Q), T_colmajor_hex=f32hex(Tc),
                live_cov_hex=f64hex(livec), cov3_hex=f64hex(cov), H3_hex=f64hex(H3), H6_colmajor_hex=f64hex(H6),
                cov3=cov.tolist(), H3=H3.tolist())


def main():
    os.makedirs(GOLD, exist_ok=True)
    L = ref()
    cases = []
    # KAT-1: 5 explicit points
    P = [(1.0, 0.5), (2.0, -1.0), (3.5, 0.25), (0.5, 2.0), (-1.0, 1.5)]
    Q = [(1.2550874948501587, 0.3773355185985565), (2.3748416900634766, -0.9903373122215271),
         (3.7775561809539795, 0.4081680178642273), (0.5928352475166321, 1.8299250602722168),
         (-0.8447542786598206, 1.2076728343963623)]
    T = transform(0.1, 0.3, -0.2)
    cases.append(cov_case(L, "kat1", P, Q, T))
    # KAT-2: n = 300 > 200 exercises the truncation quirk (cov.h:307)
    rng = np.random.default_rng(12345)
    ang = np.linspace(-2.356, 2.356, 300)
    r = rng.uniform(1, 10, 300)
    P = np.stack([r * np.cos(ang), r * np.sin(ang)], 1).astype(np.float32)
    c, s = T[0, 0], T[1, 0]
    Q = (np.stack([c * P[:, 0] - s * P[:, 1] + T[0, 3], s * P[:, 0] + c * P[:, 1] + T[1, 3]], 1)
         + rng.normal(0, 0.01, (300, 2))).astype(np.float32)
    cases.append(cov_case(L, "kat2_cap200", P, Q, T))
    # assorted sizes / poses, including model longer than data, a single point pair and exactly 200 / 201
    rng = np.random.default_rng(2024)
    for k, (n, extra, theta, tx, ty, sig) in enumerate([
            (1, 0, 0.0, 0.0, 0.0, 0.0), (3, 2, -0.7, 1.5, -2.0, 0.02), (17, 0, 2.9, -4.0, 0.3, 0.05),
            (200, 0, 0.05, 0.1, 0.1, 0.01), (201, 5, -0.05, -0.1, 0.2, 0.01), (1081, 0, 0.3, 0.5, -0.5, 0.01),
            (1081, 0, -3.1, 10.0, -7.0, 0.03), (64, 0, 1.2, 0.0, 0.0, 0.3)]):
        ang = np.sort(rng.uniform(-2.356, 2.356, n))
        rr = rng.uniform(0.5, 25.0, n)
        P = np.stack([rr * np.cos(ang) + 0.2, rr * np.sin(ang)], 1).astype(np.float32)
        Tk = transform(theta, tx, ty)
        c, s = Tk[0, 0], Tk[1, 0]
        Qm = np.stack([c * P[:, 0] - s * P[:, 1] + Tk[0, 3], s * P[:, 0] + c * P[:, 1] + Tk[1, 3]], 1)
        Qm = Qm + rng.normal(0, sig, Qm.shape) if sig > 0 else Qm
        if extra:
            Qm = np.concatenate([Qm, rng.uniform(-5, 5, (extra, 2))])
        cases.append(cov_case(L, f"rand{k}_n{n}", P, Qm.astype(np.float32), Tk,
                              live=(0.5, 0.5, 0.3) if k % 2 == 0 else (0.25, 0.125, 0.0625)))
    with open(os.path.join(GOLD, "cov_ref.json"), "w") as f:
        json.dump(dict(source="/root/reference/src/icp_cov/cov_func_point_to_point.h compiled by oracle/build_ref.sh",
                       cases=cases), f, indent=0)

    # math_utils: AngleMod<float>, inverseTransformPoint (the runIcp guess, dpg_slam.cc:364-370), transformPoint
    rng = np.random.default_rng(7)
    am_in = np.concatenate([np.array([0, 3.2, -3.2, 7, 100, -0.5, np.pi, -np.pi, 2 * np.pi, 1e-8, 6.2831855, -9.42477796],
                                     np.float32), rng.uniform(-50, 50, 200).astype(np.float32)])
    am_out = np.array([L.ref_angle_mod(float(a)) for a in am_in], np.float32)
    inv_in, inv_out, fwd_out = [], [], []
    for _ in range(300):
        p1 = rng.uniform(-100, 100, 2).astype(np.float32); th1 = np.float32(rng.uniform(-6.5, 6.5))
        p2 = (p1 + rng.uniform(-3, 3, 2)).astype(np.float32); th2 = np.float32(rng.uniform(-6.5, 6.5))
        o = np.zeros(3, np.float32)
        L.ref_inverse_transform_point(p2.ctypes.data, th2, p1.ctypes.data, th1, o.ctypes.data)   # node_2 in node_1
        o2 = np.zeros(3, np.float32)
        L.ref_transform_point(p2.ctypes.data, th2, p1.ctypes.data, th1, o2.ctypes.data)
        inv_in.append(np.array([p1[0], p1[1], th1, p2[0], p2[1], th2], np.float32))
        inv_out.append(o); fwd_out.append(o2)
    with open(os.path.join(GOLD, "math_utils_ref.json"), "w") as f:
        json.dump(dict(source="/root/reference/src/dpg_slam/math_utils.{h,cc} compiled by oracle/build_ref.sh",
                       angle_mod_in_hex=f32hex(am_in), angle_mod_out_hex=f32hex(am_out),
                       pose_pairs_hex=f32hex(np.stack(inv_in)), inverse_transform_out_hex=f32hex(np.stack(inv_out)),
                       transform_out_hex=f32hex(np.stack(fwd_out))), f, indent=0)
    print("wrote", os.listdir(GOLD))


if __name__ == "__main__":
    sys.exit(main())
