"""CPU: the oracle (oracle/dpg_oracle.c) against the golden vectors produced from the reference's
own compiled code (tools/make_golden.py -> tests/golden/), SURVEY.md Appendix B known answers and
analytic properties.  This is what pins the oracle before the GPU parity tests trust it."""
import os

import numpy as np
import pytest

from dpg_slam_b200 import synth
from dpg_slam_b200._abi import (METRIC_POINT_TO_LINE, STOP_DEGENERATE, COV_CENSI_CORR, COV_CENSI_INDEXPAIR, COV_REFERENCE_LIVE, FLAG_CONVERGED,
                                FLAG_COV_SINGULAR, FLAG_EMPTY_INPUT, STOP_ITERATIONS, STOP_MASK,
                                STOP_NO_CORRESPONDENCES, Params)
from oracle import oracle_py as O
from dpg_slam_b200._abi import SEARCH_PROJECTIVE, SEARCH_BRUTE


def T_from_colmajor(Tc):
    return np.array([Tc[0], Tc[1], Tc[12], Tc[13]], np.float32)      # T(0,0), T(1,0), T(0,3), T(1,3)


# ---- covariance: pinned against cov_func_point_to_point.h compiled from /root/reference ----------------
def test_cov_matches_compiled_reference(cov_golden):
    for c in cov_golden:
        n_h = c["P"].shape[0]
        st, cov, H = O.cov_censi(c["P"], c["Q"][:n_h], n_h, c["n_d_used"], T_from_colmajor(c["T_colmajor"]),
                                 sensor_var=c["sensor_var"], live=c["live_in"])
        # Hessian restricted to (x, y, yaw): reference d2J_dX2 rows/cols (0,1,3), cov.h:90-160
        assert np.allclose(H, c["H3"], rtol=1e-12, atol=1e-9), c["name"]
        if c["singular"]:
            continue                     # the reference's 6x6 is singular here (z/pitch/roll block), 3x3 may not be
        assert st == 0, c["name"]
        rel = np.abs(cov - c["cov3"]).max() / np.abs(c["cov3"]).max()
        assert rel < 1e-9, (c["name"], rel)


def test_cov_live_output_is_the_constant_diagonal(cov_golden):
    # cov.h:572-575: the reference's live output is diag(sx2, sy2, st2) whatever the inputs
    for c in cov_golden:
        want = np.diag(np.array(c["live_in"], np.float32).astype(np.float64))
        assert np.array_equal(c["live_cov"], want), c["name"]
        p = Params.defaults(cov_mode=COV_REFERENCE_LIVE, laser_x_variance=c["live_in"][0],
                            laser_y_variance=c["live_in"][1], laser_theta_variance=c["live_in"][2])
        res = O.run_pair(c["P"], c["Q"], [0, 0, 0], p)
        assert np.array_equal(np.array(res.cov).reshape(3, 3), want)


def test_cov_kat1_survey_appendix_b():
    P = np.array([(1.0, 0.5), (2.0, -1.0), (3.5, 0.25), (0.5, 2.0), (-1.0, 1.5)], np.float32)
    Q = np.array([(1.2550874948501587, 0.3773355185985565), (2.3748416900634766, -0.9903373122215271),
                  (3.7775561809539795, 0.4081680178642273), (0.5928352475166321, 1.8299250602722168),
                  (-0.8447542786598206, 1.2076728343963623)], np.float32)
    th = np.float32(0.1)
    T = np.array([np.cos(np.float64(th)), np.sin(np.float64(th)), 0.3, -0.2], np.float32)
    st, cov, H = O.cov_censi(P, Q, 5, 5, T)
    assert st == 0
    assert np.allclose(H, [[10, 0, -7.665528090894], [0, 10, 11.291132763709],
                           [-7.665528090894, 11.291132763709, 52.197628147121]], atol=1e-9)
    assert np.allclose(cov, [[0.004699660395, -0.001029669124, 0.000912736075],
                             [-0.001029669124, 0.005515331835, -0.001343246184],
                             [0.000912736075, -0.001343246184, 0.001190702161]], atol=1e-12)


def test_cov_cap_quirk_h_over_all_d_over_first_200(cov_golden):
    c = next(c for c in cov_golden if c["name"] == "kat2_cap200")
    T = T_from_colmajor(c["T_colmajor"])
    _, capped, _ = O.cov_censi(c["P"], c["Q"], 300, 200, T)
    _, uncapped, _ = O.cov_censi(c["P"], c["Q"], 300, 300, T)
    assert np.abs(capped - c["cov3"]).max() / np.abs(c["cov3"]).max() < 1e-9
    assert np.allclose(uncapped[0, 0], 6.687777482667e-05, rtol=1e-9)       # SURVEY.md KAT-2 without the cap
    assert not np.allclose(capped, uncapped, rtol=1e-3)


def test_cov_singular_falls_back_to_live_diag():
    P = np.zeros((4, 2), np.float32)
    st, cov, _ = O.cov_censi(P, P, 4, 4, [1, 0, 0, 0])
    assert st == FLAG_COV_SINGULAR
    assert np.array_equal(cov, np.diag([0.5, 0.5, np.float64(np.float32(0.3))]))


# ---- guess construction: pinned against math_utils.cc compiled from /root/reference ---------------------
def test_angle_mod_bit_exact(math_golden):
    got = np.array([O.angle_mod(float(a)) for a in math_golden["angle_in"]], np.float32)
    assert np.array_equal(got.view(np.uint32), math_golden["angle_out"].view(np.uint32))


def test_relative_guess_bit_exact(math_golden):
    # runIcp builds the guess with inverseTransformPoint(node_2 pose, node_1 pose), dpg_slam.cc:364-368
    for row, want in zip(math_golden["pose_pairs"], math_golden["inv_out"]):
        g = O.relative_guess(row[0:2], float(row[2]), row[3:5], float(row[5]))
        assert np.array_equal(g.view(np.uint32), want.view(np.uint32))


def test_guess_matrix_entries():
    T = O.guess_matrix([0.3, -0.2, 0.1])
    assert T[2] == np.float32(0.3) and T[3] == np.float32(-0.2)
    assert T[0] == np.float32(np.cos(np.float64(np.float32(0.1)))) and T[1] == np.float32(np.sin(np.float64(np.float32(0.1))))


# ---- scan -> cloud and down-sampling (dpg_node.cc:8-26, dpg_slam.cc:346-360) ----------------------------
def test_ranges_to_cloud_drops_max_range_and_applies_laser_offset():
    sc = synth.Scanner(n_beams=9)
    r = np.array([1, 30, 2, 31, 3, 29.999, 4, np.inf, 5], np.float32)
    xy = O.ranges_to_cloud(r, sc)
    assert xy.shape == (6, 2)                                   # 30, 31, inf dropped (range >= range_max)
    inc = np.float32((np.float64(np.float32(sc.angle_max) - np.float32(sc.angle_min))) / 8.0)
    kept = [0, 2, 4, 5, 6, 8]
    for k, i in enumerate(kept):
        a = np.float32(inc * np.float32(i) + np.float32(sc.angle_min))
        want = (np.float32(0.2) + np.float32(np.float64(r[i]) * np.cos(np.float64(a))),
                np.float32(0.0) + np.float32(np.float64(r[i]) * np.sin(np.float64(a))))
        assert xy[k, 0] == want[0] and xy[k, 1] == want[1]


def test_downsample_keeps_every_dth_compacted_index():
    xy = np.arange(46, dtype=np.float32).reshape(23, 2)
    assert np.array_equal(O.downsample(xy, 5), xy[::5])
    assert np.array_equal(O.downsample(xy, 1), xy)
    assert O.downsample(xy[:0], 5).shape == (0, 2)


# ---- ICP loop (PCL semantics restated; parity UNPINNED, properties only) ----------
def _room_pair(seed=1):
    wl = synth.config_room_pair(seed=seed)
    pts, off = O.clouds_from_ranges(wl.ranges, wl.scanner)
    return wl, pts[off[0]:off[1]], pts[off[1]:off[2]]


def test_icp_recovers_known_offset_noise_free():
    sc = synth.Scanner(noise_sigma=0.0)
    poses = np.array([[0.0, 0.0, 0.0], [0.30, -0.20, 0.10]])
    ranges = synth.cast_scans(synth.world_room(), poses, sc, 1)
    tgt, src = O.ranges_to_cloud(ranges[0], sc), O.ranges_to_cloud(ranges[1], sc)
    p = Params.defaults(downsample_divisor=1)
    res = O.run_pair(src, tgt, [0.35, -0.15, 0.12], p)
    assert res.status & FLAG_CONVERGED
    # sampling differs between the two scans, so ICP recovers the offset to a few mm, not 1e-6
    assert abs(res.tx - 0.30) < 5e-3 and abs(res.ty + 0.20) < 5e-3 and abs(res.theta - 0.10) < 2e-3


def test_icp_identical_clouds_converges_to_identity_in_one_step():
    _, tgt, _ = _room_pair()
    p = Params.defaults(downsample_divisor=1)
    res = O.run_pair(tgt, tgt, [0, 0, 0], p)
    assert res.iterations == 1 and res.mse == 0.0 and res.n_correspondences == len(tgt)
    assert (res.tx, res.ty, res.theta) == (0.0, 0.0, 0.0) and res.status & FLAG_CONVERGED


def test_grid_search_equals_brute_force():
    wl = synth.config_corridor(n_pairs=6, seed=11)
    pts, off = O.clouds_from_ranges(wl.ranges, wl.scanner)
    for div in (1, 5):
        p = Params.defaults(downsample_divisor=div, cov_mode=COV_CENSI_CORR)
        a, _ = O.run_batch(pts, off, wl.src_idx, wl.tgt_idx, wl.guess, p, fast=0)
        b, _ = O.run_batch(pts, off, wl.src_idx, wl.tgt_idx, wl.guess, p, fast=1)
        assert a.tobytes() == b.tobytes()


def test_correspondences_reciprocal_gate_and_ties():
    # target points on a lattice: source point exactly between two targets -> lowest index wins
    tgt = np.array([(0, 0), (1, 0), (2, 0), (10, 10)], np.float32)
    src = np.array([(0.5, 0), (2.1, 0), (5, 5), (0.4, 0.0)], np.float32)
    p = Params.defaults()
    k, corr, d2 = O.correspondences(src, tgt, p.copy(use_reciprocal=0))
    assert list(corr) == [0, 2, -1, 0] and k == 3                 # tie at 0.5 -> index 0; (5,5) beyond 0.6 m
    k, corr, _ = O.correspondences(src, tgt, p)
    assert list(corr) == [-1, 2, -1, 0] and k == 2                # target 0 prefers source 3 (0.4 < 0.5)
    assert d2[0] == np.float32(0.25)
    # gate is <=: a distance of exactly floor32(0.36) passes
    thr = np.float32(0.36) if np.float64(np.float32(0.36)) <= 0.6 * 0.6 else np.nextafter(np.float32(0.36), np.float32(0))
    src2 = np.array([(np.sqrt(np.float64(thr)), 0)], np.float32)
    if np.float32(src2[0, 0] * src2[0, 0]) <= thr:
        k, corr, _ = O.correspondences(src2, tgt[:1], p.copy(use_reciprocal=0))
        assert k == 1


def test_icp_stop_reasons_and_flags():
    _, tgt, src = _room_pair()
    p = Params.defaults(downsample_divisor=1)
    far = O.run_pair(src, tgt, [50, 50, 0], p)                    # nothing within 0.6 m
    assert far.status & STOP_MASK == STOP_NO_CORRESPONDENCES and not far.status & FLAG_CONVERGED
    assert far.iterations == 0 and far.tx == 50 and far.ty == 50  # the guess is returned, record always written
    one = O.run_pair(src, tgt, [0.35, -0.15, 0.12], p.copy(max_iterations=1))
    assert one.status & STOP_MASK == STOP_ITERATIONS and one.status & FLAG_CONVERGED and one.iterations == 1
    empty = O.run_pair(src[:0], tgt, [0, 0, 0], p)
    assert empty.status & FLAG_EMPTY_INPUT and empty.status & STOP_MASK == STOP_NO_CORRESPONDENCES
    two = O.run_pair(src[:2], tgt, [0.3, -0.2, 0.1], p)
    assert two.status & STOP_MASK == STOP_NO_CORRESPONDENCES      # < 3 correspondences (PCL min_number_correspondences_)


def test_censi_corr_covariance_is_spd_and_small():
    wl = synth.config_loop_closure(n_pairs=12, n_scans=40, seed=5)
    pts, off = O.clouds_from_ranges(wl.ranges, wl.scanner)
    p = Params.defaults(downsample_divisor=1, cov_mode=COV_CENSI_CORR)
    res, _ = O.run_batch(pts, off, wl.src_idx, wl.tgt_idx, wl.guess, p, fast=1)
    ok = (res["status"] & FLAG_COV_SINGULAR) == 0
    assert ok.sum() >= 8
    for c in res["cov"][ok]:
        c = c.reshape(3, 3)
        assert np.allclose(c, c.T, rtol=1e-9, atol=1e-18)
        assert np.all(np.linalg.eigvalsh(0.5 * (c + c.T)) > 0)


def test_indexpair_mode_uses_full_clouds_and_shorter_length():
    wl, tgt, src = _room_pair()
    p = Params.defaults(cov_mode=COV_CENSI_INDEXPAIR)             # divisor 5 for ICP, FULL clouds for the covariance
    res = O.run_pair(src, tgt[:-7], wl.guess[0], p)
    nh = min(len(src), len(tgt) - 7)
    T = [res.rot_c, res.rot_s, res.tx, res.ty]
    _, want, _ = O.cov_censi(src, tgt[:-7], nh, 200, T)
    assert np.array_equal(np.array(res.cov).reshape(3, 3), want)


def test_enumerate_pairs_matches_reference_loop_order():
    rng = np.random.default_rng(3)
    xy = rng.uniform(0, 12, (60, 2)).astype(np.float32)
    ps = (np.arange(60) // 25).astype(np.int32)
    src, tgt = O.enumerate_pairs(xy, ps, 5.0, 2.0)
    want = []
    for i in range(1, 60):                                        # dpg_slam.cc:79-107
        want.append((i, i - 1))
        for j in range(0, i - 1):
            d = np.sqrt(np.float32(np.float32((xy[j, 0] - xy[i, 0]) ** 2) + np.float32((xy[j, 1] - xy[i, 1]) ** 2)))
            if d <= (5.0 if ps[j] == ps[i] else 2.0):
                want.append((i, j))
    assert list(zip(src.tolist(), tgt.tolist())) == want


# ---- point-to-line metric (north-star extension; the oracle is its definition) ---------------------------
def test_point_to_line_converges_faster_and_closer():
    wl = synth.config_loop_closure(n_pairs=40, n_scans=40, seed=3)
    pts, off = O.clouds_from_ranges(wl.ranges, wl.scanner)
    p2p = Params.defaults(downsample_divisor=1)
    p2l = p2p.copy(metric=METRIC_POINT_TO_LINE)
    a, _ = O.run_batch(pts, off, wl.src_idx, wl.tgt_idx, wl.guess, p2p, fast=1)
    b, _ = O.run_batch(pts, off, wl.src_idx, wl.tgt_idx, wl.guess, p2l, fast=1)
    assert np.all(b["status"] & FLAG_CONVERGED)
    assert b["iterations"].mean() < 0.6 * a["iterations"].mean()
    err = lambda r: np.median(np.hypot(r["tx"] - wl.truth[:, 0], r["ty"] - wl.truth[:, 1]))
    assert err(b) < err(a)
    c, _ = O.run_batch(pts, off, wl.src_idx, wl.tgt_idx, wl.guess, p2l, fast=0)
    assert b.tobytes() == c.tobytes()                         # grid search == brute force for this metric too


def test_point_to_line_degenerate_geometry_is_flagged():
    # a single straight wall: sliding along it is unobservable -> normal equations singular
    x = np.linspace(-3, 3, 200, dtype=np.float32)
    wall = np.stack([x, np.full_like(x, 2.0)], 1)
    p = Params.defaults(downsample_divisor=1, metric=METRIC_POINT_TO_LINE)
    res = O.run_pair(wall, wall, [0.0, 0.05, 0.0], p)
    assert res.status & STOP_MASK == STOP_DEGENERATE and not res.status & FLAG_CONVERGED
    # the same wall plus a perpendicular one is well posed
    y = np.linspace(-1, 2, 100, dtype=np.float32)
    corner = np.concatenate([wall, np.stack([np.full_like(y, 3.0), y], 1)])
    res = O.run_pair(corner, corner, [0.03, 0.05, 0.01], p)
    assert res.status & FLAG_CONVERGED and abs(res.tx) < 1e-3 and abs(res.ty) < 1e-3 and abs(res.theta) < 1e-3


def test_point_to_line_isolated_points_fall_back_to_point_rows():
    # targets farther apart than the gate: no usable segment, every row is a point row -> well posed
    rng = np.random.default_rng(1)
    tgt = (rng.uniform(-20, 20, (60, 2))).astype(np.float32)
    th = 0.01
    R = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
    src = ((tgt - np.array([0.05, -0.03])) @ R).astype(np.float32)
    p = Params.defaults(downsample_divisor=1, metric=METRIC_POINT_TO_LINE)
    res = O.run_pair(src, tgt, [0, 0, 0], p)
    assert res.status & FLAG_CONVERGED and res.n_correspondences == 60
    assert abs(res.theta - th) < 1e-4


def test_factor_sqrt_information():
    from dpg_slam_b200._abi import FLAG_FACTOR_INVALID, RESULT_DTYPE
    r = np.zeros(3, RESULT_DTYPE)
    C = np.array([[4e-5, -5e-6, 3e-6], [-5e-6, 5e-5, -1e-6], [3e-6, -1e-6, 2e-6]])
    r["cov"][0] = C.reshape(9)
    r["cov"][1] = np.diag([0.5, 0.5, 0.3]).reshape(9)            # the reference's live covariance
    r["cov"][2] = np.array([[1, 2, 0], [2, 1, 0], [0, 0, 1]], float).reshape(9)   # indefinite
    r["tx"], r["theta"] = [1, 2, 3], [0.1, 0.2, 0.3]
    f = O.factors(r, [1, 2, 3], [0, 1, 2])
    R = f["sqrt_info"][0].reshape(3, 3)
    assert np.allclose(R.T @ R, np.linalg.inv(C), rtol=1e-12) and np.all(np.tril(R, -1) == 0)
    assert np.allclose(f["sqrt_info"][1].reshape(3, 3), np.diag([2 ** 0.5, 2 ** 0.5, (1 / 0.3) ** 0.5]))
    assert f["status"][2] & FLAG_FACTOR_INVALID and not f["sqrt_info"][2].any()
    assert list(f["from_node"]) == [0, 1, 2] and list(f["to_node"]) == [1, 2, 3] and f["tx"][1] == 2


# ---- rigid step vs Eigen's published Umeyama algorithm in float32 3-D (what PCL's TransformationEstimationSVD runs) ----
def _umeyama_float32_3d(P, Q):
    """Eigen::umeyama(src, dst, with_scaling=false) restated (Geometry/Umeyama.h): float32, 3-D points with z = 0."""
    f = np.float32
    src = np.concatenate([P, np.zeros((len(P), 1))], 1).astype(f).T       # 3 x n
    dst = np.concatenate([Q, np.zeros((len(Q), 1))], 1).astype(f).T
    n = f(src.shape[1])
    mu_s = (src.sum(1) / n).astype(f)
    mu_d = (dst.sum(1) / n).astype(f)
    sd, dd = (src - mu_s[:, None]).astype(f), (dst - mu_d[:, None]).astype(f)
    sigma = ((dd @ sd.T) / n).astype(f)
    U, d, Vt = np.linalg.svd(sigma.astype(f))
    S = np.ones(3, f)
    if np.linalg.det(U.astype(np.float64)) * np.linalg.det(Vt.astype(np.float64)) < 0:
        S[2] = -1
    R = (U * S) @ Vt
    t = mu_d - R @ mu_s
    return R.astype(f), t.astype(f)


def test_planar_closed_form_equals_float32_umeyama_svd():
    """The oracle's binary64 planar Procrustes step is the z = 0 case of PCL's float SVD step: they agree to the
    float32 noise of the SVD path (SURVEY.md A.6 measured <= 1.4e-6 m / 4e-8 rad), far inside the 1e-5 bar."""
    wl = synth.config_loop_closure(n_pairs=10, n_scans=30, seed=9)
    pts, off = O.clouds_from_ranges(wl.ranges, wl.scanner)
    p = Params.defaults(downsample_divisor=1, max_iterations=1)
    worst_t, worst_r = 0.0, 0.0
    for k in range(wl.n_pairs):
        s, t = wl.src_idx[k], wl.tgt_idx[k]
        src, tgt = pts[off[s]:off[s + 1]], pts[off[t]:off[t + 1]]
        res = O.icp(src, tgt, wl.guess[k], p)
        if res.n_correspondences < 3:
            continue
        G = O.guess_matrix(wl.guess[k])
        cur = O.transform_points(G, src)
        _, corr, _ = O.correspondences(cur, tgt, p)
        m = corr >= 0
        R, tt = _umeyama_float32_3d(cur[m], tgt[corr[m]])
        assert abs(R[2, 2] - 1.0) < 1e-6 and abs(tt[2]) < 1e-6          # stays planar (dpg_slam.cc:418-426 checks this)
        # final = step * guess, as the Matrix4f product
        c, sn = np.float64(G[0]), np.float64(G[1])
        Rg = np.array([[c, -sn], [sn, c]])
        Rf = R[:2, :2].astype(np.float64) @ Rg
        tf = R[:2, :2].astype(np.float64) @ np.array([G[2], G[3]], np.float64) + tt[:2]
        th = np.arctan2(Rf[1, 0], Rf[0, 0])
        worst_t = max(worst_t, abs(tf[0] - res.tx), abs(tf[1] - res.ty))
        worst_r = max(worst_r, abs(th - res.theta))
    assert worst_t < 5e-6 and worst_r < 2e-6, (worst_t, worst_r)


# ---- projective search (north-star extension; the oracle is its definition, so it is pinned by properties) ---------
def test_beam_key_is_monotone_in_the_bearing():
    rng = np.random.default_rng(5)
    ang = np.sort(rng.uniform(-np.pi, np.pi, 4000))
    r = rng.uniform(0.05, 30.0, ang.size)
    ox, oy = 0.2, 0.0
    keys = np.array([O.lib().orc_beam_key(np.float32(ox + rr * np.cos(a)), np.float32(oy + rr * np.sin(a)), ox, oy)
                     for a, rr in zip(ang, r)])
    assert np.all(keys > -2.0 - 1e-6) and np.all(keys <= 2.0 + 1e-6)
    # monotone up to float rounding of nearly equal bearings
    assert np.all(np.diff(keys) > -1e-5)
    assert O.lib().orc_beam_key(0.2, 0.0, 0.2, 0.0) == 0.0                      # the origin itself
    assert O.lib().orc_beam_key(1.2, 0.0, 0.2, 0.0) == 0.0                      # straight ahead
    assert O.lib().orc_beam_key(0.2, 1.0, 0.2, 0.0) == 1.0                      # +90 degrees
    assert O.lib().orc_beam_key(-1.0, 0.0, 0.2, 0.0) == 2.0                     # behind: the (-pi, pi] cut


def test_projective_with_a_window_covering_the_scan_equals_exact_search():
    """W >= n makes every point a candidate: the projective pass must then return what brute force returns."""
    wl = synth.config_loop_closure(n_pairs=4, n_scans=12, n_beams=361, seed=8)
    pts, off = O.clouds_from_ranges(wl.ranges, wl.scanner)
    for k in range(wl.n_pairs):
        s, t = wl.src_idx[k], wl.tgt_idx[k]
        src, tgt = pts[off[s]:off[s + 1]], pts[off[t]:off[t + 1]]
        T = O.guess_matrix(wl.guess[k])
        cur = O.transform_points(T, src)
        for rec in (0, 1):
            pe = Params.defaults(downsample_divisor=1, search=SEARCH_BRUTE, use_reciprocal=rec)
            pp = Params.defaults(downsample_divisor=1, search=SEARCH_PROJECTIVE, use_reciprocal=rec, projective_window=1024)
            ke, ce, de = O.correspondences(cur, tgt, pe)
            kp, cp, dp = O.correspondences(cur, tgt, pp, src_orig=src, T=T)
            assert ke == kp and np.array_equal(ce, cp)
            assert np.array_equal(de.view(np.uint32), dp.view(np.uint32))
    pe = Params.defaults(downsample_divisor=1, search=SEARCH_BRUTE, cov_mode=COV_CENSI_CORR)
    pp = pe.copy(search=SEARCH_PROJECTIVE, projective_window=1024)
    re, _ = O.run_batch(pts, off, wl.src_idx, wl.tgt_idx, wl.guess, pe, fast=0)
    rp, _ = O.run_batch(pts, off, wl.src_idx, wl.tgt_idx, wl.guess, pp, fast=0)
    assert re.tobytes() == rp.tobytes()


def test_projective_small_window_stays_close_to_exact_search():
    wl = synth.config_corridor(n_pairs=24, n_beams=721, seed=6)
    pts, off = O.clouds_from_ranges(wl.ranges, wl.scanner)
    pe = Params.defaults(downsample_divisor=1)
    pp = pe.copy(search=SEARCH_PROJECTIVE, projective_window=8)
    re, _ = O.run_batch(pts, off, wl.src_idx, wl.tgt_idx, wl.guess, pe, fast=1)
    rp, _ = O.run_batch(pts, off, wl.src_idx, wl.tgt_idx, wl.guess, pp, fast=0)
    assert np.all(rp["status"] & FLAG_CONVERGED)
    d = np.hypot(re["tx"] - rp["tx"], re["ty"] - rp["ty"])
    assert np.median(d) < 2e-3 and np.median(np.abs(re["theta"] - rp["theta"])) < 1e-3
    err_e = np.hypot(re["tx"] - wl.truth[:, 0], re["ty"] - wl.truth[:, 1])
    err_p = np.hypot(rp["tx"] - wl.truth[:, 0], rp["ty"] - wl.truth[:, 1])
    assert np.median(err_p) < 1.5 * np.median(err_e) + 5e-3


def test_projective_definition_on_unsorted_and_degenerate_clouds():
    """The bisection runs over the stored order whether or not the keys are sorted: clouds in random order, with duplicates,
    empty or single-point clouds, and queries behind the sensor stay well defined; a window covering the cloud still equals
    brute force, and a tiny window only ever returns indices inside the window around the bisection result."""
    rng = np.random.default_rng(11)
    for n_s, n_t in ((0, 5), (5, 0), (1, 1), (7, 40), (64, 33), (150, 150)):
        src = rng.uniform(-4, 4, (n_s, 2)).astype(np.float32)
        tgt = rng.uniform(-4, 4, (n_t, 2)).astype(np.float32)
        if n_t > 10:
            tgt[5] = tgt[3]                                            # duplicate point: exact tie
        T = np.array([1, 0, 0, 0], np.float32)
        for rec in (0, 1):
            pe = Params.defaults(search=SEARCH_BRUTE, use_reciprocal=rec, max_correspondence_distance=2.5)
            pw = pe.copy(search=SEARCH_PROJECTIVE, projective_window=1024, sensor_x=0.3, sensor_y=-0.2)
            ke, ce, _ = O.correspondences(src, tgt, pe)
            kw, cw, _ = O.correspondences(src, tgt, pw, src_orig=src, T=T)
            assert ke == kw and np.array_equal(ce, cw), (n_s, n_t, rec)
        if n_s and n_t:
            W = 2
            p2 = Params.defaults(search=SEARCH_PROJECTIVE, use_reciprocal=0, projective_window=W, sensor_x=0.3, sensor_y=-0.2,
                                 max_correspondence_distance=30.0)
            _, c2, _ = O.correspondences(src, tgt, p2, src_orig=src, T=T)
            keys = np.array([O.lib().orc_beam_key(x, y, 0.3, -0.2) for x, y in tgt], np.float32)
            for i, (x, y) in enumerate(src):
                kq = np.float32(O.lib().orc_beam_key(x, y, 0.3, -0.2))
                lo, hi = 0, n_t
                while lo < hi:
                    mid = (lo + hi) >> 1
                    if keys[mid] < kq:
                        lo = mid + 1
                    else:
                        hi = mid
                cand = range(max(0, lo - W), min(n_t, lo + W))
                if len(cand) == 0:
                    assert c2[i] == -1
                else:
                    d = [(np.float32(x - tgt[j, 0]) ** 2 + np.float32(y - tgt[j, 1]) ** 2, j) for j in cand]
                    assert c2[i] == min(d)[1], (n_s, n_t, i)


# ---- tie rule, outlier rejectors, the online caller (ABI v3; defined by the oracle) ------------------------------------
def test_sticky_tie_rule_prefers_the_previous_neighbour():
    """Among exact ties the previous pass's neighbour wins, else the lowest index; a non-minimiser seed is ignored;
    the reciprocal test keeps a pair unless a source point is STRICTLY closer (both askers of an exact tie are kept)."""
    from dpg_slam_b200._abi import Params
    tgt = np.array([(0, 0), (1, 0), (2, 0), (10, 10)], np.float32)
    src = np.array([(0.5, 0), (1.5, 0), (0.25, 0)], np.float32)          # 0: ties targets 0/1, 1: ties 1/2, 2: nearest 0
    p = Params.defaults(use_reciprocal=0)
    k, corr, d2, nn = O.correspondences(src, tgt, p, prev_nn=[-1, -1, -1])
    assert list(corr) == [0, 1, 0] and list(nn) == [0, 1, 0]            # no history: lowest index
    k, corr, d2, nn = O.correspondences(src, tgt, p, prev_nn=[1, 2, 1])
    assert list(corr) == [1, 2, 0] and list(nn) == [1, 2, 0]            # history among the minimisers wins; 1 is not one for point 2
    k, corr, d2, nn = O.correspondences(src, tgt, p, prev_nn=[3, 0, 2])
    assert list(corr) == [0, 1, 0]                                      # seeds that are not minimisers are ignored
    # reciprocal, asker wins ties: source points 0 and 1 are both exactly 0.5 from target 1
    src2 = np.array([(0.5, 0), (1.5, 0)], np.float32)
    k, corr, _, _ = O.correspondences(src2, tgt[:3], Params.defaults(), prev_nn=[1, 1])
    assert list(corr) == [1, 1] and k == 2
    # ... but a strictly closer source point takes the target
    src3 = np.array([(0.5, 0), (1.25, 0)], np.float32)
    k, corr, _, _ = O.correspondences(src3, tgt[:3], Params.defaults(), prev_nn=[1, 1])
    assert list(corr) == [-1, 1] and k == 1
    # grid search == brute force with seeds too
    rng = np.random.default_rng(5)
    lat = np.stack([rng.integers(0, 9, 200) * 0.25, rng.integers(0, 9, 200) * 0.25], 1).astype(np.float32)
    q = np.stack([rng.integers(0, 17, 150) * 0.125, rng.integers(0, 17, 150) * 0.125], 1).astype(np.float32)
    prev = rng.integers(-1, 200, 150).astype(np.int32)
    a = O.correspondences(q, lat, Params.defaults(), fast=0, prev_nn=prev)
    b = O.correspondences(q, lat, Params.defaults(), fast=1, prev_nn=prev)
    assert np.array_equal(a[1], b[1]) and np.array_equal(a[3], b[3]) and a[0] == b[0]


def test_outlier_threshold_definitions():
    from dpg_slam_b200._abi import OUTLIER_MEDIAN, OUTLIER_TRIMMED, Params
    rng = np.random.default_rng(9)
    for K in (1, 2, 3, 4, 10, 257, 1000):
        d = (rng.random(K) ** 2).astype(np.float32)
        d[: K // 3] = d[0]                                               # exact ties
        srt = np.sort(d)
        for ratio in (0.3, 0.5, 0.9, 1.0):
            tau = O.outlier_threshold(d, Params.defaults(outlier_mode=OUTLIER_TRIMMED, outlier_param=ratio))
            keep = min(max(3, int(np.floor(ratio * K))), K)
            assert np.float32(tau) == srt[keep - 1]
            assert (d <= np.float32(tau)).sum() >= keep                  # ties at tau are all kept
        for factor in (0.5, 1.0, 2.5):
            tau = O.outlier_threshold(d, Params.defaults(outlier_mode=OUTLIER_MEDIAN, outlier_param=factor))
            lim = np.float64(srt[K // 2]) * factor
            assert np.float64(np.float32(tau)) <= lim < np.float64(np.nextafter(np.float32(tau), np.float32(np.inf)))
    assert O.outlier_threshold(np.ones(5, np.float32), Params.defaults()) == np.inf      # NONE: nothing is rejected


def test_outlier_rejection_makes_icp_robust_to_a_new_object():
    """A third of the source scan sees an object 0.3 m in front of the walls that the target scan does not contain
    (dynamic environment): stock ICP is pulled by it, the trimmed and the median rejector are not."""
    from dpg_slam_b200._abi import FLAG_CONVERGED, OUTLIER_MEDIAN, OUTLIER_TRIMMED, Params
    wl, tgt, src = _room_pair()
    moved = src.copy()
    k0, k1 = len(src) // 3, 2 * len(src) // 3
    laser = np.array([0.2, 0.0], np.float32)
    v = moved[k0:k1] - laser
    moved[k0:k1] = (v * (1.0 - 0.3 / np.linalg.norm(v, axis=1, keepdims=True)) + laser).astype(np.float32)
    truth = wl.truth[0]
    err = {}
    for name, kw in (("none", {}), ("trimmed", dict(outlier_mode=OUTLIER_TRIMMED, outlier_param=0.6)),
                     ("median", dict(outlier_mode=OUTLIER_MEDIAN, outlier_param=1.5))):
        r = O.run_pair(moved, tgt, wl.guess[0], Params.defaults(downsample_divisor=1, cov_mode=2, **kw))
        assert r.status & FLAG_CONVERGED
        err[name] = float(np.hypot(r.tx - truth[0], r.ty - truth[1]))
    assert err["none"] > 0.2 and err["trimmed"] < 0.25 * err["none"] and err["median"] < 0.25 * err["none"], err


def test_enumerate_online_is_the_reference_loop():
    """updatePoseGraphObsConstraints (dpg_slam.cc:255-300): successive (new, preceding), then node i < size - 2 gated
    against — and attached to — the PRECEDING node."""
    rng = np.random.default_rng(3)
    for n in (0, 1, 2, 3, 4, 5, 60, 500):
        xy = rng.uniform(0, 12, (n, 2)).astype(np.float32)
        ps = (np.arange(n) // 150).astype(np.int32)
        src, tgt = O.enumerate_online(xy, ps, 5.0, 2.0)
        want = []
        if n >= 2:
            new, pre = n - 1, n - 2
            want.append((new, pre))
            size = n - 1                                                 # dpg_nodes_.size() before the push
            if size > 1:
                for i in range(max(0, size - 2)):
                    d = np.float32(np.sqrt(np.float32(np.float32((xy[i, 0] - xy[pre, 0]) ** 2) + np.float32((xy[i, 1] - xy[pre, 1]) ** 2))))
                    if d <= (5.0 if ps[i] == ps[pre] else 2.0):
                        want.append((pre, i))
        assert list(zip(src.tolist(), tgt.tolist())) == want, n


# ---- neighbour search: golden vectors from a real FLANN kd-tree (tools/make_flann_golden.py) ---------------------------
def _flann_cases():
    import json
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "flann_nn.json")
    return json.load(open(path))["cases"]


def _hex_f32(h, cols=None):
    a = np.array([int(x, 16) for x in h], np.uint32).view(np.float32)
    return a.reshape(-1, cols) if cols else a


def test_oracle_correspondences_equal_flann_golden_vectors():
    """tests/golden/flann_nn.json holds what a real FLANN single kd-tree (the library PCL's KdTreeFLANN wraps, exact search)
    returned for the forward and backward neighbour queries of PCL's reciprocal correspondence estimation at 18 iterates of
    config-2 / config-3 pairs.  The oracle's forward neighbours and squared distances (bit for bit) and its reciprocal sets
    against them — no OpenCV needed to run this."""
    n_q = 0
    for c in _flann_cases():
        S, T, Tm = _hex_f32(c["source_hex"], 2), _hex_f32(c["target_hex"], 2), _hex_f32(c["T_hex"])
        jf = np.array(c["flann_forward_index"]); d2f = _hex_f32(c["flann_forward_d2_hex"]); back = np.array(c["flann_backward_index"])
        cur = O.transform_points(Tm, S)
        inside = d2f.astype(np.float64) <= 0.36
        _, fwd, fwd_d2 = O.correspondences(cur, T, Params.defaults(use_reciprocal=0))
        assert np.array_equal(fwd >= 0, inside), (c["workload"], c["pair"], c["iterate"])
        assert np.array_equal(fwd[inside], jf[inside])
        assert np.array_equal(fwd_d2[inside].view(np.uint32), d2f[inside].view(np.uint32))
        _, rec, _ = O.correspondences(cur, T, Params.defaults(use_reciprocal=1))
        assert np.array_equal(rec, np.where(inside & (back == np.arange(len(cur))), jf, -1))
        n_q += int(inside.sum())
    assert n_q > 2500
