"""TEST-ONLY, independent emulation of what PCL's IterativeClosestPoint does with the reference's settings
(reference call sites src/dpg_slam/dpg_slam.cc:404-416,445; parameters.h:146,159,173,201), written from
SURVEY.md Appendix A.1-A.5 with NONE of the oracle's machinery: a kd-tree (scipy cKDTree, exact search) instead of
grids or boxes, float32 3-D points with z = 0, float32 Umeyama through an SVD (numpy / LAPACK) instead of the planar
closed form, float32 4x4 matrix products for `src' = step * src'` and `final = step * final`, float sums for the MSE,
and PCL's DefaultConvergenceCriteria in PCL's order.  PCL itself is absent from this image, so this does not pin the
ICP loop; it removes self-confirmation: the oracle (fixed-point moments, binary64 closed-form step, own tie rule) and
this emulation share no code and no arithmetic shortcuts, and the test reports where they part.
"""
from __future__ import annotations

import numpy as np
from scipy.spatial import cKDTree

F = np.float32


class FlannTree:
    """Exact nearest neighbour through a REAL FLANN single kd-tree — OpenCV's bundled copy of the library PCL links
    (pcl::KdTreeFLANN = flann::KDTreeSingleIndex, leaf size 15, L2 on float32 3-D points, exact search: PCL passes
    checks = -1, eps = 0).  Returns FLANN's own neighbour choice (its order among exact ties is whatever its tree walk
    yields) and FLANN's own binary32 squared distance."""

    def __init__(self, pts3):
        import cv2
        self._pts = np.ascontiguousarray(pts3, F)
        self._index = cv2.flann.Index(self._pts, dict(algorithm=4, leaf_max_size=15))     # FLANN_INDEX_KDTREE_SINGLE

    def query(self, q3):
        ind, d2 = self._index.knnSearch(np.ascontiguousarray(q3, F), 1, params=dict(checks=-1, eps=0.0, sorted=True))
        return d2[:, 0].astype(F), ind[:, 0].astype(np.int64)


class ScipyTree:
    def __init__(self, pts3):
        self._tree = cKDTree(np.asarray(pts3, np.float64))

    def query(self, q3):
        d, j = self._tree.query(np.asarray(q3, np.float64), k=1)
        return None, j


def guess_matrix(guess) -> np.ndarray:
    """dpg_slam.cc:374-378: Matrix4f from (dx, dy, dtheta); cos/sin of the float angle, stored as float."""
    g = np.asarray(guess, F)
    c, s = F(np.cos(np.float64(g[2]))), F(np.sin(np.float64(g[2])))
    T = np.eye(4, dtype=F)
    T[0, 0], T[0, 1], T[0, 3] = c, -s, g[0]
    T[1, 0], T[1, 1], T[1, 3] = s, c, g[1]
    return T


def transform(T: np.ndarray, pts3: np.ndarray) -> np.ndarray:
    """pcl::transformPointCloud with a Matrix4f: float32 throughout."""
    R, t = T[:3, :3].astype(F), T[:3, 3].astype(F)
    return (pts3.astype(F) @ R.T + t).astype(F)


def _jacobi_svd_eigen_style(a: np.ndarray):
    """Two-sided Jacobi SVD of a square float32 matrix, every operation in float32: the algorithm of Eigen 3.3's
    JacobiSVD (what PCL's TransformationEstimationSVD runs through Eigen::umeyama) — scale by the largest entry, sweep over
    the (p, q) pairs while an off-diagonal entry exceeds 2 eps x the largest diagonal entry, each 2x2 block first made
    symmetric by a left rotation, then diagonalised by a Jacobi rotation; singular values made positive by flipping columns
    of U, then sorted.  RESTATED FROM THE PUBLISHED ALGORITHM (Eigen is absent from this image): it cannot be checked
    against Eigen here; it is a third, independent float32 SVD for the report, of the same family."""
    n = a.shape[0]
    tiny = F(np.finfo(np.float32).tiny)
    precision = F(2.0) * F(np.finfo(np.float32).eps)
    scale = F(np.abs(a).max())
    if scale == 0:
        scale = F(1.0)
    W = (a.astype(F) / scale).astype(F)
    U, V = np.eye(n, dtype=F), np.eye(n, dtype=F)

    def rot_left(M, p, q, c, s):          # rows p, q of M <- J [row p; row q], J = [c s; -s c]
        x, y = M[p, :].copy(), M[q, :].copy()
        M[p, :] = (c * x + s * y).astype(F)
        M[q, :] = (-s * x + c * y).astype(F)

    def rot_right(M, p, q, c, s):         # cols p, q of M <- [col p, col q] J, J = [c s; -s c]
        x, y = M[:, p].copy(), M[:, q].copy()
        M[:, p] = (c * x - s * y).astype(F)
        M[:, q] = (s * x + c * y).astype(F)

    def make_jacobi(x, y, z):             # J with J^T [x y; y z] J diagonal
        deno = F(2.0) * abs(y)
        if deno < tiny:
            return F(1.0), F(0.0)
        tau = F((x - z) / deno)
        w = F(np.sqrt(F(tau * tau) + F(1.0)))
        t = F(1.0) / F(tau + w) if tau > 0 else F(1.0) / F(tau - w)
        sign_t = F(1.0) if t > 0 else F(-1.0)
        nrm = F(1.0) / F(np.sqrt(F(t * t) + F(1.0)))
        return nrm, F(-sign_t * F(y / abs(y)) * abs(t) * nrm)

    max_diag = F(np.abs(np.diag(W)).max())
    for _sweep in range(60):
        finished = True
        for p in range(1, n):
            for q in range(p):
                thr = max(tiny, F(precision * max_diag))
                if abs(W[p, q]) > thr or abs(W[q, p]) > thr:
                    finished = False
                    m00, m01, m10, m11 = W[p, p], W[p, q], W[q, p], W[q, q]
                    t, d = F(m00 + m11), F(m10 - m01)
                    if abs(d) < tiny:
                        c1, s1 = F(1.0), F(0.0)
                    else:
                        u = F(t / d)
                        tmp = F(np.sqrt(F(1.0) + F(u * u)))
                        s1, c1 = F(F(1.0) / tmp), F(u / tmp)
                    # the 2x2 block after the symmetrising rotation
                    b00, b01 = F(c1 * m00 + s1 * m10), F(c1 * m01 + s1 * m11)
                    b11 = F(-s1 * m01 + c1 * m11)
                    cr, sr = make_jacobi(b00, b01, b11)
                    # j_left = rot1 * j_right^T
                    cl = F(c1 * cr - s1 * (-sr))
                    sl = F(c1 * (-sr) + s1 * cr)
                    rot_left(W, p, q, cl, sl)
                    rot_right(U, p, q, cl, -sl)               # U <- U * j_left^T
                    rot_right(W, p, q, cr, sr)
                    rot_right(V, p, q, cr, sr)
                    max_diag = max(max_diag, F(max(abs(W[p, p]), abs(W[q, q]))))
        if finished:
            break
    sv = np.zeros(n, F)
    for i in range(n):
        aii = W[i, i]
        sv[i] = abs(aii)
        if aii < 0:
            U[:, i] = -U[:, i]
    sv = (sv * scale).astype(F)
    for i in range(n):                                        # selection sort, as Eigen: largest first, swap columns
        k = int(np.argmax(sv[i:])) + i
        if sv[k] == 0:
            break
        if k != i:
            sv[[i, k]] = sv[[k, i]]
            U[:, [i, k]] = U[:, [k, i]]
            V[:, [i, k]] = V[:, [k, i]]
    return U, sv, V.T.copy()


def _svd3(sigma: np.ndarray, engine: str):
    """float32 SVD of the 3x3 covariance: LAPACK (numpy) or OpenCV's Jacobi SVD — a one-sided Jacobi iteration in float32,
    the family Eigen's JacobiSVD (what PCL runs) belongs to.  Two engines, so that the report can say how much of the
    stop-iteration noise is the SVD implementation's."""
    if engine == "eigen_jacobi":
        return _jacobi_svd_eigen_style(np.asarray(sigma, F))
    if engine == "opencv":
        import cv2
        w, u, vt = cv2.SVDecomp(np.ascontiguousarray(sigma, F))
        return u.astype(F), w.ravel().astype(F), vt.astype(F)
    return np.linalg.svd(sigma.astype(F))


def umeyama_float32(src3: np.ndarray, dst3: np.ndarray, svd: str = "lapack") -> np.ndarray:
    """Eigen::umeyama(src, dst, with_scaling = false) in float32, as TransformationEstimationSVD<.., float> calls it
    (App. A.3-5): means, covariance (1/n) sum (dst - mu_d)(src - mu_s)^T, SVD, R = U diag(1, 1, +-1) V^T."""
    n = F(src3.shape[0])
    mu_s = (src3.sum(axis=0, dtype=F) / n).astype(F)
    mu_d = (dst3.sum(axis=0, dtype=F) / n).astype(F)
    sd, dd = (src3 - mu_s).astype(F), (dst3 - mu_d).astype(F)
    sigma = ((dd.T @ sd) * (F(1.0) / n)).astype(F)
    U, _, Vt = _svd3(sigma, svd)
    S = np.ones(3, F)
    if np.linalg.det(U.astype(np.float64)) * np.linalg.det(Vt.astype(np.float64)) < 0:
        S[2] = F(-1.0)
    R = ((U * S) @ Vt).astype(F)
    T = np.eye(4, dtype=F)
    T[:3, :3] = R
    T[:3, 3] = (mu_d - R @ mu_s).astype(F)
    return T


def icp(src_xy, tgt_xy, guess, max_iterations=500, eps=5e-9, max_dist=0.6, reciprocal=True, trace=None, nn="scipy", svd="lapack"):
    """-> dict(T (4x4 float32), converged, iterations, stop, n_corr, mse).  stop in {"iterations", "transform",
    "abs_mse", "no_correspondences"}.  ``trace`` (a list) receives (final BEFORE the pass as (c, s, tx, ty), K) per pass.
    ``nn``: "scipy" (cKDTree) or "flann" (OpenCV's FLANN single kd-tree: PCL's own neighbour library); ``svd``: "lapack" or
    "opencv" (a float32 Jacobi SVD)."""
    Tree = FlannTree if nn == "flann" else ScipyTree
    src = np.zeros((len(src_xy), 3), F)
    tgt = np.zeros((len(tgt_xy), 3), F)
    if len(src_xy):
        src[:, :2] = np.asarray(src_xy, F)
    if len(tgt_xy):
        tgt[:, :2] = np.asarray(tgt_xy, F)
    final = guess_matrix(guess)                                         # A.2
    cur = transform(final, src)
    out = dict(T=final, converged=False, iterations=0, stop="no_correspondences", n_corr=0, mse=0.0)
    if len(src) == 0 or len(tgt) == 0:
        return out
    tree_t = Tree(tgt)                                                  # A.1: target tree built once
    max_d2 = np.float64(max_dist) * np.float64(max_dist)                # double threshold vs float distance (A.3-2)
    mse_prev = np.finfo(np.float64).max
    it = 0
    while True:
        d2_tree, j = tree_t.query(cur)
        diff = (cur - tgt[j]).astype(F)
        d2 = (diff[:, 0] * diff[:, 0] + diff[:, 1] * diff[:, 1] + diff[:, 2] * diff[:, 2]).astype(F)   # FLANN L2_Simple
        if d2_tree is not None:
            d2 = d2_tree                                                # FLANN's own value (bit-equal, tested)
        ok = ~(d2.astype(np.float64) > max_d2)
        if reciprocal:
            tree_s = Tree(cur)                                          # rebuilt every iteration: the source moved
            _, back = tree_s.query(tgt[j])
            ok &= back == np.arange(len(cur))
        idx = np.nonzero(ok)[0]
        K = len(idx)
        out["n_corr"] = K
        if trace is not None:
            trace.append((np.array([final[0, 0], final[1, 0], final[0, 3], final[1, 3]], F), K))
        if K < 3:                                                       # A.3-4
            out.update(converged=False, stop="no_correspondences", iterations=it, T=final)
            return out
        step = umeyama_float32(cur[idx], tgt[j[idx]], svd)              # A.3-5
        cur = transform(step, cur)                                      # A.3-6
        final = (step @ final).astype(F)
        it += 1
        mse = float(np.sum(d2[idx].astype(np.float64)) / np.float64(K))
        out.update(T=final, iterations=it, mse=mse)
        # A.5 DefaultConvergenceCriteria, in PCL's order
        if it >= max_iterations:
            out.update(converged=True, stop="iterations")
            return out
        cos_angle = 0.5 * (np.float64(step[0, 0]) + np.float64(step[1, 1]) + np.float64(step[2, 2]) - 1.0)
        t2 = np.float64(step[0, 3]) ** 2 + np.float64(step[1, 3]) ** 2 + np.float64(step[2, 3]) ** 2
        if cos_angle >= 1.0 - eps and t2 <= eps:
            out.update(converged=True, stop="transform")
            return out
        if abs(mse - mse_prev) < 1e-12:
            out.update(converged=True, stop="abs_mse")
            return out
        mse_prev = mse


def pose_of(T: np.ndarray):
    """dpg_slam.cc:434-439: (tx, ty, theta)"""
    return float(T[0, 3]), float(T[1, 3]), float(np.arctan2(np.float64(T[1, 0]), np.float64(T[0, 0])))
