/*
 * dpg_oracle.h — CPU restatement ("oracle") of DPG-SLAM's scan-matching path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (dpg_slam_b200/, include/) may call,
 * link or import this; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs do, and only as the checker or the timed CPU baseline.
 *
 * PARITY STATUS
 *   covariance (orc_cov_*):   PINNED — checked against the reference's own expressions compiled
 *                             from /root/reference/src/icp_cov/cov_func_point_to_point.h
 *                             (oracle/build_ref.sh -> oracle/_ref/libdpgref.so, goldens in
 *                             tests/golden/cov_ref_*.json) and SURVEY.md Appendix B KAT-1/KAT-2.
 *   guess / AngleMod / scan->cloud / downsample: PINNED the same way (math_utils.cc compiled
 *                             against stub Eigen types) where the code is in /root/reference.
 *   ICP loop (orc_icp):       PARITY UNPINNED — the arithmetic is owned by PCL
 *                             (pcl::IterativeClosestPoint, un-vendored, un-pinned:
 *                             CMakeLists.txt:44 `find_package(PCL 1.3 REQUIRED)`, era 1.8.1/1.10.0),
 *                             absent from /root/reference and from this image; the reference holds
 *                             no tests or golden vectors for it.  The loop below restates PCL's
 *                             published algorithm as documented in SURVEY.md Appendix A and is
 *                             anchored on the reference's call sites (dpg_slam.cc:387-416,445).
 *                             One half of it IS checked against the real thing: the neighbour search
 *                             (orc_correspondences*) returns the indices, the binary32 squared
 *                             distances and the reciprocal sets of a real FLANN KDTreeSingleIndex —
 *                             the library PCL's KdTreeFLANN wraps; OpenCV's bundled copy is in this
 *                             image — on the benchmark configs (tests/test_pcl_emulation.py).  The rigid
 *                             step (Eigen's Umeyama / JacobiSVD) and the convergence glue stay
 *                             restated-only and are cross-checked against an independent float32 SVD
 *                             emulation of the whole loop (tests/pcl_emulation.py).
 *
 * The types of the public C ABI (include/dpgicp.h) are reused for parameters and results so the
 * parity tests compare like with like.
 */
#ifndef DPG_ORACLE_H
#define DPG_ORACLE_H

#include <stdint.h>
#include "../include/dpgicp.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- a2: guess construction (dpg_slam.cc:364-378, math_utils.cc:20-34, math_utils.h:13-16) --- */
float orc_angle_mod(float a);
/* pose of node_2 in node_1's frame: guess = (dx, dy, dtheta) */
void  orc_relative_guess(const float p1[2], float th1, const float p2[2], float th2, float guess[3]);
/* T = (c, s, tx, ty) of the Matrix4f guess (dpg_slam.cc:374-378) */
void  orc_guess_matrix(const float guess[3], float T[4]);

/* ---- a3/a4: scan -> cloud, down-sampling (dpg_slam.cc:488-513, dpg_measurement.h:41-46,102-104,
 *      dpg_node.cc:8-26, dpg_slam.cc:346-360) ------------------------------------------------- */
/* returns number of points written to xy (capacity n_beams) */
int   orc_ranges_to_cloud(const float *ranges, int n_beams, float angle_min, float angle_max,
                          float range_max, float lx, float ly, float ltheta, float *xy);
int   orc_downsample(const float *xy, int n, int divisor, float *out_xy);

/* ---- a5: ICP building blocks ------------------------------------------------------------------ */
void  orc_transform_points(const float T[4], const float *xy, int n, float *out_xy);
/* one correspondence pass (src already transformed). corr[i] = target index or -1. Returns K.
 * fast = 0: brute force;  fast = 1: uniform-grid exact search (same answers, used for timing).  */
int   orc_correspondences(const float *src_t, int ns, const float *tgt, int nt,
                          const dpgicp_params *p, int fast, int32_t *corr, float *d2);
/* the same with what DPGICP_SEARCH_PROJECTIVE needs besides: the untransformed source (its beam-order keys)
 * and T = (c, s, tx, ty), the transform that produced src_t; both NULL = src_t is the original, T = identity */
int   orc_correspondences_ex(const float *src_t, int ns, const float *tgt, int nt, const dpgicp_params *p, int fast,
                             const float *src_orig, const float *T, int32_t *corr, float *d2);
/* the general form: prev_nn (size ns, may be NULL) is the sticky tie preference in / this pass's gated forward neighbour
 * out (exact searches; see the tie rule in include/dpgicp.h); d2 must be an array of ns floats; the outlier rejector
 * of p->outlier_mode is applied last                                                                             */
int   orc_correspondences_seeded(const float *src_t, int ns, const float *tgt, int nt, const dpgicp_params *p, int fast,
                                 const float *src_orig, const float *T, int32_t *prev_nn, int32_t *corr, float *d2);
/* rejection threshold tau of DPGICP_OUTLIER_* for K accepted squared distances (+inf for NONE) */
float orc_outlier_threshold(const float *d2_accepted, int K, const dpgicp_params *p);
/* bearing key of DPGICP_SEARCH_PROJECTIVE (include/dpgicp.h) */
float orc_beam_key(float px, float py, float ox, float oy);

/* optional per-iteration trace: T_iter[4*k] = accumulated (c,s,tx,ty) BEFORE iteration k's
 * correspondence pass (k = 0 is the guess), n_corr[k], so tests can replay any iterate.         */
typedef struct orc_trace {
  int    capacity;      /* in: number of iterations the arrays can hold */
  int    count;         /* out */
  float *T_iter;        /* 4 * capacity */
  int32_t *n_corr;      /* capacity */
} orc_trace;

/* ICP on already down-sampled clouds.  Fills tx,ty,theta,rot_c,rot_s,iterations,status,
 * n_correspondences,mse of *out (cov untouched).  corr_last (size ns, may be NULL) receives the
 * correspondences of the last executed iteration.                                              */
void  orc_icp(const float *src, int ns, const float *tgt, int nt, const float guess[3],
              const dpgicp_params *p, int fast, dpgicp_result *out, orc_trace *trace);
/* the same; nn_state (size ns, may be NULL) receives the last pass's forward neighbours (seeds of the covariance pass) */
void  orc_icp_ex(const float *src, int ns, const float *tgt, int nt, const float guess[3],
                 const dpgicp_params *p, int fast, dpgicp_result *out, orc_trace *trace, int32_t *nn_state);

/* ---- a6: covariance ---------------------------------------------------------------------------- */
/* Censi/Prakhya planar form on index-paired arrays P[k] <-> Q[k]:
 * H over k < n_h, D-term over k < n_d.  T = (c, s, tx, ty) float entries of the final transform.
 * Returns status flags (0 or DPGICP_FLAG_COV_SINGULAR).  H3 (9, may be NULL) receives the Hessian. */
uint32_t orc_cov_censi(const float *P, const float *Q, int n_h, int n_d, const float T[4],
                       double sensor_var, const float live_diag[3], double cov[9], double H3[9]);

/* ---- a1: the whole runIcp (dpg_slam.cc:362-446): full clouds in, record out -------------------- */
void  orc_run_pair(const float *source_full, int n_source, const float *target_full, int n_target,
                   const float guess[3], const dpgicp_params *p, int fast, dpgicp_result *out);

/* batch over a scan store in CSR form (offsets in points), OpenMP over pairs when threads > 1;
 * returns the number of threads actually used.                                                  */
int   orc_run_batch(const float *points, const int64_t *offsets, int n_scans,
                    const int32_t *src_idx, const int32_t *tgt_idx, const float *guess,
                    int64_t n_pairs, const dpgicp_params *p, int fast, int threads,
                    dpgicp_result *out);

/* addObservationConstraint hand-off (dpg_slam.cc:331-338): record -> factor with the upper-triangular
 * square-root information R (R^T R = cov^-1), as noiseModel::Gaussian::Covariance derives it      */
void  orc_factor(const dpgicp_result *rec, int32_t src, int32_t tgt, dpgicp_factor *out);

/* callers' pair enumeration (dpg_slam.cc:79-107): returns count; writes up to capacity pairs    */
int64_t orc_enumerate_pairs(const float *node_xy, const int32_t *node_pass, int n_nodes,
                            float same_pass_radius, float other_pass_radius,
                            int32_t *src_idx, int32_t *tgt_idx, int64_t capacity);

/* updatePoseGraphObsConstraints (dpg_slam.cc:255-300) for the newest node n-1 with preceding node n-2: the successive
 * pair (src n-1, tgt n-2), then every i < n-3 (the loop bound dpg_nodes_.size() - 2 with size = n-1) within the gate of
 * the PRECEDING node -> (src n-2, tgt i)                                                                           */
int64_t orc_enumerate_online(const float *node_xy, const int32_t *node_pass, int n_nodes,
                             float same_pass_radius, float other_pass_radius,
                             int32_t *src_idx, int32_t *tgt_idx, int64_t capacity);

#ifdef __cplusplus
}
#endif
#endif
