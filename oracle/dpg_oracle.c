/*
 * dpg_oracle.c — CPU restatement of DPG-SLAM's runIcp + calculate_ICP_COV.  TEST INFRASTRUCTURE;
 * see dpg_oracle.h for who may use it and for the parity status (ICP loop: PARITY UNPINNED,
 * PCL is absent; covariance: pinned against the compiled reference header).
 *
 * Build with -O2 -ffp-contract=off (no FMA contraction: the arithmetic contract below is stated
 * in individually rounded IEEE-754 operations).
 *
 * ARITHMETIC CONTRACT (shared with the CUDA path, DESIGN.md §"arithmetic contract")
 *   points            binary32 (x, y), |coord| <= 1000
 *   transform         x' = ((c*x) + ((-s)*y)) + tx ; y' = ((s*x) + (c*y)) + ty      binary32, no FMA
 *                     (Matrix4f * Vector4f evaluated column by column, PCL transformCloud)
 *   distance          d2 = (dx*dx) + (dy*dy), dx = px - qx, dy = py - qy           binary32, no FMA
 *                     (FLANN L2_Simple accumulates diff*diff in coordinate order; z term is +0)
 *   nearest neighbour minimum d2; among exact ties the point matched in the previous pass of the same pair wins
 *                     ("sticky"), else the lowest index.  FLANN's tie order is unspecified;
 *                     this rule lets the GPU skip the index bookkeeping whenever nothing strictly improves.
 *   reciprocity       pair (i, j) is kept iff no source point is STRICTLY closer to target j than i is
 *                     (the asker wins exact ties)
 *   gate              d2 <= fl32_floor(max_correspondence_distance^2 computed in binary64)
 *   moment sums       exact int64 fixed point: round-to-nearest-even of value * 2^S
 *                     S = 32 for sums of coordinates, 28 for sums of coordinate products,
 *                     40 for the sum of squared distances  -> independent of summation order
 *   rigid step        binary64, individually rounded ops, closed planar Procrustes (the z = 0
 *                     case of PCL's Umeyama/SVD step), entries rounded to binary32
 *   composition       final = step * final, binary32, k-ordered dot products, no FMA
 */
#include "dpg_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define SCALE_LIN  4294967296.0          /* 2^32 */
#define SCALE_PROD 268435456.0           /* 2^28 */
#define SCALE_D2   1099511627776.0       /* 2^40 */

/* ------------------------------------------------------------------------------------------------
 * a2. guess
 * ---------------------------------------------------------------------------------------------- */

/* math_utils.h:13-16 with T = float: the expression is evaluated in double (M_PI is double) and
 * assigned back to the float. */
float orc_angle_mod(float a) {
  double d = (double)a;
  d -= (M_PI * 2.0) * rint(d / (M_PI * 2.0));
  return (float)d;
}

/* math_utils.cc:20-34 as called from dpg_slam.cc:364-368: src = node_2 pose, target frame = node_1.
 * Eigen::Rotation2Df(-th1) takes cosf/sinf of the float angle; the 2x2 * 2x1 product is a
 * k-ordered dot product in float. */
void orc_relative_guess(const float p1[2], float th1, const float p2[2], float th2, float guess[3]) {
  float tx = p2[0] - p1[0];
  float ty = p2[1] - p1[1];
  float ang = -th1;
  float c = cosf(ang), s = sinf(ang);
  float m01 = -s;
  guess[0] = (c * tx) + (m01 * ty);
  guess[1] = (s * tx) + (c * ty);
  guess[2] = orc_angle_mod(th2 - th1);
}

/* dpg_slam.cc:374-378: `cos(est_angle_displ)` on a float argument inside a Matrix4f initialiser.
 * Restated as the binary64 libm value rounded to binary32 (the ::cos(double) overload); the
 * std::cos(float) overload differs by at most one binary32 ulp. */
void orc_guess_matrix(const float guess[3], float T[4]) {
  T[0] = (float)cos((double)guess[2]);
  T[1] = (float)sin((double)guess[2]);
  T[2] = guess[0];
  T[3] = guess[1];
}

/* ------------------------------------------------------------------------------------------------
 * a3 / a4. scan -> cloud, down-sampling
 * ---------------------------------------------------------------------------------------------- */

/* createNode dpg_slam.cc:497-506: angle_inc = (max-min)/(n-1.0) [double divide, stored float],
 * angle = angle_inc*i + angle_min in float.  MeasurementPoint dpg_measurement.h:41-46 labels
 * range >= max_range MAX_RANGE; getPointInLaserFrame dpg_measurement.h:102-104 is
 * (range*cos(angle), range*sin(angle)) — restated with the double overloads, rounded to float.
 * getCachedPointCloudFromNode dpg_node.cc:13-22 skips MAX_RANGE and applies transformPoint
 * (math_utils.cc:5-18): Rotation2Df(ltheta) * p + t. */
int orc_ranges_to_cloud(const float *ranges, int n_beams, float angle_min, float angle_max,
                        float range_max, float lx, float ly, float ltheta, float *xy) {
  float angle_inc = (float)(((double)(float)(angle_max - angle_min)) / ((double)n_beams - 1.0));
  float c = cosf(ltheta), s = sinf(ltheta);
  float m01 = -s;
  int n = 0;
  for (int i = 0; i < n_beams; ++i) {
    float r = ranges[i];
    if (r >= range_max) continue;
    float angle = angle_inc * (float)i + angle_min;
    float px = (float)((double)r * cos((double)angle));
    float py = (float)((double)r * sin((double)angle));
    float rx = (c * px) + (m01 * py);
    float ry = (s * px) + (c * py);
    xy[2 * n + 0] = lx + rx;
    xy[2 * n + 1] = ly + ry;
    ++n;
  }
  return n;
}

/* dpg_slam.cc:346-360: keep compacted indices 0, d, 2d, ... */
int orc_downsample(const float *xy, int n, int divisor, float *out_xy) {
  if (divisor < 1) divisor = 1;
  int m = 0;
  for (int i = 0; i < n; ++i) {
    if (i % divisor == 0) {
      out_xy[2 * m] = xy[2 * i];
      out_xy[2 * m + 1] = xy[2 * i + 1];
      ++m;
    }
  }
  return m;
}

/* ------------------------------------------------------------------------------------------------
 * a5. ICP building blocks (PCL semantics)
 * ---------------------------------------------------------------------------------------------- */

void orc_transform_points(const float T[4], const float *xy, int n, float *out_xy) {
  const float c = T[0], s = T[1], tx = T[2], ty = T[3];
  const float ms = -s;
  for (int i = 0; i < n; ++i) {
    float x = xy[2 * i], y = xy[2 * i + 1];
    float nx = ((c * x) + (ms * y)) + tx;
    float ny = ((s * x) + (c * y)) + ty;
    out_xy[2 * i] = nx;
    out_xy[2 * i + 1] = ny;
  }
}

static inline float dist2(float ax, float ay, float bx, float by) {
  float dx = ax - bx, dy = ay - by;
  return (dx * dx) + (dy * dy);
}

/* largest binary32 value <= the binary64 threshold: `d2 > max_dist_sqr` (float vs double in PCL's
 * correspondence_estimation.hpp) is then the same predicate as the float compare d2 > thr.       */
static float gate_threshold(const dpgicp_params *p) {
  double d = p->max_correspondence_distance * p->max_correspondence_distance;
  float f = (float)d;
  if ((double)f > d) f = nextafterf(f, -INFINITY);
  return f;
}

/* brute-force (d2, index) argmin of point (qx,qy) over a cloud */
static inline void nn_brute(float qx, float qy, const float *cloud, int n, int *best_i, float *best_d) {
  float bd = INFINITY;
  int bi = -1;
  for (int j = 0; j < n; ++j) {
    float d = dist2(qx, qy, cloud[2 * j], cloud[2 * j + 1]);
    if (d < bd) { bd = d; bi = j; }
  }
  *best_i = bi;
  *best_d = bd;
}

/* --- exact uniform-grid search used only to time a non-strawman CPU baseline ------------------- */
typedef struct {
  double x0, y0, inv;
  int nx, ny;
  int *start;   /* nx*ny + 1 */
  int *items;   /* n, point indices sorted by cell, ascending index inside a cell */
} grid_t;

static void grid_build(grid_t *g, const float *cloud, int n, double cell) {
  double xmin = DBL_MAX, ymin = DBL_MAX, xmax = -DBL_MAX, ymax = -DBL_MAX;
  for (int i = 0; i < n; ++i) {
    double x = cloud[2 * i], y = cloud[2 * i + 1];
    if (x < xmin) xmin = x;
    if (x > xmax) xmax = x;
    if (y < ymin) ymin = y;
    if (y > ymax) ymax = y;
  }
  if (n == 0) { xmin = ymin = 0; xmax = ymax = 0; }
  g->x0 = xmin; g->y0 = ymin; g->inv = 1.0 / cell;
  g->nx = (int)floor((xmax - xmin) * g->inv) + 1;
  g->ny = (int)floor((ymax - ymin) * g->inv) + 1;
  int cells = g->nx * g->ny;
  g->start = (int *)calloc((size_t)cells + 1, sizeof(int));
  g->items = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
  for (int i = 0; i < n; ++i) {
    int cx = (int)floor((cloud[2 * i] - g->x0) * g->inv);
    int cy = (int)floor((cloud[2 * i + 1] - g->y0) * g->inv);
    g->start[cy * g->nx + cx + 1]++;
  }
  for (int c = 0; c < cells; ++c) g->start[c + 1] += g->start[c];
  int *fill = (int *)malloc(sizeof(int) * (size_t)(cells > 0 ? cells : 1));
  memcpy(fill, g->start, sizeof(int) * (size_t)cells);
  for (int i = 0; i < n; ++i) {
    int cx = (int)floor((cloud[2 * i] - g->x0) * g->inv);
    int cy = (int)floor((cloud[2 * i + 1] - g->y0) * g->inv);
    g->items[fill[cy * g->nx + cx]++] = i;
  }
  free(fill);
}

static void grid_free(grid_t *g) { free(g->start); free(g->items); }

/* (d2, index) argmin restricted to the 3x3 cells around the query.  With cell > gate distance every
 * point whose computed d2 passes the gate lies in those cells, so after gating the answer equals
 * the brute-force one (a best candidate with d2 > thr is reported as "no neighbour" either way). */
static inline void nn_grid(const grid_t *g, float qx, float qy, const float *cloud, int *best_i, float *best_d) {
  float bd = INFINITY;
  int bi = -1;
  int cx = (int)floor(((double)qx - g->x0) * g->inv);
  int cy = (int)floor(((double)qy - g->y0) * g->inv);
  for (int yy = cy - 1; yy <= cy + 1; ++yy) {
    if (yy < 0 || yy >= g->ny) continue;
    for (int xx = cx - 1; xx <= cx + 1; ++xx) {
      if (xx < 0 || xx >= g->nx) continue;
      int c = yy * g->nx + xx;
      for (int k = g->start[c]; k < g->start[c + 1]; ++k) {
        int j = g->items[k];
        float d = dist2(qx, qy, cloud[2 * j], cloud[2 * j + 1]);
        if (d < bd || (d == bd && j < bi)) { bd = d; bi = j; }
      }
    }
  }
  *best_i = bi;
  *best_d = bd;
}

/* --- DPGICP_SEARCH_PROJECTIVE (include/dpgicp.h): approximate neighbour by projection onto the other
 * scan's beam order.  North-star extension with no reference counterpart; this text is its definition. */

/* bearing key of v = point - sensor origin: a monotone function of atan2(vy, vx) on (-pi, pi] computed
 * without libm, in individually rounded binary32 operations (the division is IEEE round-to-nearest) */
float orc_beam_key(float px, float py, float ox, float oy) {
  const float vx = px - ox, vy = py - oy;
  const float a = fabsf(vx) + fabsf(vy);
  const float t = a > 0.0f ? vy / a : 0.0f;
  if (vx >= 0.0f) return t;
  return vy >= 0.0f ? 2.0f - t : -2.0f - t;
}

/* plain bisection over the stored order (well defined whether or not the keys are sorted) */
static inline int key_lower_bound(const float *keys, int n, float k) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (keys[mid] < k) lo = mid + 1; else hi = mid;
  }
  return lo;
}

/* (d2, index) argmin of (qx, qy) over cloud[c - W, c + W) clipped to [0, n) */
static inline void nn_window(float qx, float qy, const float *cloud, int n, int c, int W, int *best_i, float *best_d) {
  float bd = INFINITY;
  int bi = -1;
  const int j0 = c - W < 0 ? 0 : c - W, j1 = c + W > n ? n : c + W;
  for (int j = j0; j < j1; ++j) {
    const float d = dist2(qx, qy, cloud[2 * j], cloud[2 * j + 1]);
    if (d < bd) { bd = d; bi = j; }
  }
  *best_i = bi;
  *best_d = bd;
}

static int correspondences_projective(const float *src_t, int ns, const float *tgt, int nt, const dpgicp_params *p,
                                      const float *src_orig, const float T[4], int32_t *corr, float *d2) {
  const float thr = gate_threshold(p);
  const float ox = p->sensor_x, oy = p->sensor_y;
  const int W = p->projective_window;
  float *tkey = (float *)malloc(sizeof(float) * (size_t)(nt > 0 ? nt : 1));
  float *skey = (float *)malloc(sizeof(float) * (size_t)(ns > 0 ? ns : 1));
  for (int j = 0; j < nt; ++j) tkey[j] = orc_beam_key(tgt[2 * j], tgt[2 * j + 1], ox, oy);
  for (int i = 0; i < ns; ++i) skey[i] = orc_beam_key(src_orig[2 * i], src_orig[2 * i + 1], ox, oy);
  int K = 0;
  for (int i = 0; i < ns; ++i) {
    int j, ir;
    float d, dr;
    corr[i] = -1;
    const float qx = src_t[2 * i], qy = src_t[2 * i + 1];
    nn_window(qx, qy, tgt, nt, key_lower_bound(tkey, nt, orc_beam_key(qx, qy, ox, oy)), W, &j, &d);
    if (d2) d2[i] = d;
    if (j < 0 || d > thr) continue;
    if (p->use_reciprocal) {
      /* the matched target point in the source scan's own frame: R^T (r - t), binary32, no FMA */
      const float rx = tgt[2 * j], ry = tgt[2 * j + 1];
      const float ex = rx - T[2], ey = ry - T[3];
      const float bx = (T[0] * ex) + (T[1] * ey);
      const float by = (T[0] * ey) - (T[1] * ex);
      nn_window(rx, ry, src_t, ns, key_lower_bound(skey, ns, orc_beam_key(bx, by, ox, oy)), W, &ir, &dr);
      if (dr > thr || ir != i) continue;
    }
    corr[i] = j;
    ++K;
  }
  free(tkey); free(skey);
  return K;
}

/* ---- outlier rejection (north-star "outlier-trim logic"; the reference registers no rejector — dpg_slam.cc:408-412 —
 * so this is OFF by default and defined HERE).  Both modes need one order statistic of the accepted squared
 * distances (binary32, compared as numbers; ties at the threshold are all kept, so the result does not depend on
 * any ordering of equal values):
 *   DPGICP_OUTLIER_TRIMMED  (PCL CorrespondenceRejectorTrimmed analogue): keep_n = max(3, floor(param * K));
 *                           tau = keep_n-th smallest d2; keep d2 <= tau.
 *   DPGICP_OUTLIER_MEDIAN   (PCL CorrespondenceRejectorMedianDistance analogue): med = sorted d2 [K / 2];
 *                           tau = largest binary32 <= (double)med * param; keep d2 <= tau.                    */
static int cmp_f32(const void *a, const void *b) {
  const float x = *(const float *)a, y = *(const float *)b;
  return (x > y) - (x < y);
}

float orc_outlier_threshold(const float *d2_accepted, int K, const dpgicp_params *p) {
  if (p->outlier_mode == DPGICP_OUTLIER_NONE || K < 1) return INFINITY;
  float *v = (float *)malloc(sizeof(float) * (size_t)K);
  memcpy(v, d2_accepted, sizeof(float) * (size_t)K);
  qsort(v, (size_t)K, sizeof(float), cmp_f32);
  float tau;
  if (p->outlier_mode == DPGICP_OUTLIER_TRIMMED) {
    int keep = (int)floor(p->outlier_param * (double)K);
    if (keep < 3) keep = 3;
    if (keep > K) keep = K;
    tau = v[keep - 1];
  } else {
    const double lim = (double)v[K / 2] * p->outlier_param;
    tau = (float)lim;
    if ((double)tau > lim) tau = nextafterf(tau, -INFINITY);
  }
  free(v);
  return tau;
}

static int reject_outliers(int ns, const dpgicp_params *p, int32_t *corr, const float *d2, int K) {
  if (p->outlier_mode == DPGICP_OUTLIER_NONE || K < 1) return K;
  float *acc = (float *)malloc(sizeof(float) * (size_t)K);
  int k = 0;
  for (int i = 0; i < ns; ++i) if (corr[i] >= 0) acc[k++] = d2[i];
  const float tau = orc_outlier_threshold(acc, K, p);
  free(acc);
  int kept = 0;
  for (int i = 0; i < ns; ++i) {
    if (corr[i] < 0) continue;
    if (d2[i] <= tau) ++kept; else corr[i] = -1;
  }
  return kept;
}

/* PCL CorrespondenceEstimation::determine[Reciprocal]Correspondences (Appendix A.3-2): for each
 * source index in order: forward NN, gate, reciprocal NN over the *current* source, require i' == i.
 * src_orig / T (the untransformed source and the transform that produced src_t) are only read by
 * DPGICP_SEARCH_PROJECTIVE; NULL means "src_t is the original, T = identity".
 * prev_nn (size ns, may be NULL; exact searches only) carries the forward neighbour of the previous pass of the same
 * pair: in: the tie preference (-1 = none), out: this pass's gated forward neighbour (-1 = none within the gate).
 * d2 is a required scratch/output array of ns floats.                                                          */
int orc_correspondences_seeded(const float *src_t, int ns, const float *tgt, int nt, const dpgicp_params *p, int fast,
                               const float *src_orig, const float *T, int32_t *prev_nn, int32_t *corr, float *d2) {
  const float thr = gate_threshold(p);
  int K = 0;
  if (ns <= 0 || nt <= 0) {
    for (int i = 0; i < ns; ++i) { corr[i] = -1; d2[i] = INFINITY; if (prev_nn) prev_nn[i] = -1; }
    return 0;
  }
  if (p->search == DPGICP_SEARCH_PROJECTIVE) {
    static const float ident[4] = {1.0f, 0.0f, 0.0f, 0.0f};
    K = correspondences_projective(src_t, ns, tgt, nt, p, src_orig ? src_orig : src_t, T ? T : ident, corr, d2);
    return reject_outliers(ns, p, corr, d2, K);
  }
  grid_t gt, gs;
  if (fast) {
    double cell = p->max_correspondence_distance * 1.001;
    grid_build(&gt, tgt, nt, cell);
    if (p->use_reciprocal) grid_build(&gs, src_t, ns, cell);
  }
  for (int i = 0; i < ns; ++i) {
    int j, ir;
    float d, dr;
    corr[i] = -1;
    if (fast) nn_grid(&gt, src_t[2 * i], src_t[2 * i + 1], tgt, &j, &d);
    else nn_brute(src_t[2 * i], src_t[2 * i + 1], tgt, nt, &j, &d);
    d2[i] = d;
    const int seed = prev_nn ? prev_nn[i] : -1;
    if (prev_nn) prev_nn[i] = -1;
    if (j < 0 || d > thr) continue;
    /* sticky tie rule: last pass's neighbour wins when it is among the minimisers */
    if (seed >= 0 && seed < nt && seed != j && dist2(src_t[2 * i], src_t[2 * i + 1], tgt[2 * seed], tgt[2 * seed + 1]) == d) j = seed;
    if (prev_nn) prev_nn[i] = j;
    if (p->use_reciprocal) {
      if (fast) nn_grid(&gs, tgt[2 * j], tgt[2 * j + 1], src_t, &ir, &dr);
      else nn_brute(tgt[2 * j], tgt[2 * j + 1], src_t, ns, &ir, &dr);
      (void)ir;
      if (dr < d) continue;          /* some source point is strictly closer to target j: not reciprocal */
    }
    corr[i] = j;
    ++K;
  }
  if (fast) {
    grid_free(&gt);
    if (p->use_reciprocal) grid_free(&gs);
  }
  return reject_outliers(ns, p, corr, d2, K);
}

int orc_correspondences_ex(const float *src_t, int ns, const float *tgt, int nt, const dpgicp_params *p, int fast,
                           const float *src_orig, const float *T, int32_t *corr, float *d2) {
  float *tmp = NULL;
  if (!d2) d2 = tmp = (float *)malloc(sizeof(float) * (size_t)(ns > 0 ? ns : 1));
  const int K = orc_correspondences_seeded(src_t, ns, tgt, nt, p, fast, src_orig, T, NULL, corr, d2);
  free(tmp);
  return K;
}

int orc_correspondences(const float *src_t, int ns, const float *tgt, int nt,
                        const dpgicp_params *p, int fast, int32_t *corr, float *d2) {
  return orc_correspondences_ex(src_t, ns, tgt, nt, p, fast, NULL, NULL, corr, d2);
}

/* exact fixed-point moment sums of one correspondence set */
typedef struct {
  int64_t spx, spy, sqx, sqy;     /* 2^32 */
  int64_t sxx, sxy, syx, syy;     /* 2^28: sum px*qx, px*qy, py*qx, py*qy */
  int64_t sd2;                    /* 2^40 */
  int32_t k;
} moments_t;

static inline int64_t fx(double v) { return (int64_t)llrint(v); }   /* round-to-nearest-even */

static void accumulate_moments(const float *src_t, const float *tgt, int ns,
                               const int32_t *corr, const float *d2, moments_t *m) {
  memset(m, 0, sizeof(*m));
  for (int i = 0; i < ns; ++i) {
    int j = corr[i];
    if (j < 0) continue;
    double px = src_t[2 * i], py = src_t[2 * i + 1];
    double qx = tgt[2 * j], qy = tgt[2 * j + 1];
    m->spx += fx(px * SCALE_LIN);
    m->spy += fx(py * SCALE_LIN);
    m->sqx += fx(qx * SCALE_LIN);
    m->sqy += fx(qy * SCALE_LIN);
    m->sxx += fx((px * qx) * SCALE_PROD);
    m->sxy += fx((px * qy) * SCALE_PROD);
    m->syx += fx((py * qx) * SCALE_PROD);
    m->syy += fx((py * qy) * SCALE_PROD);
    m->sd2 += fx((double)d2[i] * SCALE_D2);
    m->k++;
  }
}

/* PCL TransformationEstimationSVD on z = 0 data == planar Procrustes (Appendix A.3-5, A.6).
 * step = (c, s, tx, ty) in binary32. */
static void solve_rigid_p2p(const moments_t *m, float step[4]) {
  /* divisions by K and by the norm are multiplications by one reciprocal each (a short dependent
   * chain: this runs on one GPU thread per pass); the difference from exact division is a few
   * binary64 ulps, far below the binary32 rounding of the step entries */
  const double invK = 1.0 / (double)m->k;
  const double spx = (double)m->spx * (1.0 / SCALE_LIN);
  const double spy = (double)m->spy * (1.0 / SCALE_LIN);
  const double sqx = (double)m->sqx * (1.0 / SCALE_LIN);
  const double sqy = (double)m->sqy * (1.0 / SCALE_LIN);
  const double dot = (double)(m->sxx + m->syy) * (1.0 / SCALE_PROD);
  const double crs = (double)(m->sxy - m->syx) * (1.0 / SCALE_PROD);
  const double a = dot - (((spx * sqx) + (spy * sqy)) * invK);
  const double b = crs - (((spx * sqy) - (spy * sqx)) * invK);
  const double h = sqrt((a * a) + (b * b));
  double c = 1.0, s = 0.0;
  if (h > 0.0) { const double rh = 1.0 / h; c = a * rh; s = b * rh; }
  const double mpx = spx * invK, mpy = spy * invK, mqx = sqx * invK, mqy = sqy * invK;
  const double tx = mqx - ((c * mpx) - (s * mpy));
  const double ty = mqy - ((s * mpx) + (c * mpy));
  step[0] = (float)c;
  step[1] = (float)s;
  step[2] = (float)tx;
  step[3] = (float)ty;
}

/* ---- point-to-line metric (north-star extension; no reference counterpart, defined HERE) -------
 * For an accepted correspondence (p = current source point, q = target point j):
 *   neighbour  j2 = whichever of j-1, j+1 exists and is closer to p (binary32 d2, tie -> j-1)
 *   segment    usable if 0 < d2(q, q_j2) <= gate (binary32): the beam neighbours lie on one surface
 *   line row   n = perp(q_j2 - q) / |q_j2 - q|,  r = n . (p - q),  J = [n_x, n_y, n_y p_x - n_x p_y]
 *   fallback   (no usable segment) two point rows: r = p - q, J = [[1, 0, -p_y], [0, 1, p_x]]
 * Normal equations A = sum J^T J (6 sums), b = -sum J^T r (3 sums): every term is computed in binary64
 * with individually rounded operations and added as exact 2^28 fixed point (order independent).
 * One Gauss-Newton step per ICP iteration: Cholesky of the 3x3, delta = (dx, dy, dth); the rotation
 * of the step is the Cayley map of dth/2 (rational: no libm, so CPU and GPU agree bit for bit):
 *   u = dth / 2,  c = (1 - u^2) / (1 + u^2),  s = 2u / (1 + u^2).                                   */
typedef struct {
  int64_t a11, a12, a13, a22, a23, a33, b1, b2, b3;   /* 2^28 */
  int64_t sd2;                                         /* 2^40 */
  int32_t k;
} normal_eq_t;

static void accumulate_normal_eq(const float *src_t, const float *tgt, int ns, int nt, const int32_t *corr,
                                 const float *d2, float gate, normal_eq_t *m) {
  memset(m, 0, sizeof(*m));
  for (int i = 0; i < ns; ++i) {
    const int j = corr[i];
    if (j < 0) continue;
    const float pxf = src_t[2 * i], pyf = src_t[2 * i + 1];
    const float qxf = tgt[2 * j], qyf = tgt[2 * j + 1];
    int j2 = -1;
    float best = INFINITY;
    if (j - 1 >= 0) { best = dist2(pxf, pyf, tgt[2 * (j - 1)], tgt[2 * (j - 1) + 1]); j2 = j - 1; }
    if (j + 1 < nt) {
      const float dn = dist2(pxf, pyf, tgt[2 * (j + 1)], tgt[2 * (j + 1) + 1]);
      if (dn < best) { best = dn; j2 = j + 1; }
    }
    int line = 0;
    double nx = 0.0, ny = 0.0;
    if (j2 >= 0) {
      const float seg = dist2(tgt[2 * j2], tgt[2 * j2 + 1], qxf, qyf);
      if (seg > 0.0f && seg <= gate) {
        const double tx = (double)tgt[2 * j2] - (double)qxf, ty = (double)tgt[2 * j2 + 1] - (double)qyf;
        const double len = sqrt((tx * tx) + (ty * ty));
        nx = -ty / len;
        ny = tx / len;
        line = 1;
      }
    }
    const double px = pxf, py = pyf;
    const double ex = px - (double)qxf, ey = py - (double)qyf;
    if (line) {
      const double r = (nx * ex) + (ny * ey);
      const double j3 = (ny * px) - (nx * py);
      m->a11 += fx((nx * nx) * SCALE_PROD); m->a12 += fx((nx * ny) * SCALE_PROD); m->a13 += fx((nx * j3) * SCALE_PROD);
      m->a22 += fx((ny * ny) * SCALE_PROD); m->a23 += fx((ny * j3) * SCALE_PROD); m->a33 += fx((j3 * j3) * SCALE_PROD);
      m->b1 += fx((nx * r) * SCALE_PROD); m->b2 += fx((ny * r) * SCALE_PROD); m->b3 += fx((j3 * r) * SCALE_PROD);
    } else {
      m->a11 += fx(1.0 * SCALE_PROD); m->a13 += fx((-py) * SCALE_PROD);
      m->a22 += fx(1.0 * SCALE_PROD); m->a23 += fx(px * SCALE_PROD);
      m->a33 += fx(((px * px) + (py * py)) * SCALE_PROD);
      m->b1 += fx(ex * SCALE_PROD); m->b2 += fx(ey * SCALE_PROD);
      m->b3 += fx(((px * ey) - (py * ex)) * SCALE_PROD);
    }
    m->sd2 += fx((double)d2[i] * SCALE_D2);
    m->k++;
  }
}

/* returns 0 when the normal equations are not positive definite */
static int solve_rigid_p2l(const normal_eq_t *m, float step[4]) {
  const double inv = 1.0 / SCALE_PROD;
  const double a11 = (double)m->a11 * inv, a12 = (double)m->a12 * inv, a13 = (double)m->a13 * inv;
  const double a22 = (double)m->a22 * inv, a23 = (double)m->a23 * inv, a33 = (double)m->a33 * inv;
  const double b1 = -((double)m->b1 * inv), b2 = -((double)m->b2 * inv), b3 = -((double)m->b3 * inv);
  if (!(a11 > 0.0)) return 0;
  const double l11 = sqrt(a11);
  const double l21 = a12 / l11, l31 = a13 / l11;
  const double d22 = a22 - (l21 * l21);
  if (!(d22 > 0.0)) return 0;
  const double l22 = sqrt(d22);
  const double l32 = (a23 - (l31 * l21)) / l22;
  const double d33 = (a33 - (l31 * l31)) - (l32 * l32);
  if (!(d33 > 0.0)) return 0;
  const double l33 = sqrt(d33);
  const double y1 = b1 / l11;
  const double y2 = (b2 - (l21 * y1)) / l22;
  const double y3 = ((b3 - (l31 * y1)) - (l32 * y2)) / l33;
  const double x3 = y3 / l33;
  const double x2 = (y2 - (l32 * x3)) / l22;
  const double x1 = ((y1 - (l21 * x2)) - (l31 * x3)) / l11;
  if (!isfinite(x1) || !isfinite(x2) || !isfinite(x3)) return 0;
  const double u = 0.5 * x3;
  const double uu = u * u;
  const double den = 1.0 + uu;
  step[0] = (float)((1.0 - uu) / den);
  step[1] = (float)((u + u) / den);
  step[2] = (float)x1;
  step[3] = (float)x2;
  return 1;
}

/* final = step * final on the (c, s, tx, ty) parametrisation of the Matrix4f product (A.3-6) */
static void compose(const float st[4], float fin[4]) {
  const float c = st[0], s = st[1], ms = -st[1];
  float nc = (c * fin[0]) + (ms * fin[1]);
  float ns = (s * fin[0]) + (c * fin[1]);
  float ntx = ((c * fin[2]) + (ms * fin[3])) + st[2];
  float nty = ((s * fin[2]) + (c * fin[3])) + st[3];
  fin[0] = nc; fin[1] = ns; fin[2] = ntx; fin[3] = nty;
}

void orc_icp_ex(const float *src, int ns, const float *tgt, int nt, const float guess[3],
                const dpgicp_params *p, int fast, dpgicp_result *out, orc_trace *trace, int32_t *nn_state) {
  float fin[4];
  orc_guess_matrix(guess, fin);
  out->iterations = 0;
  out->n_correspondences = 0;
  out->mse = 0.0;
  uint32_t status = DPGICP_STOP_NONE;
  if (trace) trace->count = 0;

  float *cur = (float *)malloc(sizeof(float) * 2 * (size_t)(ns > 0 ? ns : 1));
  int32_t *corr = (int32_t *)malloc(sizeof(int32_t) * (size_t)(ns > 0 ? ns : 1));
  float *d2 = (float *)malloc(sizeof(float) * (size_t)(ns > 0 ? ns : 1));
  /* forward neighbour of the previous pass (the sticky tie preference); the caller's copy ends up holding the last
   * pass's, which seeds the covariance pass at the final pose exactly as the CUDA path's shared-memory array does */
  int32_t *nn_own = NULL;
  if (!nn_state) nn_state = nn_own = (int32_t *)malloc(sizeof(int32_t) * (size_t)(ns > 0 ? ns : 1));
  for (int i = 0; i < ns; ++i) nn_state[i] = -1;
  /* A.2: the guess is applied to the source once, then steps are applied incrementally */
  orc_transform_points(fin, src, ns, cur);

  if (ns <= 0 || nt <= 0) status |= DPGICP_FLAG_EMPTY_INPUT;

  double mse_prev = DBL_MAX;
  const double rot_thr = 1.0 - p->transformation_epsilon;
  for (;;) {
    if (trace && trace->count < trace->capacity) {
      memcpy(trace->T_iter + 4 * trace->count, fin, sizeof(fin));
    }
    int K = orc_correspondences_seeded(cur, ns, tgt, nt, p, fast, src, fin, nn_state, corr, d2);
    if (trace && trace->count < trace->capacity) trace->n_corr[trace->count++] = K;
    out->n_correspondences = K;
    if (K < 3) {                                   /* A.3-4: min_number_correspondences_ = 3 */
      status |= DPGICP_STOP_NO_CORRESPONDENCES;
      break;
    }
    float st[4];
    double sd2;
    if (p->metric == DPGICP_METRIC_POINT_TO_LINE) {
      normal_eq_t ne;
      accumulate_normal_eq(cur, tgt, ns, nt, corr, d2, gate_threshold(p), &ne);
      if (!solve_rigid_p2l(&ne, st)) {
        status |= DPGICP_STOP_DEGENERATE;
        break;
      }
      sd2 = (double)ne.sd2;
    } else {
      moments_t m;
      accumulate_moments(cur, tgt, ns, corr, d2, &m);
      solve_rigid_p2p(&m, st);
      sd2 = (double)m.sd2;
    }
    orc_transform_points(st, cur, ns, cur);        /* A.3-6 */
    compose(st, fin);
    out->iterations++;
    out->mse = (sd2 * (1.0 / SCALE_D2)) * (1.0 / (double)K);

    /* A.5 DefaultConvergenceCriteria, in PCL's order */
    if (out->iterations >= p->max_iterations) {
      status |= DPGICP_STOP_ITERATIONS | DPGICP_FLAG_CONVERGED;
      break;
    }
    float tr = ((st[0] + st[0]) + 1.0f) - 1.0f;    /* coeff(0,0)+coeff(1,1)+coeff(2,2)-1 in float */
    double cos_angle = 0.5 * (double)tr;
    float tsq = (st[2] * st[2]) + (st[3] * st[3]); /* float arithmetic, then compared as double */
    if (cos_angle >= rot_thr && (double)tsq <= p->transformation_epsilon) {
      status |= DPGICP_STOP_TRANSFORM | DPGICP_FLAG_CONVERGED;
      break;
    }
    if (fabs(out->mse - mse_prev) < 1e-12) {
      status |= DPGICP_STOP_ABS_MSE | DPGICP_FLAG_CONVERGED;
      break;
    }
    mse_prev = out->mse;
  }
  out->rot_c = fin[0];
  out->rot_s = fin[1];
  out->tx = fin[2];
  out->ty = fin[3];
  out->theta = atan2f(fin[1], fin[0]);             /* Rotation2Df::fromRotationMatrix().angle() */
  out->status = status;
  free(cur); free(corr); free(d2); free(nn_own);
}

void orc_icp(const float *src, int ns, const float *tgt, int nt, const float guess[3],
             const dpgicp_params *p, int fast, dpgicp_result *out, orc_trace *trace) {
  orc_icp_ex(src, ns, tgt, nt, guess, p, fast, out, trace, NULL);
}

/* ------------------------------------------------------------------------------------------------
 * a6. covariance — planar closed form of cov_func_point_to_point.h
 * ---------------------------------------------------------------------------------------------- */

static int inv3_sym(const double H[9], double Hi[9]) {
  const double a = H[0], b = H[1], c = H[2], d = H[4], e = H[5], f = H[8];
  const double c00 = d * f - e * e;
  const double c01 = c * e - b * f;
  const double c02 = b * e - c * d;
  const double det = a * c00 + b * c01 + c * c02;
  if (!(fabs(det) > 0.0) || !isfinite(det)) return 0;
  const double id = 1.0 / det;
  Hi[0] = c00 * id; Hi[1] = c01 * id; Hi[2] = c02 * id;
  Hi[3] = Hi[1];    Hi[4] = (a * f - c * c) * id; Hi[5] = (b * c - a * e) * id;
  Hi[6] = Hi[2];    Hi[7] = Hi[5]; Hi[8] = (a * d - b * b) * id;
  for (int i = 0; i < 9; ++i) if (!isfinite(Hi[i])) return 0;
  return 1;
}

uint32_t orc_cov_censi(const float *P, const float *Q, int n_h, int n_d, const float T[4],
                       double sensor_var, const float live_diag[3], double cov[9], double H3[9]) {
  /* cov.h:26-35: x, y are the float entries widened; a = (double)atan2f(T(1,0), T(0,0)) */
  const double x = (double)T[2], y = (double)T[3];
  const double a = (double)atan2f(T[1], T[0]);
  const double c = cos(a), s = sin(a);
  double H[9] = {0};
  for (int k = 0; k < n_h; ++k) {                 /* cov.h:45-283 restricted to rows/cols (0,1,3) */
    const double px = P[2 * k], py = P[2 * k + 1], qx = Q[2 * k], qy = Q[2 * k + 1];
    const double A = px * c - py * s, B = px * s + py * c;
    const double dx = x - qx, dy = y - qy;
    H[0] += 2.0;           H[2] += -2.0 * B;
    H[4] += 2.0;           H[5] += 2.0 * A;
    H[8] += -2.0 * (A * dx + B * dy);
  }
  H[1] = H[3] = 0.0; H[6] = H[2]; H[7] = H[5];
  double M[9] = {0};
  for (int k = 0; k < n_d; ++k) {                 /* cov.h:311-530 restricted the same way */
    const double px = P[2 * k], py = P[2 * k + 1], qx = Q[2 * k], qy = Q[2 * k + 1];
    const double A = px * c - py * s, B = px * s + py * c;
    const double dx = x - qx, dy = y - qy;
    const double D[3][4] = {
        {2.0 * c, -2.0 * s, -2.0, 0.0},
        {2.0 * s, 2.0 * c, 0.0, -2.0},
        {2.0 * (-s * dx + c * dy), -2.0 * (c * dx + s * dy), 2.0 * B, -2.0 * A}};
    for (int r = 0; r < 3; ++r)
      for (int q = 0; q < 3; ++q) {
        double acc = 0.0;
        for (int z = 0; z < 4; ++z) acc += D[r][z] * D[q][z];
        M[3 * r + q] += acc;
      }
  }
  if (H3) memcpy(H3, H, sizeof(H));
  double Hi[9];
  if (n_h <= 0 || !inv3_sym(H, Hi)) {
    memset(cov, 0, 9 * sizeof(double));
    cov[0] = live_diag[0]; cov[4] = live_diag[1]; cov[8] = live_diag[2];
    return DPGICP_FLAG_COV_SINGULAR;
  }
  /* cov.h:553-560: Hinv * D * (0.01 I) * D^T * Hinv */
  double Tm[9];
  for (int r = 0; r < 3; ++r)
    for (int q = 0; q < 3; ++q) {
      double acc = 0.0;
      for (int z = 0; z < 3; ++z) acc += Hi[3 * r + z] * M[3 * z + q];
      Tm[3 * r + q] = acc;
    }
  for (int r = 0; r < 3; ++r)
    for (int q = 0; q < 3; ++q) {
      double acc = 0.0;
      for (int z = 0; z < 3; ++z) acc += Tm[3 * r + z] * Hi[3 * z + q];
      cov[3 * r + q] = sensor_var * acc;
    }
  for (int i = 0; i < 9; ++i)
    if (!isfinite(cov[i])) {
      memset(cov, 0, 9 * sizeof(double));
      cov[0] = live_diag[0]; cov[4] = live_diag[1]; cov[8] = live_diag[2];
      return DPGICP_FLAG_COV_SINGULAR;
    }
  return 0;
}

/* ------------------------------------------------------------------------------------------------
 * a1. runIcp restated (dpg_slam.cc:362-446)
 * ---------------------------------------------------------------------------------------------- */
void orc_run_pair(const float *source_full, int n_source, const float *target_full, int n_target,
                  const float guess[3], const dpgicp_params *p, int fast, dpgicp_result *out) {
  memset(out, 0, sizeof(*out));
  const int div = p->downsample_divisor < 1 ? 1 : p->downsample_divisor;
  float *src = (float *)malloc(sizeof(float) * 2 * (size_t)(n_source > 0 ? n_source : 1));
  float *tgt = (float *)malloc(sizeof(float) * 2 * (size_t)(n_target > 0 ? n_target : 1));
  const int ns = orc_downsample(source_full, n_source, div, src);     /* dpg_slam.cc:397-402 */
  const int nt = orc_downsample(target_full, n_target, div, tgt);
  int32_t *nn_state = (int32_t *)malloc(sizeof(int32_t) * (size_t)(ns > 0 ? ns : 1));
  orc_icp_ex(src, ns, tgt, nt, guess, p, fast, out, NULL, nn_state);  /* dpg_slam.cc:404-416 */

  const float live[3] = {p->laser_x_variance, p->laser_y_variance, p->laser_theta_variance};
  const float T[4] = {out->rot_c, out->rot_s, out->tx, out->ty};
  memset(out->cov, 0, sizeof(out->cov));
  if (p->cov_mode == DPGICP_COV_REFERENCE_LIVE) {                      /* cov.h:572-575 */
    out->cov[0] = live[0]; out->cov[4] = live[1]; out->cov[8] = live[2];
  } else if (p->cov_mode == DPGICP_COV_CENSI_INDEXPAIR) {
    /* dpg_slam.cc:429-431 passes the FULL clouds, paired by index; bound by the shorter cloud
     * (the reference reads out of bounds when N1 < N2) */
    int nh = n_source < n_target ? n_source : n_target;
    int nd = nh;
    if (p->cov_cap > 0 && nd > p->cov_cap) nd = p->cov_cap;           /* cov.h:307 */
    out->status |= orc_cov_censi(source_full, target_full, nh, nd, T, p->cov_sensor_variance, live,
                                 out->cov, NULL);
  } else {
    /* CENSI_CORR: correspondences at the final pose, in source order */
    float *cur = (float *)malloc(sizeof(float) * 2 * (size_t)(ns > 0 ? ns : 1));
    int32_t *corr = (int32_t *)malloc(sizeof(int32_t) * (size_t)(ns > 0 ? ns : 1));
    float *P = (float *)malloc(sizeof(float) * 2 * (size_t)(ns > 0 ? ns : 1));
    float *Q = (float *)malloc(sizeof(float) * 2 * (size_t)(ns > 0 ? ns : 1));
    float *cd2 = (float *)malloc(sizeof(float) * (size_t)(ns > 0 ? ns : 1));
    orc_transform_points(T, src, ns, cur);
    orc_correspondences_seeded(cur, ns, tgt, nt, p, fast, src, T, nn_state, corr, cd2);
    free(cd2);
    int k = 0;
    for (int i = 0; i < ns; ++i)
      if (corr[i] >= 0) {
        P[2 * k] = src[2 * i]; P[2 * k + 1] = src[2 * i + 1];
        Q[2 * k] = tgt[2 * corr[i]]; Q[2 * k + 1] = tgt[2 * corr[i] + 1];
        ++k;
      }
    int nd = k;
    if (p->cov_cap > 0 && nd > p->cov_cap) nd = p->cov_cap;
    out->status |= orc_cov_censi(P, Q, k, nd, T, p->cov_sensor_variance, live, out->cov, NULL);
    free(cur); free(corr); free(P); free(Q);
  }
  free(src); free(tgt); free(nn_state);
}

int orc_run_batch(const float *points, const int64_t *offsets, int n_scans,
                  const int32_t *src_idx, const int32_t *tgt_idx, const float *guess,
                  int64_t n_pairs, const dpgicp_params *p, int fast, int threads,
                  dpgicp_result *out) {
  (void)n_scans;
  int used = 1;
#ifdef _OPENMP
  if (threads < 1) threads = omp_get_max_threads();
  used = threads;
#pragma omp parallel for schedule(dynamic, 4) num_threads(threads)
#endif
  for (int64_t k = 0; k < n_pairs; ++k) {
    const int s = src_idx[k], t = tgt_idx[k];
    orc_run_pair(points + 2 * offsets[s], (int)(offsets[s + 1] - offsets[s]),
                 points + 2 * offsets[t], (int)(offsets[t + 1] - offsets[t]),
                 guess + 3 * k, p, fast, out + k);
  }
  return used;
}

/* dpg_slam.cc:331-338: Pose2(tx, ty, theta) + Gaussian::Covariance(cov) -> sqrt information */
void orc_factor(const dpgicp_result *r, int32_t src, int32_t tgt, dpgicp_factor *f) {
  f->from_node = tgt; f->to_node = src;
  f->tx = r->tx; f->ty = r->ty; f->theta = r->theta;
  const double a = r->cov[0], b = r->cov[1], c = r->cov[2], d = r->cov[4], e = r->cov[5], g = r->cov[8];
  const double c00 = (d * g) - (e * e);
  const double c01 = (c * e) - (b * g);
  const double c02 = (b * e) - (c * d);
  const double det = ((a * c00) + (b * c01)) + (c * c02);
  int ok = (det > 0.0) && isfinite(det);
  double R[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  if (ok) {
    const double id = 1.0 / det;
    const double i00 = c00 * id, i01 = c01 * id, i02 = c02 * id;
    const double i11 = ((a * g) - (c * c)) * id;
    const double i12 = ((b * c) - (a * e)) * id;
    const double i22 = ((a * d) - (b * b)) * id;
    ok = i00 > 0.0;
    if (ok) {
      const double r00 = sqrt(i00);
      const double r01 = i01 / r00, r02 = i02 / r00;
      const double p11 = i11 - (r01 * r01);
      ok = p11 > 0.0;
      if (ok) {
        const double r11 = sqrt(p11);
        const double r12 = (i12 - (r01 * r02)) / r11;
        const double p22 = (i22 - (r02 * r02)) - (r12 * r12);
        ok = p22 > 0.0;
        if (ok) {
          R[0] = r00; R[1] = r01; R[2] = r02; R[4] = r11; R[5] = r12; R[8] = sqrt(p22);
          for (int q = 0; q < 9; ++q) ok = ok && isfinite(R[q]);
        }
      }
    }
  }
  for (int q = 0; q < 9; ++q) f->sqrt_info[q] = ok ? R[q] : 0.0;
  f->status = r->status | (ok ? 0u : DPGICP_FLAG_FACTOR_INVALID);
}

/* dpg_slam.cc:79-107 (reoptimize): for node i > 0: the successive pair (i-1 -> i), then every
 * j < i-1 whose float distance passes the same-pass / other-pass gate.  source = node i. */
int64_t orc_enumerate_pairs(const float *node_xy, const int32_t *node_pass, int n_nodes,
                            float same_pass_radius, float other_pass_radius,
                            int32_t *src_idx, int32_t *tgt_idx, int64_t capacity) {
  int64_t n = 0;
  for (int i = 1; i < n_nodes; ++i) {
    if (n < capacity) { src_idx[n] = i; tgt_idx[n] = i - 1; }
    ++n;
    for (int j = 0; j < i - 1; ++j) {
      float dx = node_xy[2 * j] - node_xy[2 * i];
      float dy = node_xy[2 * j + 1] - node_xy[2 * i + 1];
      float dist = sqrtf((dx * dx) + (dy * dy));
      float thr = (node_pass[j] == node_pass[i]) ? same_pass_radius : other_pass_radius;
      if (dist <= thr) {
        if (n < capacity) { src_idx[n] = i; tgt_idx[n] = j; }
        ++n;
      }
    }
  }
  return n;
}

/* dpg_slam.cc:255-300 (updatePoseGraphObsConstraints): dpg_nodes_ holds nodes 0 .. n-2 (preceding = n-2), the new
 * node is n-1.  Loop closures compare against and attach to the PRECEDING node (dpg_slam.cc:278,295,299); the loop
 * runs over i < dpg_nodes_.size() - 2 (dpg_slam.cc:275). */
int64_t orc_enumerate_online(const float *node_xy, const int32_t *node_pass, int n_nodes,
                             float same_pass_radius, float other_pass_radius,
                             int32_t *src_idx, int32_t *tgt_idx, int64_t capacity) {
  int64_t n = 0;
  if (n_nodes < 2) return 0;
  const int nw = n_nodes - 1, pre = n_nodes - 2;
  if (n < capacity) { src_idx[n] = nw; tgt_idx[n] = pre; }
  ++n;
  const int size = n_nodes - 1;                       /* dpg_nodes_.size() before the push */
  if (size > 1) {
    for (int i = 0; i < size - 2; ++i) {
      float dx = node_xy[2 * i] - node_xy[2 * pre];
      float dy = node_xy[2 * i + 1] - node_xy[2 * pre + 1];
      float dist = sqrtf((dx * dx) + (dy * dy));
      float thr = (node_pass[i] == node_pass[pre]) ? same_pass_radius : other_pass_radius;
      if (dist <= thr) {
        if (n < capacity) { src_idx[n] = pre; tgt_idx[n] = i; }
        ++n;
      }
    }
  }
  return n;
}
