/*
 * ref_harness.cc — C entry points around the REFERENCE'S OWN sources, compiled where they lie:
 *     /root/reference/src/icp_cov/cov_func_point_to_point.h   (calculate_ICP_COV)
 *     /root/reference/src/dpg_slam/math_utils.{h,cc}          (transformPoint, inverseTransformPoint, AngleMod)
 * against the container stand-ins in oracle/ref_stubs/ (Eigen and PCL are absent from this image).
 * TEST INFRASTRUCTURE: used by tests/ and tools/make_golden.py to pin oracle/dpg_oracle.c.
 * Built by oracle/build_ref.sh into oracle/_ref/libdpgref.so (git-ignored).  No reference source
 * is copied into this repository: REF_ROOT is passed as an include path.
 *
 * What the harness adds on top of the reference's text:
 *   - ref_cov_live():      calls calculate_ICP_COV and returns its live output (cov.h:572-575) and,
 *                          through the MatrixXd destruction tap, the local matrices the function
 *                          computed and discarded: d2J_dX2 (6x6, cov.h:41-283), d2J_dZdX (6x6n,
 *                          cov.h:311-528) and cov_z's size (cov.h:553-554).
 *   - ref_cov_intended():  evaluates the reference's COMMENTED-OUT product (cov.h:560)
 *                          d2J_dX2^-1 * d2J_dZdX * cov_z * d2J_dZdX^T * d2J_dX2^-1 on those tapped
 *                          matrices with plain Gauss-Jordan, and the (0,1,3) selection of cov.h:564-566.
 */
#include <cmath>
#include <cstring>
#include <memory>
#include <utility>
#include <vector>

#include <pcl/impl/point_types.hpp>
#include <pcl/point_cloud.h>

namespace Eigen { namespace tap {
thread_local Slot ring[3];
thread_local bool enabled = false;
} }

#include "icp_cov/cov_func_point_to_point.h"
#include "dpg_slam/math_utils.h"

namespace {
typedef pcl::PointCloud<pcl::PointXYZ> Cloud;

Cloud::Ptr make_cloud(const float *xy, int n) {
  Cloud::Ptr c(new Cloud);
  c->points.resize((size_t)n);
  for (int i = 0; i < n; ++i) { c->points[i].x = xy[2 * i]; c->points[i].y = xy[2 * i + 1]; c->points[i].z = 0.f; c->points[i].pad = 1.f; }
  return c;
}

bool invert6(const double *A, double *Ai) {   /* Gauss-Jordan with partial pivoting, column-major in/out */
  double a[6][12];
  for (int r = 0; r < 6; ++r)
    for (int c = 0; c < 6; ++c) { a[r][c] = A[c * 6 + r]; a[r][6 + c] = (r == c) ? 1.0 : 0.0; }
  for (int k = 0; k < 6; ++k) {
    int p = k;
    for (int r = k + 1; r < 6; ++r) if (std::fabs(a[r][k]) > std::fabs(a[p][k])) p = r;
    if (!(std::fabs(a[p][k]) > 0.0)) return false;
    if (p != k) for (int c = 0; c < 12; ++c) std::swap(a[p][c], a[k][c]);
    const double inv = 1.0 / a[k][k];
    for (int c = 0; c < 12; ++c) a[k][c] *= inv;
    for (int r = 0; r < 6; ++r) {
      if (r == k) continue;
      const double f = a[r][k];
      if (f != 0.0) for (int c = 0; c < 12; ++c) a[r][c] -= f * a[k][c];
    }
  }
  for (int r = 0; r < 6; ++r)
    for (int c = 0; c < 6; ++c) Ai[c * 6 + r] = a[r][6 + c];
  return true;
}
}  // namespace

extern "C" {

/* Runs the reference function.  n_model must be >= n_data (the reference indexes model_qi[s] for
 * s < data_pi.size(), cov.h:45-51).  Outputs: live_cov[9] row-major; H6[36] column-major d2J_dX2;
 * D (column-major 6 x 6n, capacity d_cap doubles, may be NULL); returns n (columns/6) or -1.      */
int ref_cov_live(const float *data_xy, int n_data, const float *model_xy, int n_model, const float T_colmajor[16],
                 float sx, float sy, float st, double live_cov[9], double H6[36], double *D, long d_cap,
                 int *cov_z_dim) {
  if (n_model < n_data) return -1;
  Cloud::Ptr p = make_cloud(data_xy, n_data), q = make_cloud(model_xy, n_model);
  Eigen::Matrix4f T;
  std::memcpy(T.m, T_colmajor, sizeof(T.m));
  Eigen::MatrixXd out;
  for (auto &s : Eigen::tap::ring) { s.rows = s.cols = 0; s.data.clear(); }
  Eigen::tap::enabled = true;
  calculate_ICP_COV(p, q, T, out, sx, sy, st);
  Eigen::tap::enabled = false;
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) live_cov[3 * r + c] = out(r, c);
  /* locals die in reverse declaration order: cov_z, d2J_dZdX, d2J_dX2 -> ring[2], ring[1], ring[0] */
  const Eigen::tap::Slot &h = Eigen::tap::ring[0], &d = Eigen::tap::ring[1], &z = Eigen::tap::ring[2];
  if (h.rows != 6 || h.cols != 6) return -1;
  std::memcpy(H6, h.data.data(), 36 * sizeof(double));
  int n = 0;
  if (d.rows == 6 && d.cols % 6 == 0) n = d.cols / 6;          /* n == 0: empty matrices are not tapped */
  if (cov_z_dim) *cov_z_dim = z.rows == z.cols ? z.rows : -1;
  if (D && (long)d.data.size() <= d_cap && n > 0) std::memcpy(D, d.data.data(), d.data.size() * sizeof(double));
  return n;
}

/* cov.h:560 + 564-566 evaluated on the matrices the reference computed. cov_z = sensor_var * I. */
int ref_cov_intended(const float *data_xy, int n_data, const float *model_xy, int n_model, const float T_colmajor[16],
                     double sensor_var, double cov3[9], double H3[9]) {
  double live[9], H6[36];
  std::vector<double> D((size_t)36 * 200 + 36);
  int zdim = 0;
  const int n = ref_cov_live(data_xy, n_data, model_xy, n_model, T_colmajor, 0.f, 0.f, 0.f, live, H6, D.data(),
                             (long)D.size(), &zdim);
  if (n <= 0) return -1;
  const int sel[3] = {0, 1, 3};
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) {
      cov3[3 * r + c] = 0.0;
      if (H3) H3[3 * r + c] = H6[sel[c] * 6 + sel[r]];
    }
  double Hi[36];
  if (!invert6(H6, Hi)) return -2;
  const int m = 6 * n;
  /* M = D * (sensor_var I) * D^T  (6x6) */
  double M[36] = {0};
  for (int c = 0; c < m; ++c)
    for (int i = 0; i < 6; ++i) {
      const double dic = D[(size_t)c * 6 + i];
      for (int j = 0; j < 6; ++j) M[j * 6 + i] += dic * sensor_var * D[(size_t)c * 6 + j];
    }
  double A[36] = {0}, B[36] = {0};
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 6; ++j) { double s = 0; for (int k = 0; k < 6; ++k) s += Hi[k * 6 + i] * M[j * 6 + k]; A[j * 6 + i] = s; }
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 6; ++j) { double s = 0; for (int k = 0; k < 6; ++k) s += A[k * 6 + i] * Hi[j * 6 + k]; B[j * 6 + i] = s; }
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) cov3[3 * r + c] = B[sel[c] * 6 + sel[r]];
  return n;
}

/* math_utils.cc:20-34 / 5-18 and math_utils.h:13-16, the reference's own code */
void ref_inverse_transform_point(const float src_pt[2], float src_angle, const float tgt_pos[2], float tgt_angle,
                                 float out[3]) {
  auto r = math_utils::inverseTransformPoint(Eigen::Vector2f(src_pt[0], src_pt[1]), src_angle,
                                             Eigen::Vector2f(tgt_pos[0], tgt_pos[1]), tgt_angle);
  out[0] = r.first.x(); out[1] = r.first.y(); out[2] = r.second;
}
void ref_transform_point(const float src_pt[2], float src_angle, const float pos[2], float angle, float out[3]) {
  auto r = math_utils::transformPoint(Eigen::Vector2f(src_pt[0], src_pt[1]), src_angle,
                                      Eigen::Vector2f(pos[0], pos[1]), angle);
  out[0] = r.first.x(); out[1] = r.first.y(); out[2] = r.second;
}
float ref_angle_mod(float a) { return math_utils::AngleMod<float>(a); }

}  /* extern "C" */
