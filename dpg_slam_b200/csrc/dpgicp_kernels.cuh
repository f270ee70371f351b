/*
 * dpgicp_kernels.cuh — sm_100a device code of the scan-matching back end.
 *
 * One persistent CTA aligns one scan pair at a time, pulled from an atomic work queue (ICP
 * iteration counts vary 5..500 per pair, so static assignment would idle SMs).  Both clouds of
 * the pair are staged once in shared memory (1-D TMA bulk copies when rows are contiguous) and
 * every ICP iteration — correspondence search, reciprocity check, moment reduction, rigid solve,
 * convergence test, source update — runs on-chip; only the 112-byte result record goes back to
 * HBM.  Replaces the inside of DpgSLAM::runIcp (reference src/dpg_slam/dpg_slam.cc:362-446):
 * PCL's kd-tree correspondence estimation + SVD step + convergence criteria
 * and calculate_ICP_COV (src/icp_cov/cov_func_point_to_point.h:24-585, planar closed form
 * SURVEY.md App. B).
 *
 * Arithmetic contract (identical to oracle/dpg_oracle.c): binary32 transform and
 * distances in individually rounded operations (no FMA: __f*_rn intrinsics and --fmad=false),
 * (d2, index)-lexicographic nearest neighbour, exact int64 fixed-point moment sums (order
 * independent, so any thread layout gives the same bits), binary64 closed-form planar step.
 *
 * Exact pruned search: points of a scan are in beam order, so 32 consecutive points form a
 * spatially compact group with an axis-aligned bounding box.  A warp handles 32 consecutive
 * queries (one per lane); box-to-box lower bounds (evaluated lane-parallel over groups) select
 * the groups that can still beat the tile's current bound, which is seeded with the previous
 * iteration's neighbour.  Lower bounds use the same rounding sequence as the distance itself, so
 * by monotonicity of rounding they never exceed a computed distance: pruning is exact, no
 * epsilons.  SEARCH_BRUTE scans every group through the same code.
 */
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "dpgicp.h"

namespace dpg {

constexpr int   kGroup = 32;        /* points per bounding-box group = lanes per warp            */
constexpr float kPad   = 1.0e30f;   /* coordinate of padded slots: any d2 against it is +inf      */
constexpr double kScaleLin  = 4294967296.0;      /* 2^32 */
constexpr double kScaleProd = 268435456.0;       /* 2^28 */
constexpr double kScaleD2   = 1099511627776.0;   /* 2^40 */

struct PairTask {            /* 24 bytes: pair indices + the guess as matrix entries (host libm) */
  int32_t src, tgt;
  float c, s, tx, ty;
};

struct StoreView {           /* scan store: padded rows of float2, beam order, MAX_RANGE dropped  */
  const float2 *pts;
  const int32_t *count;
  int32_t pitch;             /* points per row, even (rows are 16-byte aligned for TMA)           */
  int32_t n_scans;
};

struct KernelParams {
  StoreView store;
  const PairTask *tasks;
  dpgicp_result *results;
  unsigned long long *queue;      /* work-queue head                                              */
  unsigned long long *counters;   /* [0] iterations [1] correspondences [2] distance evals [3] box tests */
  long long n_pairs;
  int32_t n_cap;                  /* smem capacity per cloud in points, multiple of 32            */
  int32_t max_iterations, use_reciprocal, divisor, metric, cov_mode, cov_cap;
  float gate;                     /* binary32 floor of max_correspondence_distance^2              */
  double eps, rot_thr, sensor_var;
  float live[3];
  /* dpgicp_correspondences hook: when corr_out != nullptr the kernel runs ONE pass for pair 0    */
  int32_t *corr_out;
  float *corr_d2_out;
};

/* ------------------------------------------------------------------------------------------------
 * small device helpers
 * ---------------------------------------------------------------------------------------------- */
__device__ __forceinline__ float dist2(float ax, float ay, float bx, float by) {
  const float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by);
  return __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
}

__device__ __forceinline__ float2 xform(float c, float s, float tx, float ty, float2 p) {
  float2 r;
  r.x = __fadd_rn(__fadd_rn(__fmul_rn(c, p.x), __fmul_rn(-s, p.y)), tx);
  r.y = __fadd_rn(__fadd_rn(__fmul_rn(s, p.x), __fmul_rn(c, p.y)), ty);
  return r;
}

/* Blackwell warp-wide float min/max in one instruction (SASS CREDUX.MIN/MAX.F32) */
__device__ __forceinline__ float warp_min(float v) {
  float r;
  asm volatile("redux.sync.min.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));
  return r;
}
__device__ __forceinline__ float warp_max(float v) {
  float r;
  asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));
  return r;
}

__device__ __forceinline__ long long warp_sum_i64(long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = __dadd_rn(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

/* lower bound of dist2(q, p) over all p in box b = (lox, loy, hix, hiy); same rounding sequence */
__device__ __forceinline__ float lb_point_box(float qx, float qy, float4 b) {
  const float ex = fmaxf(fmaxf(__fsub_rn(b.x, qx), __fsub_rn(qx, b.z)), 0.0f);
  const float ey = fmaxf(fmaxf(__fsub_rn(b.y, qy), __fsub_rn(qy, b.w)), 0.0f);
  return __fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey));
}
/* lower bound over all q in box a, p in box b */
__device__ __forceinline__ float lb_box_box(float4 a, float4 b) {
  const float ex = fmaxf(fmaxf(__fsub_rn(b.x, a.z), __fsub_rn(a.x, b.z)), 0.0f);
  const float ey = fmaxf(fmaxf(__fsub_rn(b.y, a.w), __fsub_rn(a.y, b.w)), 0.0f);
  return __fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey));
}

/* ---- mbarrier + 1-D TMA bulk copy (SASS UBLKCP) ------------------------------------------------ */
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

/* ------------------------------------------------------------------------------------------------
 * shared-memory layout of one CTA
 * ---------------------------------------------------------------------------------------------- */
struct SmemLayout {
  float2 *tgt;      /* n_cap */
  float2 *src;      /* n_cap, current (incrementally transformed) source                           */
  int32_t *nn;      /* n_cap, forward neighbour of the previous pass (seed) / final correspondences */
  float4 *tbox;     /* n_cap/32 */
  float4 *sbox;     /* n_cap/32 */
  int32_t *tcnt;    /* n_cap/32 accepted per source tile (rank for the covariance cap)              */
  long long *red;   /* 16 int64 moment sums                                                         */
  double *dpart;    /* kMaxWarps * 12 partial double sums (deterministic order)                     */
  float *step;      /* 4 */
  int32_t *ctl;     /* [0] pair index lo [1] pair index hi [2] stop [3] K                            */
  uint64_t *mbar;
};
constexpr int kMaxWarps = 16;

__host__ __device__ inline size_t smem_bytes(int n_cap) {
  const int g = n_cap / kGroup;
  return (size_t)n_cap * (8 + 8 + 4) + (size_t)g * (16 + 16 + 4) + 16 * 8 + kMaxWarps * 12 * 8 + 16 + 32 + 16 + 64;
}

__device__ __forceinline__ SmemLayout carve(unsigned char *base, int n_cap) {
  const int g = n_cap / kGroup;
  SmemLayout L;
  size_t o = 0;
  L.tgt = (float2 *)(base + o);  o += (size_t)n_cap * 8;
  L.src = (float2 *)(base + o);  o += (size_t)n_cap * 8;
  L.tbox = (float4 *)(base + o); o += (size_t)g * 16;
  L.sbox = (float4 *)(base + o); o += (size_t)g * 16;
  L.red = (long long *)(base + o); o += 16 * 8;
  L.dpart = (double *)(base + o);  o += kMaxWarps * 12 * 8;
  L.mbar = (uint64_t *)(base + o); o += 16;
  L.nn = (int32_t *)(base + o);   o += (size_t)n_cap * 4;
  L.tcnt = (int32_t *)(base + o); o += (size_t)g * 4;
  L.step = (float *)(base + o);   o += 16;
  L.ctl = (int32_t *)(base + o);  o += 32;
  return L;
}

/* bounding box of the 32 points held one per lane (invalid lanes contribute nothing) */
__device__ __forceinline__ float4 warp_box(float2 p, bool valid) {
  const float inf = __int_as_float(0x7f800000);
  float4 b;
  b.x = warp_min(valid ? p.x : inf);
  b.y = warp_min(valid ? p.y : inf);
  b.z = warp_max(valid ? p.x : -inf);
  b.w = warp_max(valid ? p.y : -inf);
  return b;
}

/* ------------------------------------------------------------------------------------------------
 * Exact nearest neighbour of one query per lane over a grouped cloud in shared memory.
 *   bd/bj in: current bound (gate or seed), out: (d2, index)-lexicographic minimum among points
 *   with d2 <= initial bd.  `qbox` is the bounding box of the valid lanes' queries.
 *   PRUNED = false scans every group.
 * ---------------------------------------------------------------------------------------------- */
template <bool PRUNED>
__device__ __forceinline__ void nn_search(const float2 *__restrict__ cloud, const float4 *__restrict__ boxes,
                                          int n_groups, float qx, float qy, bool valid, float4 qbox,
                                          float &bd, int &bj, unsigned &scans, unsigned &tests) {
  const int lane = threadIdx.x & 31;
  float bmax = 0.0f;
  if (PRUNED) bmax = warp_max(valid ? bd : -1.0f);
  for (int base = 0; base < n_groups; base += 32) {
    unsigned mask;
    if (PRUNED) {
      const int g = base + lane;
      bool cand = false;
      if (g < n_groups) cand = lb_box_box(qbox, boxes[g]) <= bmax;
      mask = __ballot_sync(0xffffffffu, cand);
      ++tests;
    } else {
      const int rem = n_groups - base;
      mask = rem >= 32 ? 0xffffffffu : ((1u << rem) - 1u);
    }
    while (mask) {
      const int g = base + __ffs(mask) - 1;
      mask &= mask - 1;
      if (PRUNED) {
        const bool need = valid && (lb_point_box(qx, qy, boxes[g]) <= bd);
        if (!__any_sync(0xffffffffu, need)) continue;
      }
      ++scans;
      const float4 *pp = reinterpret_cast<const float4 *>(cloud + g * kGroup);
      float gd = __int_as_float(0x7f800000);
      int gj = 0;
#pragma unroll
      for (int t = 0; t < kGroup / 2; ++t) {
        const float4 p = pp[t];                       /* two points per LDS.128, broadcast       */
        const float d0 = dist2(qx, qy, p.x, p.y);
        const float d1 = dist2(qx, qy, p.z, p.w);
        if (d0 < gd) { gd = d0; gj = 2 * t; }
        if (d1 < gd) { gd = d1; gj = 2 * t + 1; }
      }
      const int j = g * kGroup + gj;
      if (gd < bd || (gd == bd && j < bj)) { bd = gd; bj = j; }
    }
  }
}

/* Is there a point i' in the grouped cloud with (d2(i', r), i') < (bd, self) lexicographically?
 * (reciprocity test of PCL's determineReciprocalCorrespondences: the query r = tgt[j] must have the
 * source point `self` as ITS nearest neighbour.) */
template <bool PRUNED>
__device__ __forceinline__ bool beaten_search(const float2 *__restrict__ cloud, const float4 *__restrict__ boxes,
                                              int n_groups, float rx, float ry, bool active, float4 rbox,
                                              float bd, int self, unsigned &scans, unsigned &tests) {
  const int lane = threadIdx.x & 31;
  bool beaten = false;
  float bmax = 0.0f;
  if (PRUNED) bmax = warp_max(active ? bd : -1.0f);
  for (int base = 0; base < n_groups; base += 32) {
    unsigned mask;
    if (PRUNED) {
      const int g = base + lane;
      bool cand = false;
      if (g < n_groups) cand = lb_box_box(rbox, boxes[g]) <= bmax;
      mask = __ballot_sync(0xffffffffu, cand);
      ++tests;
    } else {
      const int rem = n_groups - base;
      mask = rem >= 32 ? 0xffffffffu : ((1u << rem) - 1u);
    }
    while (mask) {
      const int g = base + __ffs(mask) - 1;
      mask &= mask - 1;
      if (PRUNED) {
        const bool need = active && !beaten && (lb_point_box(rx, ry, boxes[g]) <= bd);
        if (!__any_sync(0xffffffffu, need)) continue;
      }
      ++scans;
      const float4 *pp = reinterpret_cast<const float4 *>(cloud + g * kGroup);
      /* lowest distance in the group and the first index attaining it */
      float gd = __int_as_float(0x7f800000);
      int gj = 0;
#pragma unroll
      for (int t = 0; t < kGroup / 2; ++t) {
        const float4 p = pp[t];
        const float d0 = dist2(p.x, p.y, rx, ry);
        const float d1 = dist2(p.z, p.w, rx, ry);
        if (d0 < gd) { gd = d0; gj = 2 * t; }
        if (d1 < gd) { gd = d1; gj = 2 * t + 1; }
      }
      const int i2 = g * kGroup + gj;
      if (gd < bd || (gd == bd && i2 < self)) beaten = true;
    }
  }
  return beaten;
}

/* ------------------------------------------------------------------------------------------------
 * One correspondence pass for one source tile (32 consecutive source points, one per lane).
 * Returns accept flag; j = matched target index, d = its squared distance.  Updates the seed.
 * ---------------------------------------------------------------------------------------------- */
template <bool PRUNED>
__device__ __forceinline__ bool match_tile(const SmemLayout &L, int tile, int ns, int n_groups_s,
                                           int n_groups_t, float gate, bool reciprocal, float2 &q,
                                           int &j_out, float &d_out, bool &fwd_ok, unsigned &scans,
                                           unsigned &tests) {
  const int lane = threadIdx.x & 31;
  const int i = tile * kGroup + lane;
  const bool valid = i < ns;
  q = L.src[i];
  float bd = gate;
  int bj = 0x7fffffff;
  if (PRUNED) {
    const int seed = valid ? L.nn[i] : -1;
    if (seed >= 0) {
      const float2 p = L.tgt[seed];
      const float d0 = dist2(q.x, q.y, p.x, p.y);
      if (d0 <= gate) { bd = d0; bj = seed; }
    }
  }
  nn_search<PRUNED>(L.tgt, L.tbox, n_groups_t, q.x, q.y, valid, L.sbox[tile], bd, bj, scans, tests);
  fwd_ok = valid && (bj != 0x7fffffff);
  j_out = bj;
  d_out = bd;
  bool accept = fwd_ok;
  if (reciprocal) {
    float2 r = make_float2(0.f, 0.f);
    if (fwd_ok) r = L.tgt[bj];
    const float4 rbox = warp_box(r, fwd_ok);
    if (__any_sync(0xffffffffu, fwd_ok)) {
      const bool beaten =
          beaten_search<PRUNED>(L.src, L.sbox, n_groups_s, r.x, r.y, fwd_ok, rbox, bd, i, scans, tests);
      accept = fwd_ok && !beaten;
    }
  }
  return accept;
}

/* ------------------------------------------------------------------------------------------------
 * covariance finishing (3x3 algebra), mirrors oracle/dpg_oracle.c orc_cov_censi
 * sums: [0] n_h [1] SA [2] SB [3] SE   (Hessian, all pairs)
 *       [4] n_d [5] Sdx [6] Sdy [7] SA_d [8] SB_d [9] S(dx^2+dy^2) [10] S(A^2+B^2)  (capped set)
 * ---------------------------------------------------------------------------------------------- */
__device__ inline uint32_t finish_cov(const double *S, double sensor_var, const float *live, double *cov) {
  double H[9], M[9], Hi[9];
  H[0] = 2.0 * S[0]; H[1] = 0.0;        H[2] = -2.0 * S[2];
  H[3] = 0.0;        H[4] = 2.0 * S[0]; H[5] = 2.0 * S[1];
  H[6] = H[2];       H[7] = H[5];       H[8] = -2.0 * S[3];
  M[0] = 8.0 * S[4]; M[1] = 0.0;        M[2] = 4.0 * (S[6] - S[8]);
  M[3] = 0.0;        M[4] = 8.0 * S[4]; M[5] = 4.0 * (S[7] - S[5]);
  M[6] = M[2];       M[7] = M[5];       M[8] = 4.0 * S[9] + 4.0 * S[10];
  bool ok = S[0] > 0.0;
  const double a = H[0], b = H[1], c = H[2], d = H[4], e = H[5], f = H[8];
  const double c00 = d * f - e * e, c01 = c * e - b * f, c02 = b * e - c * d;
  const double det = a * c00 + b * c01 + c * c02;
  if (!(fabs(det) > 0.0) || !isfinite(det)) ok = false;
  if (ok) {
    const double id = 1.0 / det;
    Hi[0] = c00 * id; Hi[1] = c01 * id; Hi[2] = c02 * id;
    Hi[3] = Hi[1];    Hi[4] = (a * f - c * c) * id; Hi[5] = (b * c - a * e) * id;
    Hi[6] = Hi[2];    Hi[7] = Hi[5];    Hi[8] = (a * d - b * b) * id;
    double Tm[9];
    for (int r = 0; r < 3; ++r)
      for (int q = 0; q < 3; ++q)
        Tm[3 * r + q] = Hi[3 * r] * M[q] + Hi[3 * r + 1] * M[3 + q] + Hi[3 * r + 2] * M[6 + q];
    for (int r = 0; r < 3; ++r)
      for (int q = 0; q < 3; ++q)
        cov[3 * r + q] =
            sensor_var * (Tm[3 * r] * Hi[q] + Tm[3 * r + 1] * Hi[3 + q] + Tm[3 * r + 2] * Hi[6 + q]);
    for (int k = 0; k < 9; ++k)
      if (!isfinite(cov[k])) ok = false;
  }
  if (!ok) {
    for (int k = 0; k < 9; ++k) cov[k] = 0.0;
    cov[0] = live[0]; cov[4] = live[1]; cov[8] = live[2];
    return DPGICP_FLAG_COV_SINGULAR;
  }
  return 0u;
}

/* per-pair terms of the covariance sums for source point p (untransformed) and target q */
__device__ __forceinline__ void cov_terms(double px, double py, double qx, double qy, double c, double s,
                                          double x, double y, bool in_h, bool in_d, double *acc) {
  const double A = px * c - py * s, B = px * s + py * c;
  const double dx = x - qx, dy = y - qy;
  if (in_h) {
    acc[0] += 1.0; acc[1] += A; acc[2] += B; acc[3] += A * dx + B * dy;
  }
  if (in_d) {
    acc[4] += 1.0; acc[5] += dx; acc[6] += dy; acc[7] += A; acc[8] += B;
    acc[9] += dx * dx + dy * dy; acc[10] += A * A + B * B;
  }
}

/* deterministic block reduction of 11 doubles: lanes by xor-shuffle, warps summed in order */
template <int WARPS>
__device__ __forceinline__ void block_sum11(double *acc, double *dpart, double *out /* thread 0 */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 11; ++k) acc[k] = warp_sum_f64(acc[k]);
  if (lane == 0)
    for (int k = 0; k < 11; ++k) dpart[warp * 12 + k] = acc[k];
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 0; k < 11; ++k) {
      double v = 0.0;
      for (int w = 0; w < WARPS; ++w) v = __dadd_rn(v, dpart[w * 12 + k]);
      out[k] = v;
    }
  }
  __syncthreads();
}

/* ------------------------------------------------------------------------------------------------
 * the persistent ICP + covariance kernel
 * ---------------------------------------------------------------------------------------------- */
template <int WARPS, bool PRUNED>
__global__ void __launch_bounds__(WARPS * 32) icp_pairs_kernel(const KernelParams P) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const SmemLayout L = carve(smem_raw, P.n_cap);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int div = P.divisor;
  uint32_t mbar_phase = 0;

  if (tid == 0) mbar_init(L.mbar, 1);
  __syncthreads();

  unsigned c_scans = 0, c_tests = 0;
  unsigned long long c_iters = 0, c_corr = 0;

  for (;;) {
    /* ---- fetch the next pair ---------------------------------------------------------------- */
    if (tid == 0) {
      const unsigned long long k = atomicAdd(P.queue, 1ull);
      L.ctl[0] = (int32_t)(k & 0xffffffffu);
      L.ctl[1] = (int32_t)(k >> 32);
    }
    __syncthreads();
    const long long pair = (long long)(((unsigned long long)(uint32_t)L.ctl[1] << 32) | (uint32_t)L.ctl[0]);
    if (pair >= P.n_pairs) break;
    const PairTask task = P.tasks[pair];
    const float2 *srow = P.store.pts + (size_t)task.src * P.store.pitch;
    const float2 *trow = P.store.pts + (size_t)task.tgt * P.store.pitch;
    const int ns_full = P.store.count[task.src], nt_full = P.store.count[task.tgt];
    const int ns = (ns_full + div - 1) / div, nt = (nt_full + div - 1) / div;
    const int gs = (ns + kGroup - 1) / kGroup, gt = (nt + kGroup - 1) / kGroup;

    /* ---- stage both clouds in shared memory ------------------------------------------------- */
    if (div == 1) {
      /* contiguous rows: two 1-D TMA bulk copies completing on one mbarrier */
      fence_proxy_async();
      __syncthreads();
      if (tid == 0) {
        const uint32_t bt = (uint32_t)((nt * 8 + 15) & ~15), bs = (uint32_t)((ns * 8 + 15) & ~15);
        mbar_expect_tx(L.mbar, bt + bs);
        if (bt) bulk_g2s(L.tgt, trow, bt, L.mbar);
        if (bs) bulk_g2s(L.src, srow, bs, L.mbar);
      }
      mbar_wait(L.mbar, mbar_phase);
      mbar_phase ^= 1u;
    } else {
      for (int k = tid; k < nt; k += WARPS * 32) L.tgt[k] = __ldg(trow + (size_t)k * div);
      for (int k = tid; k < ns; k += WARPS * 32) L.src[k] = __ldg(srow + (size_t)k * div);
    }
    __syncthreads();
    /* pad to whole groups, apply the guess (PCL transformCloud(input, guess), App. A.2), boxes */
    for (int k = nt + tid; k < gt * kGroup; k += WARPS * 32) L.tgt[k] = make_float2(kPad, kPad);
    for (int k = ns + tid; k < gs * kGroup; k += WARPS * 32) L.src[k] = make_float2(kPad, kPad);
    __syncthreads();
    for (int g = warp; g < gt; g += WARPS) {
      const int k = g * kGroup + lane;
      const float4 b = warp_box(L.tgt[k], k < nt);
      if (lane == 0) L.tbox[g] = b;
    }
    for (int g = warp; g < gs; g += WARPS) {
      const int k = g * kGroup + lane;
      float2 p = L.src[k];
      if (k < ns) { p = xform(task.c, task.s, task.tx, task.ty, p); L.src[k] = p; }
      L.nn[k] = -1;
      const float4 b = warp_box(p, k < ns);
      if (lane == 0) L.sbox[g] = b;
    }
    if (tid < 16) L.red[tid] = 0;
    __syncthreads();

    /* ---- parity hook: a single correspondence pass ------------------------------------------ */
    if (P.corr_out != nullptr) {
      for (int tile = warp; tile < gs; tile += WARPS) {
        float2 q; int j; float d; bool fwd;
        const bool acc = match_tile<PRUNED>(L, tile, ns, gs, gt, P.gate, P.use_reciprocal != 0, q, j, d, fwd,
                                            c_scans, c_tests);
        const int i = tile * kGroup + lane;
        if (i < ns) {
          P.corr_out[i] = acc ? j : -1;
          P.corr_d2_out[i] = fwd ? d : __int_as_float(0x7f800000);
        }
      }
      continue;
    }

    /* ---- ICP iterations (PCL IterativeClosestPoint::computeTransformation, App. A.3) -------- */
    float fc = task.c, fs = task.s, ftx = task.tx, fty = task.ty;   /* thread 0: final transform  */
    int iterations = 0, last_k = 0;
    uint32_t status = 0;
    double mse = 0.0, mse_prev = 1.7976931348623157e308;
    if (ns <= 0 || nt <= 0) status |= DPGICP_FLAG_EMPTY_INPUT;

    for (;;) {
      long long m_px = 0, m_py = 0, m_qx = 0, m_qy = 0, m_xx = 0, m_xy = 0, m_yx = 0, m_yy = 0, m_d2 = 0;
      int m_k = 0;
      for (int tile = warp; tile < gs; tile += WARPS) {
        float2 q; int j; float d; bool fwd;
        const bool acc = match_tile<PRUNED>(L, tile, ns, gs, gt, P.gate, P.use_reciprocal != 0, q, j, d, fwd,
                                            c_scans, c_tests);
        const int i = tile * kGroup + lane;
        if (i < ns) L.nn[i] = fwd ? j : -1;           /* seed of the next pass */
        if (acc) {
          const float2 t = L.tgt[j];
          const double px = q.x, py = q.y, qx = t.x, qy = t.y;
          m_px += __double2ll_rn(__dmul_rn(px, kScaleLin));
          m_py += __double2ll_rn(__dmul_rn(py, kScaleLin));
          m_qx += __double2ll_rn(__dmul_rn(qx, kScaleLin));
          m_qy += __double2ll_rn(__dmul_rn(qy, kScaleLin));
          m_xx += __double2ll_rn(__dmul_rn(__dmul_rn(px, qx), kScaleProd));
          m_xy += __double2ll_rn(__dmul_rn(__dmul_rn(px, qy), kScaleProd));
          m_yx += __double2ll_rn(__dmul_rn(__dmul_rn(py, qx), kScaleProd));
          m_yy += __double2ll_rn(__dmul_rn(__dmul_rn(py, qy), kScaleProd));
          m_d2 += __double2ll_rn(__dmul_rn((double)d, kScaleD2));
          m_k += 1;
        }
      }
      /* exact integer reduction: warp shuffles, then one shared atomic per warp and value */
      m_px = warp_sum_i64(m_px); m_py = warp_sum_i64(m_py); m_qx = warp_sum_i64(m_qx);
      m_qy = warp_sum_i64(m_qy); m_xx = warp_sum_i64(m_xx); m_xy = warp_sum_i64(m_xy);
      m_yx = warp_sum_i64(m_yx); m_yy = warp_sum_i64(m_yy); m_d2 = warp_sum_i64(m_d2);
      m_k = __reduce_add_sync(0xffffffffu, m_k);
      if (lane == 0) {
        unsigned long long *r = reinterpret_cast<unsigned long long *>(L.red);
        atomicAdd(r + 0, (unsigned long long)m_px); atomicAdd(r + 1, (unsigned long long)m_py);
        atomicAdd(r + 2, (unsigned long long)m_qx); atomicAdd(r + 3, (unsigned long long)m_qy);
        atomicAdd(r + 4, (unsigned long long)m_xx); atomicAdd(r + 5, (unsigned long long)m_xy);
        atomicAdd(r + 6, (unsigned long long)m_yx); atomicAdd(r + 7, (unsigned long long)m_yy);
        atomicAdd(r + 8, (unsigned long long)m_d2); atomicAdd(r + 9, (unsigned long long)(long long)m_k);
      }
      __syncthreads();

      if (tid == 0) {
        /* rigid step: planar Procrustes in binary64 (PCL TransformationEstimationSVD, z = 0) */
        const int K = (int)L.red[9];
        last_k = K;
        int stop = 0;
        if (K < 3) {                                   /* App. A.3-4 */
          status |= DPGICP_STOP_NO_CORRESPONDENCES;
          stop = 2;
        } else {
          const double Kd = (double)K;
          const double spx = __dmul_rn((double)L.red[0], 1.0 / kScaleLin);
          const double spy = __dmul_rn((double)L.red[1], 1.0 / kScaleLin);
          const double sqx = __dmul_rn((double)L.red[2], 1.0 / kScaleLin);
          const double sqy = __dmul_rn((double)L.red[3], 1.0 / kScaleLin);
          const double dot = __dmul_rn((double)(L.red[4] + L.red[7]), 1.0 / kScaleProd);
          const double crs = __dmul_rn((double)(L.red[5] - L.red[6]), 1.0 / kScaleProd);
          const double a = __dsub_rn(dot, __ddiv_rn(__dadd_rn(__dmul_rn(spx, sqx), __dmul_rn(spy, sqy)), Kd));
          const double b = __dsub_rn(crs, __ddiv_rn(__dsub_rn(__dmul_rn(spx, sqy), __dmul_rn(spy, sqx)), Kd));
          const double h = __dsqrt_rn(__dadd_rn(__dmul_rn(a, a), __dmul_rn(b, b)));
          double c = 1.0, s = 0.0;
          if (h > 0.0) { c = __ddiv_rn(a, h); s = __ddiv_rn(b, h); }
          const double mpx = __ddiv_rn(spx, Kd), mpy = __ddiv_rn(spy, Kd);
          const double mqx = __ddiv_rn(sqx, Kd), mqy = __ddiv_rn(sqy, Kd);
          const double tx = __dsub_rn(mqx, __dsub_rn(__dmul_rn(c, mpx), __dmul_rn(s, mpy)));
          const double ty = __dsub_rn(mqy, __dadd_rn(__dmul_rn(s, mpx), __dmul_rn(c, mpy)));
          const float sc = (float)c, ss = (float)s, stx = (float)tx, sty = (float)ty;
          L.step[0] = sc; L.step[1] = ss; L.step[2] = stx; L.step[3] = sty;
          /* final = step * final (App. A.3-6) */
          const float nc = __fadd_rn(__fmul_rn(sc, fc), __fmul_rn(-ss, fs));
          const float nsn = __fadd_rn(__fmul_rn(ss, fc), __fmul_rn(sc, fs));
          const float ntx = __fadd_rn(__fadd_rn(__fmul_rn(sc, ftx), __fmul_rn(-ss, fty)), stx);
          const float nty = __fadd_rn(__fadd_rn(__fmul_rn(ss, ftx), __fmul_rn(sc, fty)), sty);
          fc = nc; fs = nsn; ftx = ntx; fty = nty;
          ++iterations;
          mse = __ddiv_rn(__dmul_rn((double)L.red[8], 1.0 / kScaleD2), Kd);
          /* DefaultConvergenceCriteria (App. A.5), in PCL's order */
          const float tr = __fsub_rn(__fadd_rn(__fadd_rn(sc, sc), 1.0f), 1.0f);
          const double cos_angle = __dmul_rn(0.5, (double)tr);
          const float tsq = __fadd_rn(__fmul_rn(stx, stx), __fmul_rn(sty, sty));
          if (iterations >= P.max_iterations) {
            status |= DPGICP_STOP_ITERATIONS | DPGICP_FLAG_CONVERGED; stop = 1;
          } else if (cos_angle >= P.rot_thr && (double)tsq <= P.eps) {
            status |= DPGICP_STOP_TRANSFORM | DPGICP_FLAG_CONVERGED; stop = 1;
          } else if (fabs(__dsub_rn(mse, mse_prev)) < 1e-12) {
            status |= DPGICP_STOP_ABS_MSE | DPGICP_FLAG_CONVERGED; stop = 1;
          } else {
            mse_prev = mse;
          }
        }
        L.ctl[2] = stop;
        L.ctl[3] = K;
#pragma unroll
        for (int k = 0; k < 10; ++k) L.red[k] = 0;
      }
      __syncthreads();
      const int stop = L.ctl[2];
      c_corr += (tid == 0) ? (unsigned long long)L.ctl[3] : 0ull;
      if (stop == 2) break;
      /* src' = step * src' in place (App. A.3-6) and refresh the source boxes */
      {
        const float sc = L.step[0], ss = L.step[1], stx = L.step[2], sty = L.step[3];
        for (int g = warp; g < gs; g += WARPS) {
          const int k = g * kGroup + lane;
          float2 p = L.src[k];
          if (k < ns) { p = xform(sc, ss, stx, sty, p); L.src[k] = p; }
          const float4 b = warp_box(p, k < ns);
          if (lane == 0) L.sbox[g] = b;
        }
      }
      if (tid == 0) ++c_iters;
      __syncthreads();
      if (stop) break;
    }

    /* ---- covariance (calculate_ICP_COV) + result record -------------------------------------- */
    double S[11];
    uint32_t cov_flag = 0;
    const int cov_mode = P.cov_mode;
    if (cov_mode != DPGICP_COV_REFERENCE_LIVE) {
      /* pose as cov.h:26-35: x,y float entries widened, a = (double)atan2f(T10, T00); all threads */
      if (tid == 0) { L.step[0] = fc; L.step[1] = fs; L.step[2] = ftx; L.step[3] = fty; }
      __syncthreads();
      const float Tc = L.step[0], Ts = L.step[1], Ttx = L.step[2], Tty = L.step[3];
      const double x = (double)Ttx, y = (double)Tty;
      const double a = (double)atan2f(Ts, Tc);
      const double ca = cos(a), sa = sin(a);
      double acc[11];
#pragma unroll
      for (int k = 0; k < 11; ++k) acc[k] = 0.0;
      if (cov_mode == DPGICP_COV_CENSI_INDEXPAIR) {
        /* full clouds paired by index (dpg_slam.cc:430), Hessian over all, D-term over the cap */
        const int nh = ns_full < nt_full ? ns_full : nt_full;
        const int nd = (P.cov_cap > 0 && nh > P.cov_cap) ? P.cov_cap : nh;
        for (int k = tid; k < nh; k += WARPS * 32) {
          const float2 p = __ldg(srow + k), q = __ldg(trow + k);
          cov_terms(p.x, p.y, q.x, q.y, ca, sa, x, y, true, k < nd, acc);
        }
      } else {
        /* CENSI_CORR: correspondences at the final pose: src'' = final * src (original points) */
        __syncthreads();
        for (int g = warp; g < gs; g += WARPS) {
          const int k = g * kGroup + lane;
          float2 p = make_float2(kPad, kPad);
          if (k < ns) { p = xform(Tc, Ts, Ttx, Tty, __ldg(srow + (size_t)k * div)); }
          L.src[k] = p;
          const float4 b = warp_box(p, k < ns);
          if (lane == 0) L.sbox[g] = b;
        }
        __syncthreads();
        for (int tile = warp; tile < gs; tile += WARPS) {
          float2 q; int j; float d; bool fwd;
          const bool ok = match_tile<PRUNED>(L, tile, ns, gs, gt, P.gate, P.use_reciprocal != 0, q, j, d, fwd,
                                             c_scans, c_tests);
          const unsigned bal = __ballot_sync(0xffffffffu, ok);
          const int i = tile * kGroup + lane;
          if (i < ns) L.nn[i] = ok ? j : -1;
          if (lane == 0) L.tcnt[tile] = __popc(bal);
        }
        __syncthreads();
        for (int tile = warp; tile < gs; tile += WARPS) {
          int prefix = 0;
          for (int t = lane; t < tile; t += 32) prefix += L.tcnt[t];
          prefix = __reduce_add_sync(0xffffffffu, prefix);
          const int i = tile * kGroup + lane;
          const int j = (i < ns) ? L.nn[i] : -1;
          const unsigned bal = __ballot_sync(0xffffffffu, j >= 0);
          const int rank = prefix + __popc(bal & ((1u << lane) - 1u));
          if (j >= 0) {
            const float2 p = __ldg(srow + (size_t)i * div);
            const float2 q = L.tgt[j];
            const bool in_d = (P.cov_cap <= 0) || (rank < P.cov_cap);
            cov_terms(p.x, p.y, q.x, q.y, ca, sa, x, y, true, in_d, acc);
          }
        }
      }
      block_sum11<WARPS>(acc, L.dpart, S);
    }

    if (tid == 0) {
      dpgicp_result r;
      r.tx = ftx; r.ty = fty;
      r.theta = atan2f(fs, fc);                      /* Rotation2Df::fromRotationMatrix().angle() */
      r.rot_c = fc; r.rot_s = fs;
      r.iterations = iterations;
      r.n_correspondences = last_k;
      r.mse = mse;
#pragma unroll
      for (int k = 0; k < 9; ++k) r.cov[k] = 0.0;
      if (cov_mode == DPGICP_COV_REFERENCE_LIVE) {   /* cov.h:572-575 */
        r.cov[0] = P.live[0]; r.cov[4] = P.live[1]; r.cov[8] = P.live[2];
      } else {
        cov_flag = finish_cov(S, P.sensor_var, P.live, r.cov);
      }
      r.status = status | cov_flag;
      P.results[pair] = r;
    }
  }

  /* executed-work counters (one set of atomics per CTA) */
  if (lane == 0) {
    atomicAdd(P.counters + 2, (unsigned long long)c_scans * (kGroup * 32ull));
    atomicAdd(P.counters + 3, (unsigned long long)c_tests * 32ull);
  }
  if (tid == 0) {
    atomicAdd(P.counters + 0, c_iters);
    atomicAdd(P.counters + 1, c_corr);
  }
}

/* ------------------------------------------------------------------------------------------------
 * Standalone covariance kernel (calculate_ICP_COV call shape): one CTA per item, clouds paired by
 * index.  HBM-bound: 16 bytes read per index pair.
 * ---------------------------------------------------------------------------------------------- */
struct CovItem {
  const float2 *p, *q;      /* data_pi, model_qi (packed float2) */
  int32_t n_p, n_q;
  float c, s, tx, ty;       /* T(0,0), T(1,0), T(0,3), T(1,3)    */
};

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) cov_indexpair_kernel(const CovItem *items, int n_items, int cov_mode,
                                                                  int cov_cap, double sensor_var, float lx,
                                                                  float ly, float lt, double *cov_out,
                                                                  uint32_t *status_out) {
  __shared__ double dpart[kMaxWarps * 12];
  const int item = blockIdx.x;
  if (item >= n_items) return;
  const CovItem it = items[item];
  const float live[3] = {lx, ly, lt};
  double S[11];
  double acc[11];
#pragma unroll
  for (int k = 0; k < 11; ++k) acc[k] = 0.0;
  if (cov_mode != DPGICP_COV_REFERENCE_LIVE) {
    const double x = (double)it.tx, y = (double)it.ty;
    const double a = (double)atan2f(it.s, it.c);
    const double ca = cos(a), sa = sin(a);
    const int nh = it.n_p < it.n_q ? it.n_p : it.n_q;
    const int nd = (cov_cap > 0 && nh > cov_cap) ? cov_cap : nh;
    for (int k = threadIdx.x; k < nh; k += WARPS * 32) {
      const float2 p = __ldg(it.p + k), q = __ldg(it.q + k);
      cov_terms(p.x, p.y, q.x, q.y, ca, sa, x, y, true, k < nd, acc);
    }
  }
  block_sum11<WARPS>(acc, dpart, S);
  if (threadIdx.x == 0) {
    double cov[9];
    uint32_t flag = 0;
    if (cov_mode == DPGICP_COV_REFERENCE_LIVE) {
      for (int k = 0; k < 9; ++k) cov[k] = 0.0;
      cov[0] = lx; cov[4] = ly; cov[8] = lt;
    } else {
      flag = finish_cov(S, sensor_var, live, cov);
    }
    for (int k = 0; k < 9; ++k) cov_out[(size_t)item * 9 + k] = cov[k];
    status_out[item] = flag;
  }
}

/* ------------------------------------------------------------------------------------------------
 * Scan-store builders
 * ---------------------------------------------------------------------------------------------- */
/* CSR points (arbitrary stride) -> padded rows; flags non-finite / out-of-range coordinates */
__global__ void pack_rows_kernel(const unsigned char *__restrict__ pts, size_t stride, const long long *offsets,
                                 int n_scans, int pitch, float2 *rows, int32_t *count, int *bad) {
  const int scan = blockIdx.x;
  if (scan >= n_scans) return;
  const long long o0 = offsets[scan], o1 = offsets[scan + 1];
  const int n = (int)(o1 - o0);
  if (threadIdx.x == 0) count[scan] = n;
  for (int k = threadIdx.x; k < pitch; k += blockDim.x) {
    float2 v = make_float2(0.f, 0.f);
    if (k < n) {
      const float *f = reinterpret_cast<const float *>(pts + (size_t)(o0 + k) * stride);
      v = make_float2(f[0], f[1]);
      if (!(fabsf(v.x) <= DPGICP_MAX_ABS_COORD) || !(fabsf(v.y) <= DPGICP_MAX_ABS_COORD)) atomicExch(bad, 1);
    }
    rows[(size_t)scan * pitch + k] = v;
  }
}

/* raw ranges -> base_link points, MAX_RANGE dropped, beam order kept (createNode
 * dpg_slam.cc:497-506, dpg_measurement.h:41-46,102-104, dpg_node.cc:13-22).  One warp per scan,
 * ballot-prefix compaction.  Trig in binary64 rounded to binary32, like the oracle. */
__global__ void ranges_to_rows_kernel(const float *__restrict__ ranges, int n_scans, int n_beams, float angle_min,
                                      float angle_inc, float range_max, float lx, float ly, float lc, float ls,
                                      int pitch, float2 *rows, int32_t *count, int *bad) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= n_scans) return;
  const float *r = ranges + (size_t)warp * n_beams;
  float2 *row = rows + (size_t)warp * pitch;
  int n = 0;
  for (int base = 0; base < n_beams; base += 32) {
    const int i = base + lane;
    float rg = range_max;
    if (i < n_beams) rg = __ldg(r + i);
    const bool keep = (i < n_beams) && !(rg >= range_max);
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    if (keep) {
      const float angle = __fadd_rn(__fmul_rn(angle_inc, (float)i), angle_min);
      const float px = (float)__dmul_rn((double)rg, cos((double)angle));
      const float py = (float)__dmul_rn((double)rg, sin((double)angle));
      const float rx = __fadd_rn(__fmul_rn(lc, px), __fmul_rn(-ls, py));
      const float ry = __fadd_rn(__fmul_rn(ls, px), __fmul_rn(lc, py));
      const float2 v = make_float2(__fadd_rn(lx, rx), __fadd_rn(ly, ry));
      if (!(fabsf(v.x) <= DPGICP_MAX_ABS_COORD) || !(fabsf(v.y) <= DPGICP_MAX_ABS_COORD)) atomicExch(bad, 1);
      row[n + __popc(bal & ((1u << lane) - 1u))] = v;
    }
    n += __popc(bal);
  }
  for (int k = n + lane; k < pitch; k += 32) row[k] = make_float2(0.f, 0.f);
  if (lane == 0) count[warp] = n;
}

/* ------------------------------------------------------------------------------------------------
 * Candidate-pair enumeration (reoptimize's distance gate, dpg_slam.cc:79-107): count pass then
 * fill pass; one thread per source node i, output in the reference's (i, j) loop order.
 * ---------------------------------------------------------------------------------------------- */
__device__ __forceinline__ bool gate_pair(const float2 *xy, const int32_t *pass, int i, int j, float r_same,
                                          float r_other) {
  const float dx = __fsub_rn(xy[j].x, xy[i].x), dy = __fsub_rn(xy[j].y, xy[i].y);
  const float dist = __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
  return dist <= ((pass[j] == pass[i]) ? r_same : r_other);
}

__global__ void enumerate_count_kernel(const float2 *xy, const int32_t *pass, int n, float r_same, float r_other,
                                       unsigned long long *cnt) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned long long c = 0;
  if (i >= 1) {
    c = 1;
    for (int j = 0; j < i - 1; ++j) c += gate_pair(xy, pass, i, j, r_same, r_other) ? 1 : 0;
  }
  cnt[i] = c;
}

__global__ void enumerate_fill_kernel(const float2 *xy, const int32_t *pass, int n, float r_same, float r_other,
                                      const unsigned long long *start, int32_t *src, int32_t *tgt) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || i < 1) return;
  unsigned long long o = start[i];
  src[o] = i; tgt[o] = i - 1; ++o;
  for (int j = 0; j < i - 1; ++j)
    if (gate_pair(xy, pass, i, j, r_same, r_other)) { src[o] = i; tgt[o] = j; ++o; }
}

/* ------------------------------------------------------------------------------------------------
 * FP32-pipe probe: the roofline denominator of the distance loop, measured on the device in use.
 * 8 independent dependent-chains per thread of separately rounded FMUL + FADD (the instruction mix
 * the bit-exact distance loop is allowed to use; FMA = false) or of FFMA (FMA = true, for context).
 * ---------------------------------------------------------------------------------------------- */
template <bool FMA>
__global__ void __launch_bounds__(256) fp32_probe_kernel(float *out, int iters, float a, float b) {
  float x[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) x[k] = (float)(threadIdx.x + k) * 1e-3f;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (FMA) x[k] = __fmaf_rn(x[k], a, b);
      else x[k] = __fadd_rn(__fmul_rn(x[k], a), b);
    }
  }
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) s += x[k];
  if (s == 123.456f) out[0] = s;       /* keeps the chains alive; practically never true */
}

}  // namespace dpg
