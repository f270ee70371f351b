/*
 * dpgsynth.c — synthetic "Hokuyo-like" 2D laser scans for the batch runner, the tests and the
 * benchmark (SURVEY.md §8d scanner model).  Host-only C; it produces the *inputs* of the hot path
 * (raw ranges per beam, as the reference receives them in DpgSLAM::ObserveLaser,
 * src/dpg_slam/dpg_slam.cc:122-140) and is not part of the timed region.
 *
 * A world is a list of wall segments (x0, y0, x1, y1).  A scan is cast from the laser origin
 * (robot pose composed with the laser offset, parameters.h:319-339 default (0.2, 0, 0)) with
 * beam i at angle_min + i*angle_inc (float arithmetic as createNode, dpg_slam.cc:497-504).
 * Beams with no hit below range_max return range_max, which the path drops
 * (dpg_measurement.h:43-45).  Noise is a counter-based splitmix64 + Box-Muller Gaussian keyed by
 * (seed, scan, beam), so results do not depend on thread count or order.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static inline uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

static inline double u01(uint64_t h) { return ((double)(h >> 11) + 0.5) * (1.0 / 9007199254740992.0); }

static double gauss(uint64_t seed, uint64_t a, uint64_t b) {
  uint64_t k = splitmix64(seed ^ splitmix64(a * 0x100000001B3ull + b));
  double u1 = u01(k), u2 = u01(splitmix64(k));
  return sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
}

/* uniform in [0,1) from a hashed key; exported for the Python side's world builders */
double dpgsynth_uniform(uint64_t seed, uint64_t a, uint64_t b) {
  return u01(splitmix64(seed ^ splitmix64(a * 0x100000001B3ull + b)));
}

static int push_seg(float *segs, int n, int cap, double x0, double y0, double x1, double y1) {
  if (n < cap) {
    segs[4 * n + 0] = (float)x0; segs[4 * n + 1] = (float)y0;
    segs[4 * n + 2] = (float)x1; segs[4 * n + 3] = (float)y1;
  }
  return n + 1;
}

static int push_box(float *segs, int n, int cap, double x0, double y0, double x1, double y1) {
  n = push_seg(segs, n, cap, x0, y0, x1, y0);
  n = push_seg(segs, n, cap, x1, y0, x1, y1);
  n = push_seg(segs, n, cap, x1, y1, x0, y1);
  n = push_seg(segs, n, cap, x0, y1, x0, y0);
  return n;
}

/* rectangular room [-w/2, w/2] x [-h/2, h/2] plus one off-centre pillar; returns segment count */
int dpgsynth_world_room(double w, double h, float *segs, int cap) {
  int n = 0;
  n = push_box(segs, n, cap, -0.5 * w, -0.5 * h, 0.5 * w, 0.5 * h);
  n = push_box(segs, n, cap, 0.27 * w, 0.12 * h, 0.27 * w + 0.4, 0.12 * h + 0.4);
  return n;
}

/* Corridor of the given width along +x from x_from to x_to, with door recesses (depth ~0.5 m,
 * width 0.8-1.4 m) every `period` metres alternating sides; recess geometry is hashed from its
 * index so no two stretches look alike.  Closed at both ends. */
int dpgsynth_world_corridor(double x_from, double x_to, double width, double period, uint64_t seed,
                            float *segs, int cap) {
  int n = 0;
  const double hw = 0.5 * width;
  long k0 = (long)floor(x_from / period), k1 = (long)ceil(x_to / period);
  for (int side = 0; side < 2; ++side) {
    const double sgn = side ? -1.0 : 1.0;
    double x = x_from;
    for (long k = k0; k <= k1; ++k) {
      if (((k & 1) != 0) != (side != 0)) continue;   /* alternate sides */
      double cx = (double)k * period + period * (0.25 + 0.5 * dpgsynth_uniform(seed, (uint64_t)(k + (1l << 40)), 1));
      double rw = 0.8 + 0.6 * dpgsynth_uniform(seed, (uint64_t)(k + (1l << 40)), 2);
      double rd = 0.35 + 0.3 * dpgsynth_uniform(seed, (uint64_t)(k + (1l << 40)), 3);
      double a = cx - 0.5 * rw, b = cx + 0.5 * rw;
      if (a <= x || b >= x_to) continue;
      n = push_seg(segs, n, cap, x, sgn * hw, a, sgn * hw);
      n = push_seg(segs, n, cap, a, sgn * hw, a, sgn * (hw + rd));
      n = push_seg(segs, n, cap, a, sgn * (hw + rd), b, sgn * (hw + rd));
      n = push_seg(segs, n, cap, b, sgn * (hw + rd), b, sgn * hw);
      x = b;
    }
    n = push_seg(segs, n, cap, x, sgn * hw, x_to, sgn * hw);
  }
  n = push_seg(segs, n, cap, x_from, -hw, x_from, hw);
  n = push_seg(segs, n, cap, x_to, -hw, x_to, hw);
  return n;
}

/* Office-like world: outer square [0,size]^2 plus n_boxes axis-aligned boxes (0.5-4 m) scattered
 * by hash.  `variant` > 0 moves `moved_fraction` of the boxes (dynamic-environment sessions).   */
int dpgsynth_world_office(double size, int n_boxes, uint64_t seed, int variant, double moved_fraction,
                          float *segs, int cap) {
  int n = 0;
  n = push_box(segs, n, cap, 0.0, 0.0, size, size);
  for (int b = 0; b < n_boxes; ++b) {
    uint64_t key = (uint64_t)b;
    uint64_t s = seed;
    if (variant > 0 && dpgsynth_uniform(seed, key, 100 + (uint64_t)variant) < moved_fraction)
      s = seed + 7919ull * (uint64_t)variant;          /* this box is somewhere else in this session */
    double w = 0.5 + 3.5 * dpgsynth_uniform(s, key, 1);
    double h = 0.5 + 3.5 * dpgsynth_uniform(s, key, 2);
    double x = 1.0 + (size - w - 2.0) * dpgsynth_uniform(s, key, 3);
    double y = 1.0 + (size - h - 2.0) * dpgsynth_uniform(s, key, 4);
    n = push_box(segs, n, cap, x, y, x + w, y + h);
  }
  return n;
}

/* returns 1 if point (x,y) is at least `margin` away from every segment */
int dpgsynth_is_free(const float *segs, int nseg, double x, double y, double margin) {
  for (int i = 0; i < nseg; ++i) {
    double ax = segs[4 * i], ay = segs[4 * i + 1], bx = segs[4 * i + 2], by = segs[4 * i + 3];
    double vx = bx - ax, vy = by - ay, wx = x - ax, wy = y - ay;
    double l2 = vx * vx + vy * vy;
    double t = l2 > 0 ? (wx * vx + wy * vy) / l2 : 0.0;
    if (t < 0) t = 0;
    if (t > 1) t = 1;
    double dx = wx - t * vx, dy = wy - t * vy;
    if (dx * dx + dy * dy < margin * margin) return 0;
  }
  return 1;
}

/* returns 1 if (x,y) is inside any of the boxes of an office world (boxes are 4 consecutive
 * segments after the 4 outer ones) */
int dpgsynth_inside_box(const float *segs, int nseg, double x, double y) {
  for (int i = 4; i + 3 < nseg; i += 4) {
    double x0 = segs[4 * i], y0 = segs[4 * i + 1], x1 = segs[4 * i + 2], y1 = segs[4 * (i + 1) + 3];
    double lox = x0 < x1 ? x0 : x1, hix = x0 < x1 ? x1 : x0;
    double loy = y0 < y1 ? y0 : y1, hiy = y0 < y1 ? y1 : y0;
    if (x >= lox && x <= hix && y >= loy && y <= hiy) return 1;
  }
  return 0;
}

/*
 * Cast n_scans scans.  poses = (x, y, theta) of base_link per scan in the world frame (double).
 * ranges_out is n_scans * n_beams floats.  threads <= 0 -> OpenMP default.
 */
void dpgsynth_cast_scans(const float *segs, int nseg, const double *poses, int n_scans, int n_beams,
                         float angle_min, float angle_max, float range_min, float range_max,
                         double noise_sigma, uint64_t seed, double laser_x, double laser_y,
                         int threads, float *ranges_out) {
  const float angle_inc = (float)(((double)(float)(angle_max - angle_min)) / ((double)n_beams - 1.0));
#ifdef _OPENMP
  if (threads <= 0) threads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 8) num_threads(threads)
#else
  (void)threads;
#endif
  for (int s = 0; s < n_scans; ++s) {
    const double px = poses[3 * s], py = poses[3 * s + 1], th = poses[3 * s + 2];
    const double ox = px + cos(th) * laser_x - sin(th) * laser_y;
    const double oy = py + sin(th) * laser_x + cos(th) * laser_y;
    /* cull to segments that can be hit */
    int *near = (int *)malloc(sizeof(int) * (size_t)(nseg > 0 ? nseg : 1));
    int nn = 0;
    for (int i = 0; i < nseg; ++i) {
      double ax = segs[4 * i], ay = segs[4 * i + 1], bx = segs[4 * i + 2], by = segs[4 * i + 3];
      double lox = fmin(ax, bx) - range_max, hix = fmax(ax, bx) + range_max;
      double loy = fmin(ay, by) - range_max, hiy = fmax(ay, by) + range_max;
      if (ox >= lox && ox <= hix && oy >= loy && oy <= hiy) near[nn++] = i;
    }
    for (int b = 0; b < n_beams; ++b) {
      const float ang = angle_inc * (float)b + angle_min;
      const double dir = th + (double)ang;
      const double dx = cos(dir), dy = sin(dir);
      double best = INFINITY;
      for (int q = 0; q < nn; ++q) {
        const int i = near[q];
        const double ax = segs[4 * i], ay = segs[4 * i + 1];
        const double ex = (double)segs[4 * i + 2] - ax, ey = (double)segs[4 * i + 3] - ay;
        const double den = dx * ey - dy * ex;
        if (fabs(den) < 1e-12) continue;
        const double wx = ax - ox, wy = ay - oy;
        const double t = (wx * ey - wy * ex) / den;    /* along the ray */
        const double u = (wx * dy - wy * dx) / den;    /* along the segment */
        if (t > 0.0 && u >= 0.0 && u <= 1.0 && t < best) best = t;
      }
      float r;
      if (!(best < (double)range_max)) {
        r = range_max;
      } else {
        double v = best + noise_sigma * gauss(seed, (uint64_t)s, (uint64_t)b);
        if (v < (double)range_min) v = (double)range_min;
        r = (float)v;
        if (r >= range_max) r = range_max;
      }
      ranges_out[(size_t)s * (size_t)n_beams + (size_t)b] = r;
    }
    free(near);
  }
}
