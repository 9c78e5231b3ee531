/*
 * oracle/roialign_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C, CPU restatement of MobulaOP's ROIAlign forward pool and backward
 * scatter, used only as the checker for the CUDA path (tests/, smoke(), and
 * bench.py's cpu_baseline / --impl reference legs).  Nothing under
 * mobulaop_b200/ may import, link or execute it.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this file bit-for-bit
 * against (a) golden vectors produced by the unmodified reference running in
 * the build container (tests/golden/, generator tests/golden/gen_golden.py),
 * (b) the reference's own hand-checkable known answer and
 * (c) oracle/_ref, the reference's C++ sources compiled where they lie.
 *
 * What it follows (all paths relative to /root/reference):
 *   forward  : opzoo/ROIAlign/ROIAlign.cpp:9-76
 *   backward : opzoo/ROIAlign/ROIAlign.cpp:78-161
 *   tap math : opzoo/ROIAlign/bilinear.h:10-59 (value), :61-112 (gradient)
 *   order    : mobula/cpp/include/context/openmp_ctx.h:50-67 (single-thread
 *              parfor = ascending flat index) and common.h:29-39 (per-thread
 *              contiguous ranges for the multi-thread variants)
 *
 * The arithmetic below is organised differently from the reference (the
 * bilinear tap is decoded per axis, because row and column decoding are
 * independent) but performs the same fp32 operations in the same order, so
 * results are bit-identical on a target without FMA contraction.  Build with
 * -ffp-contract=off (see oracle/Makefile).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define ORACLE_API __attribute__((visibility("default")))

/* One axis of a bilinear tap (bilinear.h:20-48 restated per axis).
 * `pos` is the sample coordinate along the axis, `extent` the plane size.
 * The "outside" rejection (bilinear.h:15-18) involves both axes and is done by
 * the caller through axis_outside(). */
typedef struct {
  int lo, hi;   /* the two pixel indices                                   */
  float wl, wh; /* weight of `hi` (= frac) and of `lo` (= 1 - frac)         */
} axis_tap;

static inline int axis_outside(float pos, int extent) {
  /* bilinear.h:15: `y < -1.0 || y > height` (int promoted to the float type) */
  return pos < -1.0f || pos > (float)extent;
}

static inline axis_tap axis_decode(float pos, int extent) {
  axis_tap t;
  if (pos <= 0.0f) pos = 0.0f;            /* bilinear.h:20-25 */
  t.lo = (int)pos;                         /* bilinear.h:27-28 */
  if (t.lo >= extent - 1) {                /* bilinear.h:32-44 */
    t.hi = t.lo = extent - 1;
    pos = (float)t.lo;
  } else {
    t.hi = t.lo + 1;
  }
  t.wl = pos - (float)t.lo;                /* ly / lx, bilinear.h:46-47 */
  /* bilinear.h:48 computes `1. - ly` in double and rounds to float; for one
   * subtraction of fp32 operands double rounding is innocuous (53 >= 2*24+2),
   * so the fp32 subtraction gives the same bits. */
  t.wh = 1.0f - t.wl;
  return t;
}

/* Per-ROI quantities that do not depend on (c, ph, pw): ROIAlign.cpp:24-54. */
typedef struct {
  int batch;
  float start_w, start_h;
  float bin_w, bin_h;
  int grid_h, grid_w;
  float count;
} roi_geom;

static inline roi_geom roi_decode(const float* roi, float spatial_scale,
                                  int pooled_h, int pooled_w,
                                  int sampling_ratio) {
  roi_geom g;
  g.batch = (int)roi[0];                                   /* :25 */
  g.start_w = roi[1] * spatial_scale;                      /* :28 */
  g.start_h = roi[2] * spatial_scale;                      /* :29 */
  float end_w = roi[3] * spatial_scale;                    /* :30 */
  float end_h = roi[4] * spatial_scale;                    /* :31 */
  float roi_w = fmaxf(end_w - g.start_w, 1.0f);            /* :38 */
  float roi_h = fmaxf(end_h - g.start_h, 1.0f);            /* :39 */
  g.bin_h = roi_h / (float)pooled_h;                       /* :40 */
  g.bin_w = roi_w / (float)pooled_w;                       /* :41 */
  g.grid_h = sampling_ratio > 0 ? sampling_ratio
                                : (int)ceilf(roi_h / (float)pooled_h); /* :47-49 */
  g.grid_w = sampling_ratio > 0 ? sampling_ratio
                                : (int)ceilf(roi_w / (float)pooled_w); /* :50-51 */
  g.count = (float)(g.grid_h * g.grid_w);                  /* :54 */
  return g;
}

/* Sample coordinate along one axis: ROIAlign.cpp:59-61 / :63-65.
 * (start + p*bin) + (((i + .5f) * bin) / grid), left to right. */
static inline float sample_pos(float start, int p, float bin, int i, int grid) {
  return start + (float)p * bin + ((float)i + 0.5f) * bin / (float)grid;
}

static inline float pool_one(const float* plane, int height, int width,
                             const roi_geom* g, int ph, int pw) {
  float acc = 0.0f;
  for (int iy = 0; iy < g->grid_h; ++iy) {
    const float y = sample_pos(g->start_h, ph, g->bin_h, iy, g->grid_h);
    for (int ix = 0; ix < g->grid_w; ++ix) {
      const float x = sample_pos(g->start_w, pw, g->bin_w, ix, g->grid_w);
      float val = 0.0f;
      if (!(axis_outside(y, height) || axis_outside(x, width))) {
        const axis_tap ty = axis_decode(y, height);
        const axis_tap tx = axis_decode(x, width);
        const float v1 = plane[(int64_t)ty.lo * width + tx.lo];
        const float v2 = plane[(int64_t)ty.lo * width + tx.hi];
        const float v3 = plane[(int64_t)ty.hi * width + tx.lo];
        const float v4 = plane[(int64_t)ty.hi * width + tx.hi];
        const float w1 = ty.wh * tx.wh, w2 = ty.wh * tx.wl;
        const float w3 = ty.wl * tx.wh, w4 = ty.wl * tx.wl;
        val = w1 * v1 + w2 * v2 + w3 * v3 + w4 * v4;       /* bilinear.h:56 */
      }
      acc += val;                                          /* :69 */
    }
  }
  return acc / g->count;                                   /* :72 */
}

static void forward_range(int64_t begin, int64_t end, const float* data,
                          float spatial_scale, int channels, int height,
                          int width, int pooled_h, int pooled_w,
                          int sampling_ratio, const float* rois, float* out) {
  const int64_t bins = (int64_t)pooled_h * pooled_w;
  for (int64_t index = begin; index < end; ++index) {
    const int pw = (int)(index % pooled_w);
    const int ph = (int)((index / pooled_w) % pooled_h);
    const int c = (int)((index / bins) % channels);
    const int64_t n = index / bins / channels;
    const roi_geom g =
        roi_decode(rois + n * 5, spatial_scale, pooled_h, pooled_w, sampling_ratio);
    const float* plane =
        data + ((int64_t)g.batch * channels + c) * height * width;
    out[index] = pool_one(plane, height, width, &g, ph, pw);
  }
}

/* mode 0: plain `+=` (single thread, canonical order)
 * mode 1: `#pragma omp atomic` per tap (cpu_ctx.h:37-41) */
static void backward_range(int64_t begin, int64_t end, const float* dy,
                           float spatial_scale, int channels, int height,
                           int width, int pooled_h, int pooled_w,
                           int sampling_ratio, float* dx, const float* rois,
                           int atomic_mode) {
  const int64_t bins = (int64_t)pooled_h * pooled_w;
  for (int64_t index = begin; index < end; ++index) {
    const int pw = (int)(index % pooled_w);
    const int ph = (int)((index / pooled_w) % pooled_h);
    const int c = (int)((index / bins) % channels);
    const int64_t n = index / bins / channels;
    const roi_geom g =
        roi_decode(rois + n * 5, spatial_scale, pooled_h, pooled_w, sampling_ratio);
    float* plane = dx + ((int64_t)g.batch * channels + c) * height * width;
    const float dtop = dy[index];                          /* :113-115 */
    for (int iy = 0; iy < g.grid_h; ++iy) {
      const float y = sample_pos(g.start_h, ph, g.bin_h, iy, g.grid_h);
      for (int ix = 0; ix < g.grid_w; ++ix) {
        const float x = sample_pos(g.start_w, pw, g.bin_w, ix, g.grid_w);
        /* bilinear.h:67-72: outside -> indices -1, nothing is added (:148) */
        if (axis_outside(y, height) || axis_outside(x, width)) continue;
        const axis_tap ty = axis_decode(y, height);
        const axis_tap tx = axis_decode(x, width);
        const float w1 = ty.wh * tx.wh, w2 = ty.wh * tx.wl;
        const float w3 = ty.wl * tx.wh, w4 = ty.wl * tx.wl;
        const float g1 = dtop * w1 / g.count;              /* :143-146 */
        const float g2 = dtop * w2 / g.count;
        const float g3 = dtop * w3 / g.count;
        const float g4 = dtop * w4 / g.count;
        float* p1 = plane + (int64_t)ty.lo * width + tx.lo;
        float* p2 = plane + (int64_t)ty.lo * width + tx.hi;
        float* p3 = plane + (int64_t)ty.hi * width + tx.lo;
        float* p4 = plane + (int64_t)ty.hi * width + tx.hi;
        if (atomic_mode) {
#pragma omp atomic update
          *p1 += g1;
#pragma omp atomic update
          *p2 += g2;
#pragma omp atomic update
          *p3 += g3;
#pragma omp atomic update
          *p4 += g4;
        } else {
          *p1 += g1;                                       /* :149-156 */
          *p2 += g2;
          *p3 += g3;
          *p4 += g4;
        }
      }
    }
  }
}

/* parfor block partition (common.h:29-39) */
static inline void block_range(int64_t n, int nthr, int tid, int64_t* b, int64_t* e) {
  const int64_t avg = n / nthr, rest = n % nthr;
  *b = avg * tid + (tid < rest ? tid : rest);
  *e = *b + avg + (tid < rest);
}

/* ---- exported entry points (same argument order as the reference kernels) ---- */

ORACLE_API void oracle_roi_align_forward_f32(
    int64_t nthreads, const float* bottom_data, float spatial_scale, int channels,
    int height, int width, int pooled_height, int pooled_width,
    int sampling_ratio, const float* bottom_rois, float* top_data) {
  forward_range(0, nthreads, bottom_data, spatial_scale, channels, height, width,
                pooled_height, pooled_width, sampling_ratio, bottom_rois, top_data);
}

/* bottom_diff must be pre-zeroed by the caller (ROIAlign.py:33-34). */
ORACLE_API void oracle_roi_align_backward_f32(
    int64_t nthreads, const float* top_diff, float spatial_scale, int channels,
    int height, int width, int pooled_height, int pooled_width,
    int sampling_ratio, float* bottom_diff, const float* bottom_rois) {
  backward_range(0, nthreads, top_diff, spatial_scale, channels, height, width,
                 pooled_height, pooled_width, sampling_ratio, bottom_diff,
                 bottom_rois, 0);
}

/* Multi-thread variants = what the reference's OpenMP build does
 * (openmp_ctx.h:16-46): every host thread walks a contiguous index range;
 * backward adds with `omp atomic`, so its rounding order is not reproducible. */
ORACLE_API int oracle_roi_align_forward_f32_mt(
    int64_t nthreads, const float* bottom_data, float spatial_scale, int channels,
    int height, int width, int pooled_height, int pooled_width,
    int sampling_ratio, const float* bottom_rois, float* top_data, int host_threads) {
  int used = 1;
#ifdef _OPENMP
  if (host_threads <= 0) host_threads = omp_get_max_threads();
  if ((int64_t)host_threads > nthreads) host_threads = nthreads > 0 ? (int)nthreads : 1;
  used = host_threads;
#pragma omp parallel num_threads(host_threads)
  {
    int64_t b, e;
    block_range(nthreads, omp_get_num_threads(), omp_get_thread_num(), &b, &e);
    forward_range(b, e, bottom_data, spatial_scale, channels, height, width,
                  pooled_height, pooled_width, sampling_ratio, bottom_rois, top_data);
  }
#else
  (void)host_threads;
  forward_range(0, nthreads, bottom_data, spatial_scale, channels, height, width,
                pooled_height, pooled_width, sampling_ratio, bottom_rois, top_data);
#endif
  return used;
}

ORACLE_API int oracle_roi_align_backward_f32_mt(
    int64_t nthreads, const float* top_diff, float spatial_scale, int channels,
    int height, int width, int pooled_height, int pooled_width,
    int sampling_ratio, float* bottom_diff, const float* bottom_rois,
    int host_threads) {
  int used = 1;
#ifdef _OPENMP
  if (host_threads <= 0) host_threads = omp_get_max_threads();
  if ((int64_t)host_threads > nthreads) host_threads = nthreads > 0 ? (int)nthreads : 1;
  used = host_threads;
#pragma omp parallel num_threads(host_threads)
  {
    int64_t b, e;
    block_range(nthreads, omp_get_num_threads(), omp_get_thread_num(), &b, &e);
    backward_range(b, e, top_diff, spatial_scale, channels, height, width,
                   pooled_height, pooled_width, sampling_ratio, bottom_diff,
                   bottom_rois, 1);
  }
#else
  (void)host_threads;
  backward_range(0, nthreads, top_diff, spatial_scale, channels, height, width,
                 pooled_height, pooled_width, sampling_ratio, bottom_diff,
                 bottom_rois, 0);
#endif
  return used;
}

/*
 * ROI bookkeeping export, for the "index bookkeeping must be bit-exact" gate.
 * For every ROI r writes
 *   geom_i[r*4 + {0,1,2,3}] = batch index, grid_h, grid_w, 0
 *   geom_f[r*5 + {0..4}]    = start_w, start_h, bin_w, bin_h, count
 * and, when max_grid > 0, for every (r, p, i) with p < pooled and i < grid the
 * decoded axis taps along y (axis 0) and x (axis 1):
 *   tap_i[(((r*2 + axis)*pooled_max + p)*max_grid + i)*2 + {0,1}] = lo, hi
 *         (lo = hi = -1 when the coordinate is outside the plane)
 *   tap_f[... same index ... *2 + {0,1}]                         = pos, frac
 * Entries beyond a ROI's own grid are left untouched.
 */
ORACLE_API void oracle_roi_align_bookkeeping_f32(
    int num_rois, float spatial_scale, int height, int width, int pooled_height,
    int pooled_width, int sampling_ratio, const float* bottom_rois,
    int32_t* geom_i, float* geom_f, int max_grid, int32_t* tap_i, float* tap_f) {
  const int pooled_max = pooled_height > pooled_width ? pooled_height : pooled_width;
  for (int r = 0; r < num_rois; ++r) {
    const roi_geom g = roi_decode(bottom_rois + (int64_t)r * 5, spatial_scale,
                                  pooled_height, pooled_width, sampling_ratio);
    geom_i[r * 4 + 0] = g.batch;
    geom_i[r * 4 + 1] = g.grid_h;
    geom_i[r * 4 + 2] = g.grid_w;
    geom_i[r * 4 + 3] = 0;
    geom_f[r * 5 + 0] = g.start_w;
    geom_f[r * 5 + 1] = g.start_h;
    geom_f[r * 5 + 2] = g.bin_w;
    geom_f[r * 5 + 3] = g.bin_h;
    geom_f[r * 5 + 4] = g.count;
    if (max_grid <= 0) continue;
    for (int axis = 0; axis < 2; ++axis) {
      const int pooled = axis == 0 ? pooled_height : pooled_width;
      const int grid = axis == 0 ? g.grid_h : g.grid_w;
      const int extent = axis == 0 ? height : width;
      const float start = axis == 0 ? g.start_h : g.start_w;
      const float bin = axis == 0 ? g.bin_h : g.bin_w;
      for (int p = 0; p < pooled; ++p) {
        for (int i = 0; i < grid && i < max_grid; ++i) {
          const int64_t e =
              ((((int64_t)r * 2 + axis) * pooled_max + p) * max_grid + i) * 2;
          const float pos = sample_pos(start, p, bin, i, grid);
          tap_f[e + 0] = pos;
          if (axis_outside(pos, extent)) {
            tap_i[e + 0] = tap_i[e + 1] = -1;
            tap_f[e + 1] = 0.0f;
          } else {
            const axis_tap t = axis_decode(pos, extent);
            tap_i[e + 0] = t.lo;
            tap_i[e + 1] = t.hi;
            tap_f[e + 1] = t.wl;
          }
        }
      }
    }
  }
}

ORACLE_API int oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
