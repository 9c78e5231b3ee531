// "resident" kernels: the fast path for a fixed sampling ratio (ROIAlign.cpp:47-51 with
// sampling_ratio > 0) when one padded (image, channel) plane fits the shared memory of an SM.
//
// HBM sees the algorithmic traffic only: a CTA brings a plane global->shared once with the
// bulk copy engine (one cp.async.bulk per row, completion on an mbarrier), every bilinear tap
// of every ROI of that image is then a shared-memory access, and the plane never goes back
// (forward) or goes back exactly once (backward, zero-fill fused).  What bounds these kernels
// is the shared-memory pipe and the issue slots around it: a bilinear tap is a 4-byte gather
// with ~2.3 bank conflicts per warp access whatever the lane mapping, so everything else is
// kept off the critical path:
//
//   * ROI decoding is channel independent and is done once per call by a table kernel, in
//     exactly the reference's rounding (roialign_common.cuh); entries hold ready-to-use byte
//     offsets and weights and are fetched one ROI ahead of their use.
//   * Padded plane: rows 0..H-1 / columns 0..W-1 hold the data, row H / column W repeat the
//     last row / column, rows H+1, H+2 and columns W+1.. are zero.  The high tap of a sample
//     is then ALWAYS one row below / one column right of the low tap (the reference's clamp
//     y_high = y_low = H-1, bilinear.h:32-44, reads the repeated row with weight 0), and a
//     sample outside the plane (bilinear.h:15-18) needs no branch: its entry points into the
//     zero rows / columns, so it contributes exactly +0.
//   * The work -- (plane, ROI) pairs -- is split evenly over the CTAs, not plane by plane:
//     512 planes over 148 SMs would otherwise leave a 14% tail.
//
// forward : products and sums in the reference's order (bilinear.h:50-56, ROIAlign.cpp:56-74)
//           -> bit-identical output.
#include <cstdlib>
#include "roialign_kernels.h"
#include "roialign_common.cuh"
#include "ptx_sm100.cuh"

namespace mobula_b200 {

namespace {

constexpr int kFwdWarps = 16;
constexpr int kFwdThreads = kFwdWarps * 32;
constexpr int kMaxEntries = 64;      // generic forward: (PH + PW) * G entries, two per lane

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// One sample coordinate along one axis: byte offset of the low tap (rows: row * pitch * 4,
// columns: col * 4) and the fraction l = weight of the high tap.  Rejected samples point at
// the first zero row / column of the padded plane.
struct AxisSample {
  unsigned lo_off;
  float l;
};

__device__ __forceinline__ AxisSample make_sample(float start, int p, float bin, int i, int grid, int extent,
                                                  int stride_bytes) {
  const AxisTap t = axis_decode(sample_pos(start, p, bin, i, grid), extent);
  AxisSample s;
  s.lo_off = (unsigned)((t.valid ? t.lo : extent + 1) * stride_bytes);
  s.l = t.valid ? t.wl : 0.0f;
  return s;
}

// ---- tables ------------------------------------------------------------------------------
// generic forward: thread (j, k) writes entry k of the ROI at grouped position j, y entries
// first; x = lo_off, y = hi_off (always lo_off + one row / column), z = l, w = 1 - l
__global__ void __launch_bounds__(256)
resident_fwd_table_kernel(const float* __restrict__ rois, const int* __restrict__ order,
                          const int* __restrict__ img_start, int num_images, int height, int width,
                          int pitch, int pooled_h, int pooled_w, float scale, int grid,
                          uint4* __restrict__ entries) {
  ptx::chain_enter();
  const int per_roi = (pooled_h + pooled_w) * grid;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int valid = __ldg(img_start + num_images);
  if (t >= (long long)valid * per_roi) return;
  const int j = (int)(t / per_roi), k = (int)(t - (long long)j * per_roi);
  const int roi = __ldg(order + j);
  const RoiGeom g = roi_decode(rois + (size_t)roi * 5, scale, pooled_h, pooled_w, grid);
  const int ny = pooled_h * grid;
  AxisSample s;
  int stride;
  if (k < ny) {
    const int p = k / grid;
    stride = pitch * 4;
    s = make_sample(g.start_h, p, g.bin_h, k - p * grid, grid, height, stride);
  } else {
    const int kk = k - ny, p = kk / grid;
    stride = 4;
    s = make_sample(g.start_w, p, g.bin_w, kk - p * grid, grid, width, stride);
  }
  entries[t] = make_uint4(s.lo_off, s.lo_off + (unsigned)stride, __float_as_uint(s.l),
                          __float_as_uint(__fsub_rn(1.0f, s.l)));
}

// sampling_ratio == 2 forward: per ROI `ypairs` 16-byte records {lo_off(iy=0), l, lo_off(iy=1), l}
// -- one per bin row, padded with zero-row records -- followed by `xslots` 8-byte records
// {lo_off, l}, one per sample column, padded with zero-column records.
__global__ void __launch_bounds__(256)
resident_fwd2_table_kernel(const float* __restrict__ rois, const int* __restrict__ order,
                           const int* __restrict__ img_start, int num_images, int height, int width,
                           int pitch, int unit, int pooled_h, int pooled_w, float scale, int ypairs,
                           int xslots, uint2* __restrict__ entries, int dense) {
  ptx::chain_enter();
  // `unit` = bytes per pixel of the shared-memory plane: 4, or 8 for a channel pair
  const int per_roi = 2 * ypairs + xslots;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int valid = __ldg(img_start + num_images);
  if (t >= (long long)valid * per_roi) return;
  const int j = (int)(t / per_roi), k = (int)(t - (long long)j * per_roi);
  const int roi = __ldg(order + j);
  const RoiGeom g = roi_decode(rois + (size_t)roi * 5, scale, pooled_h, pooled_w, 2);
  AxisSample s;
  if (k < 2 * ypairs) {
    s.lo_off = (unsigned)((height + 1) * pitch * unit);
    s.l = 0.0f;
    if (k < 2 * pooled_h) {
      s = make_sample(g.start_h, k >> 1, g.bin_h, k & 1, 2, height, pitch * unit);
      // Row reuse (the offsets are multiples of 16, the code sits in their low two bits): how the
      // taps of this sample row relate to those of the previous sample row of the ROI --
      //   2: the same two plane rows (nothing to load), 1: its low row is the previous high row
      //   (one row to load), 0: both rows are new.  Sample rows closer than a pixel are the rule
      //   for small ROIs (bin_h / 2 < 1 below 14 feature-map rows at 7 x 7).
      if (k > 0) {
        const AxisSample prev = make_sample(g.start_h, (k - 1) >> 1, g.bin_h, (k - 1) & 1, 2, height, pitch * unit);
        if (s.lo_off == prev.lo_off) s.lo_off |= 2u;
        else if (s.lo_off == prev.lo_off + (unsigned)(pitch * unit)) s.lo_off |= 1u;
      }
    }
  } else {
    const int kk = k - 2 * ypairs;
    s.lo_off = (unsigned)((width + 1) * unit);
    s.l = 0.0f;
    if (dense) {
      // dense plane (pitch == width: no repeated column, no zero columns): the record says so itself --
      // bit 0: the sample is rejected (the lane loads nothing, its taps are 0), bit 1: the column is
      // clamped, its high tap is the low tap (bilinear.h:27-31)
      s.lo_off = 1u;
      if (kk < 2 * pooled_w) {
        const AxisTap t = axis_decode(sample_pos(g.start_w, kk >> 1, g.bin_w, kk & 1, 2), width);
        if (t.valid) {
          s.lo_off = (unsigned)(t.lo * unit) | (t.hi == t.lo ? 2u : 0u);
          s.l = t.wl;
        }
      }
    } else if (kk < 2 * pooled_w) {
      s = make_sample(g.start_w, kk >> 1, g.bin_w, kk & 1, 2, width, unit);
    }
  }
  entries[t] = make_uint2(s.lo_off, __float_as_uint(s.l));
}

// ---- plane movement ----------------------------------------------------------------------
// zero padding the loads never touch: columns [W + 1, pitch) of rows [0, H] and rows H+1, H+2
__device__ __forceinline__ void zero_padding(float* plane, int height, int width, int pitch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
  for (int r = warp; r <= height; r += warps)
    for (int x = width + 1 + lane; x < pitch; x += 32) plane[r * pitch + x] = 0.0f;
  for (int x = threadIdx.x; x < 2 * pitch; x += blockDim.x) plane[(height + 1) * pitch + x] = 0.0f;
}

// rows of `width` floats, dense in global memory, `pitch` floats apart in shared memory;
// then the repeated column W and the repeated row H.  Ends with the plane visible to all.
// `vec16`: every row starts 16-byte aligned in global memory (W % 4 == 0, aligned base).
__device__ __forceinline__ void load_padded_plane(float* plane, const float* __restrict__ src, int height,
                                                  int width, int pitch, int vec16) {
  if (vec16) {
    // 16-byte asynchronous copies, every thread a strided share: thousands in flight per SM
    // (the per-row cp.async.bulk variant reached 3 TB/s over the chip, this one is HBM bound)
    const int w4 = width >> 2, total = height * w4;
    int r = threadIdx.x / w4, c4 = threadIdx.x - r * w4;
    const int dr = blockDim.x / w4, dc = blockDim.x - dr * w4;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
      ptx::cp_async16(plane + (size_t)r * pitch + 4 * c4, src + (size_t)r * width + 4 * c4);
      r += dr;
      c4 += dc;
      if (c4 >= w4) {
        c4 -= w4;
        ++r;
      }
    }
    ptx::cp_async_wait_all();
    __syncthreads();
  } else {
    for (int i = threadIdx.x; i < height * width; i += blockDim.x) {
      const int r = i / width;
      plane[r * pitch + (i - r * width)] = __ldg(src + i);
    }
    __syncthreads();
  }
  for (int r = threadIdx.x; r < height; r += blockDim.x) plane[r * pitch + width] = plane[r * pitch + width - 1];
  __syncthreads();
  for (int x = threadIdx.x; x <= width; x += blockDim.x)
    plane[height * pitch + x] = plane[(height - 1) * pitch + x];
  __syncthreads();
}

__device__ __forceinline__ void sts_at(unsigned addr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}

__device__ __forceinline__ float lds_f32(const unsigned char* plane, unsigned off) {
  return *reinterpret_cast<const float*>(plane + off);
}

// The contiguous share [u0, u1) of this CTA out of `total` work units.
__device__ __forceinline__ void cta_share(long long total, long long& u0, long long& u1) {
  u0 = total * blockIdx.x / gridDim.x;
  u1 = total * (blockIdx.x + 1) / gridDim.x;
}

// Output rows of the ROIs whose batch index is outside [0, N): they read nothing, the rows
// are zero (the reference would read out of bounds, ROIAlign.cpp:43-44).
__device__ __forceinline__ void zero_rows_of_invalid_rois(float* __restrict__ out, const int* __restrict__ order,
                                                          const int* __restrict__ img_start, int num_images,
                                                          long long row_stride, long long row_count, int odt = kTopF32) {
  const int i0 = __ldg(img_start + num_images), i1 = __ldg(img_start + num_images + 1);
  long long z0, z1;
  cta_share((long long)(i1 - i0) * row_count, z0, z1);
  for (long long z = z0 + threadIdx.x; z < z1; z += blockDim.x) {
    const int jj = (int)(z / row_count);
    store_top(out, (size_t)__ldg(order + i0 + jj) * row_stride + (size_t)(z - jj * row_count), 0.0f, odt);
  }
}

// The next (plane, ROI range) of this CTA's share: plane (n, c), ROIs [jj0, jj1) of image n.
struct PlaneWork {
  int n, c, j0, jj0, jj1;
};

__device__ __forceinline__ PlaneWork next_plane(const int* __restrict__ img_start, int channels, long long& u,
                                                long long u1, int& n) {
  while ((long long)channels * __ldg(img_start + n + 1) <= u) ++n;
  PlaneWork w;
  w.n = n;
  w.j0 = __ldg(img_start + n);
  const int nroi = __ldg(img_start + n + 1) - w.j0;
  const long long rel = u - (long long)channels * w.j0;
  w.c = (int)(rel / nroi);
  w.jj0 = (int)(rel - (long long)w.c * nroi);
  w.jj1 = (int)min((long long)nroi, w.jj0 + (u1 - u));
  u += w.jj1 - w.jj0;
  return w;
}

// ---- generic forward (sampling ratio 1, 3, 4 and wide pooled shapes): lane = pooled bin ----
template <int G>
__device__ __forceinline__ float pool_bin(const unsigned char* plane, const uint4* __restrict__ ye,
                                          const uint4* __restrict__ xe) {
  uint4 ey[G], ex[G];
#pragma unroll
  for (int i = 0; i < G; ++i) {
    ey[i] = ye[i];
    ex[i] = xe[i];
  }
  float acc = 0.0f;
#pragma unroll
  for (int iy = 0; iy < G; ++iy) {
    const float ly = __uint_as_float(ey[iy].z), hy = __uint_as_float(ey[iy].w);
#pragma unroll
    for (int ix = 0; ix < G; ++ix) {
      const float lx = __uint_as_float(ex[ix].z), hx = __uint_as_float(ex[ix].w);
      const float v1 = lds_f32(plane, ey[iy].x + ex[ix].x), v2 = lds_f32(plane, ey[iy].x + ex[ix].y);
      const float v3 = lds_f32(plane, ey[iy].y + ex[ix].x), v4 = lds_f32(plane, ey[iy].y + ex[ix].y);
      // bilinear.h:50-56
      const float w1 = __fmul_rn(hy, hx), w2 = __fmul_rn(hy, lx);
      const float w3 = __fmul_rn(ly, hx), w4 = __fmul_rn(ly, lx);
      acc = __fadd_rn(acc, tap_sum(w1, v1, w2, v2, w3, v3, w4, v4));
    }
  }
  return acc;
}

// A warp owns one ROI at a time: its lanes fetch the ROI's entries (16 bytes each, coalesced,
// one ROI ahead), park them in the warp's slot of shared memory and every lane then picks
// the entries of its own bins.
template <int G>
__global__ void __launch_bounds__(kFwdThreads, 1)
resident_fwd_kernel(const float* __restrict__ data, float* __restrict__ out, const int* __restrict__ order,
                    const int* __restrict__ img_start, const uint4* __restrict__ entries, int num_images,
                    int channels, int c_count, int height, int width, int pitch, int pooled_h, int pooled_w,
                    int accumulate, int vec16) {
  ptx::chain_enter();
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* plane = reinterpret_cast<float*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int bins = pooled_h * pooled_w;
  const int per_roi = (pooled_h + pooled_w) * G;
  const size_t plane_floats = (size_t)(height + 3) * pitch;
  uint4* slot = reinterpret_cast<uint4*>(smem_raw + plane_floats * 4) + warp * per_roi;
  const size_t plane_elems = (size_t)height * width;
  const CountDiv div(G * G);

  zero_padding(plane, height, width, pitch);
  __syncthreads();
  if (!accumulate)
    zero_rows_of_invalid_rois(out, order, img_start, num_images, (long long)channels * bins,
                              (long long)c_count * bins);

  // this lane's first bin, and the step to its next one (bins lane, lane + 32, ...)
  const int ph0 = lane / pooled_w, pw0 = lane - ph0 * pooled_w;
  const int dph = 32 / pooled_w, dpw = 32 - dph * pooled_w;

  long long u, u1;
  cta_share((long long)c_count * __ldg(img_start + num_images), u, u1);
  int n = 0;
  while (u < u1) {
    const PlaneWork w = next_plane(img_start, c_count, u, u1, n);
    __syncthreads();   // every warp is done reading the previous plane
    load_padded_plane(plane, data + ((size_t)w.n * channels + w.c) * plane_elems, height, width, pitch,
                      vec16);

    int jj = w.jj0 + warp;
    uint4 e0 = make_uint4(0, 0, 0, 0), e1 = e0;
    int roi_next = 0;
    if (jj < w.jj1) {
      const uint4* src = entries + (size_t)(w.j0 + jj) * per_roi;
      if (lane < per_roi) e0 = __ldg(src + lane);
      if (lane + 32 < per_roi) e1 = __ldg(src + lane + 32);
      roi_next = __ldg(order + w.j0 + jj);
    }
    for (; jj < w.jj1; jj += kFwdWarps) {
      __syncwarp();    // the previous ROI's entries are no longer being read
      if (lane < per_roi) slot[lane] = e0;
      if (lane + 32 < per_roi) slot[lane + 32] = e1;
      const int roi = roi_next;
      __syncwarp();
      if (jj + kFwdWarps < w.jj1) {
        const uint4* src = entries + (size_t)(w.j0 + jj + kFwdWarps) * per_roi;
        if (lane < per_roi) e0 = __ldg(src + lane);
        if (lane + 32 < per_roi) e1 = __ldg(src + lane + 32);
        roi_next = __ldg(order + w.j0 + jj + kFwdWarps);
      }
      float* orow = out + ((size_t)roi * channels + w.c) * bins;
      int ph = ph0, pw = pw0;
      for (int b = lane; b < bins; b += 32) {
        const float y = div(pool_bin<G>(smem_raw, slot + ph * G, slot + (pooled_h + pw) * G));
        if (accumulate) orow[b] = __fadd_rn(orow[b], y);
        else orow[b] = y;
        ph += dph;
        pw += dpw;
        if (pw >= pooled_w) {
          pw -= pooled_w;
          ++ph;
        }
      }
    }
  }
}

// ---- sampling_ratio == 2 forward (every BASELINE configuration but the adaptive one) -------
// lane = one sample column of one row of bins.  Half-warp h works on the bin rows
// [ITERS * h, ITERS * h + ITERS) one after the other (consecutive rows: the taps a lane holds carry
// over from one sample row to the next, see Taps), its 16 lanes on the sample columns
// sx = 2 * pw + ix of a chunk of 8 bins; a lane computes the two samples (iy = 0, 1) of its column.  One warp instruction therefore reads two
// feature-map rows at <= 16 columns each; a lane's column record stays in registers for the
// whole ROI; the row records (one 16-byte broadcast load per half-warp and iteration) of the
// NEXT ROI are fetched while the current one is computed, so no load latency is exposed.  The
// two samples of a lane share the packed fp32x2 multiplies of sm_100 (each half rounds exactly
// like the scalar instruction).  The four samples of a bin meet in
// the lane of its first sample through two shuffles, in the reference's order
// ((s00 + s01) + s10) + s11 (ROIAlign.cpp:56-72).
// ITERS = ceil(PH / 2), NCH = chunks of 16 sample columns, ACC = add into the output.
__device__ __forceinline__ float lds_at(unsigned addr, unsigned epoch) {
  float v;
  // `epoch` changes with every plane: it only keeps loads of different planes apart
  asm("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr), "r"(epoch));
  return v;
}
__device__ __forceinline__ float lds_at4(unsigned addr, unsigned epoch) {
  float v;
  asm("ld.shared.f32 %0, [%1+4];" : "=f"(v) : "r"(addr), "r"(epoch));
  return v;
}

// The four taps of one sample column in the two plane rows of a sample row, kept in registers
// from one sample row to the next: (lo, hi) = columns x, x + 1 of the sample's low row `p` and of
// its high row `c`.  `code` (warp-half uniform, from the table): 0 = load both rows, 1 = the low
// row is the previous sample row's high row (registers move, one row is loaded), 2 = the same two
// rows as the previous sample row (nothing is loaded).  The loads are predicated, not branched
// around: the two half-warps of a warp work on different ROIs / bin rows, and a half-warp that
// loads nothing also takes no part in the bank conflicts of the other one's load.
struct Taps {
  float plo, phi, clo, chi;
};
__device__ __forceinline__ void taps_advance(Taps& t, unsigned a, unsigned b, unsigned code, unsigned epoch) {
  asm("{\n\t.reg .pred pl, ps, ph;\n\t"
      "setp.eq.u32 pl, %6, 0;\n\tsetp.eq.u32 ps, %6, 1;\n\tsetp.ne.u32 ph, %6, 2;\n\t"
      "@ps mov.f32 %0, %2;\n\t@ps mov.f32 %1, %3;\n\t"
      "@pl ld.shared.f32 %0, [%4];\n\t@pl ld.shared.f32 %1, [%4+4];\n\t"
      "@ph ld.shared.f32 %2, [%5];\n\t@ph ld.shared.f32 %3, [%5+4];\n\t}"
      : "+f"(t.plo), "+f"(t.phi), "+f"(t.clo), "+f"(t.chi)
      : "r"(a), "r"(b), "r"(code), "r"(epoch));
}
// dense plane: the high column has an address of its own (the low one again where the column is clamped);
// a lane whose sample column is rejected (`live` == 0) loads nothing: its taps stay 0, what the zero
// columns of the padded plane delivered
__device__ __forceinline__ void taps_advance_d(Taps& t, unsigned a, unsigned ah, unsigned b, unsigned bh, unsigned code,
                                               unsigned live, unsigned epoch) {
  asm("{\n\t.reg .pred pl, ps, ph, pv;\n\t"
      "setp.ne.u32 pv, %9, 0;\n\t"
      "setp.eq.and.u32 pl, %8, 0, pv;\n\tsetp.eq.u32 ps, %8, 1;\n\tsetp.ne.and.u32 ph, %8, 2, pv;\n\t"
      "@ps mov.f32 %0, %2;\n\t@ps mov.f32 %1, %3;\n\t"
      "@pl ld.shared.f32 %0, [%4];\n\t@pl ld.shared.f32 %1, [%5];\n\t"
      "@ph ld.shared.f32 %2, [%6];\n\t@ph ld.shared.f32 %3, [%7];\n\t}"
      : "+f"(t.plo), "+f"(t.phi), "+f"(t.clo), "+f"(t.chi)
      : "r"(a), "r"(ah), "r"(b), "r"(bh), "r"(code), "r"(live), "r"(epoch));
}
// the same for a channel pair (float2 per pixel)
struct Taps2 {
  float2 plo, phi, clo, chi;
};
__device__ __forceinline__ void taps2_advance(Taps2& t, unsigned a, unsigned b, unsigned code, unsigned epoch) {
  asm("{\n\t.reg .pred pl, ps, ph;\n\t"
      "setp.eq.u32 pl, %10, 0;\n\tsetp.eq.u32 ps, %10, 1;\n\tsetp.ne.u32 ph, %10, 2;\n\t"
      "@ps mov.f32 %0, %4;\n\t@ps mov.f32 %1, %5;\n\t@ps mov.f32 %2, %6;\n\t@ps mov.f32 %3, %7;\n\t"
      "@pl ld.shared.v2.f32 {%0, %1}, [%8];\n\t@pl ld.shared.v2.f32 {%2, %3}, [%8+8];\n\t"
      "@ph ld.shared.v2.f32 {%4, %5}, [%9];\n\t@ph ld.shared.v2.f32 {%6, %7}, [%9+8];\n\t}"
      : "+f"(t.plo.x), "+f"(t.plo.y), "+f"(t.phi.x), "+f"(t.phi.y), "+f"(t.clo.x), "+f"(t.clo.y), "+f"(t.chi.x),
        "+f"(t.chi.y)
      : "r"(a), "r"(b), "r"(code), "r"(epoch));
}
constexpr unsigned kRowCodeMask = 3u;

#ifndef MOBULA_FWD_DENSE_DEFAULT
#define MOBULA_FWD_DENSE_DEFAULT 1
#endif
#ifndef MOBULA_FWD2_WARPS
#define MOBULA_FWD2_WARPS 16
#endif
constexpr int kFwd2Warps = MOBULA_FWD2_WARPS;
constexpr int kFwd2Threads = kFwd2Warps * 32;

template <int ITERS, int NCH>
struct Roi2Regs {
  uint4 y[ITERS];
  uint2 x[NCH];
  int roi;
};

// Lanes without a sample column / bin row of their own re-read the last existing one: same
// addresses as a working lane, i.e. broadcasts that add no bank conflicts.
template <int ITERS, int NCH>
__device__ __forceinline__ void load_roi2(Roi2Regs<ITERS, NCH>& r, const uint2* __restrict__ ent,
                                          const int* __restrict__ order_j, int h, int q, int last_row,
                                          int last_col) {
  const uint4* ent_y = reinterpret_cast<const uint4*>(ent);
#pragma unroll
  for (int it = 0; it < ITERS; ++it) r.y[it] = __ldg(ent_y + min(ITERS * h + it, last_row));
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch) r.x[ch] = __ldg(ent + 4 * ITERS + min(16 * ch + q, last_col));
  r.roi = __ldg(order_j);
}

template <int ITERS, int NCH, bool ACC, bool TYPED>
__global__ void __launch_bounds__(kFwd2Threads, 1)
resident_fwd2_kernel(const float* __restrict__ data, float* __restrict__ out, const int* __restrict__ order,
                     const int* __restrict__ img_start, const uint2* __restrict__ entries, int num_images,
                     int channels, int c_count, int height, int width, int pitch, int pooled_h, int pooled_w,
                     int vec16, int odt) {
  ptx::chain_enter();
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* plane = reinterpret_cast<float*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int bins = pooled_h * pooled_w;
  constexpr int kPerRoi = 4 * ITERS + 16 * NCH;       // uint2 records
  const size_t plane_elems = (size_t)height * width;
  const int h = lane >> 4, q = lane & 15;
  const unsigned sbase = ptx::smem_addr(smem_raw);
  const unsigned row_bytes = (unsigned)pitch * 4u;

  zero_padding(plane, height, width, pitch);
  __syncthreads();
  if (!ACC)
    zero_rows_of_invalid_rois(out, order, img_start, num_images, (long long)channels * bins,
                              (long long)c_count * bins, TYPED ? odt : kTopF32);

  // which of this lane's results are outputs: the first sample column of a bin, of an
  // existing bin row
  bool stores[NCH];
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch) stores[ch] = (q & 1) == 0 && 8 * ch + (q >> 1) < pooled_w;

  long long u, u1;
  cta_share((long long)c_count * __ldg(img_start + num_images), u, u1);
  int n = 0;
  unsigned epoch = 0;
  while (u < u1) {
    const PlaneWork w = next_plane(img_start, c_count, u, u1, n);
    __syncthreads();   // every warp is done reading the previous plane
    load_padded_plane(plane, data + ((size_t)w.n * channels + w.c) * plane_elems, height, width, pitch,
                      vec16);
    ++epoch;

    int jj = w.jj0 + warp;
    Roi2Regs<ITERS, NCH> nxt;
    if (jj < w.jj1)
      load_roi2(nxt, entries + (size_t)(w.j0 + jj) * kPerRoi, order + w.j0 + jj, h, q, pooled_h - 1,
                2 * pooled_w - 1);
    for (; jj < w.jj1; jj += kFwd2Warps) {
      const Roi2Regs<ITERS, NCH> cur = nxt;
      if (jj + kFwd2Warps < w.jj1)
        load_roi2(nxt, entries + (size_t)(w.j0 + jj + kFwd2Warps) * kPerRoi, order + w.j0 + jj + kFwd2Warps, h, q,
                  pooled_h - 1, 2 * pooled_w - 1);

      unsigned xa[NCH], xb[NCH];      // shared-memory address of the low column in row 0 / row 1
      float2 hl2[NCH];                // (hx, lx)
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) {
        xa[ch] = sbase + cur.x[ch].x;
        xb[ch] = xa[ch] + row_bytes;
        const float lx = __uint_as_float(cur.x[ch].y), hx = __fsub_rn(1.0f, lx);
        hl2[ch] = make_float2(hx, lx);
      }
      // this lane's first output: bin (ITERS * h, q / 2) of chunk 0 -- half-warp h works on the
      // bin rows [ITERS * h, ITERS * h + ITERS), one after the other (row reuse, see Taps)
      float* optr = out + ((size_t)cur.roi * channels + w.c) * bins + (ITERS * h) * pooled_w + (q >> 1);
      Taps tp[NCH];
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) tp[ch].plo = tp[ch].phi = tp[ch].clo = tp[ch].chi = 0.0f;
#pragma unroll
      for (int it = 0; it < ITERS; ++it) {
        const uint4 ey = cur.y[it];
        const float2 ly2 = make_float2(__uint_as_float(ey.y), __uint_as_float(ey.w));
        const float2 hy2 = make_float2(__fsub_rn(1.0f, ly2.x), __fsub_rn(1.0f, ly2.y));
        const bool row_ok = ITERS * h + it < pooled_h;
        // the first sample row of a half-warp has no predecessor in its registers
        const unsigned y0 = ey.x & ~kRowCodeMask, y1 = ey.z & ~kRowCodeMask;
        const unsigned code0 = it == 0 ? 0u : (ey.x & kRowCodeMask), code1 = ey.z & kRowCodeMask;
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
          // one sample at a time (its taps are consumed before the registers move on to the next
          // sample row); the packed multiplies pair the two columns of a row: (w1, w2) = hy * (hx, lx),
          // (w3, w4) = ly * (hx, lx), products (w1 v1, w2 v2) and (w3 v3, w4 v4) as in bilinear.h:50-53.
          // The sums stay scalar: ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 (it
          // honours .rn only for the scalar forms), which would change the rounding
          float2 s;
          {
            taps_advance(tp[ch], xa[ch] + y0, xb[ch] + y0, code0, epoch);
            const float2 wa = __fmul2_rn(make_float2(hy2.x, hy2.x), hl2[ch]), wb = __fmul2_rn(make_float2(ly2.x, ly2.x), hl2[ch]);
            const float2 pa = __fmul2_rn(wa, make_float2(tp[ch].plo, tp[ch].phi));
            const float2 pb = __fmul2_rn(wb, make_float2(tp[ch].clo, tp[ch].chi));
            s.x = __fadd_rn(__fadd_rn(__fadd_rn(pa.x, pa.y), pb.x), pb.y);
          }
          {
            taps_advance(tp[ch], xa[ch] + y1, xb[ch] + y1, code1, epoch);
            const float2 wa = __fmul2_rn(make_float2(hy2.y, hy2.y), hl2[ch]), wb = __fmul2_rn(make_float2(ly2.y, ly2.y), hl2[ch]);
            const float2 pa = __fmul2_rn(wa, make_float2(tp[ch].plo, tp[ch].phi));
            const float2 pb = __fmul2_rn(wb, make_float2(tp[ch].clo, tp[ch].chi));
            s.y = __fadd_rn(__fadd_rn(__fadd_rn(pa.x, pa.y), pb.x), pb.y);
          }
          const float s0n = __shfl_down_sync(0xffffffffu, s.x, 1);
          const float s1n = __shfl_down_sync(0xffffffffu, s.y, 1);
          float t = __fadd_rn(0.0f, s.x);
          t = __fadd_rn(t, s0n);
          t = __fadd_rn(t, s.y);
          t = __fadd_rn(t, s1n);
          const float y = __fmul_rn(t, 0.25f);     // count = 4: an exact scaling
          if (stores[ch] && row_ok) {
            float* o = optr + it * pooled_w + 8 * ch;
            if (ACC) *o = __fadd_rn(*o, y);
            else if (TYPED) store_top(out, (size_t)(o - out), y, odt);      // fp16 / bf16 storage
            else *o = y;
          }
        }
      }
    }
  }
}

// ---- the same forward for an ODD number of bin rows (7 x 7: every BASELINE box head) -----------
// With two bin rows per warp step the last step of an odd pooled height is half empty (1 / 8 of
// the tap loads at 7 x 7).  Here a warp works on TWO ROIs at once -- half-warp h on its own ROI,
// one bin row per step -- so PH steps serve two ROIs instead of PH + 1.  Everything else is the
// kernel above: same records, same arithmetic, same order of sums.  The first kRowsAhead row
// records of the next ROI pair are fetched a pair ahead, the others at the start of the pair
// (they are needed kRowsAhead steps later): registers.
constexpr int kRowsAhead = 4;

template <int PH, int NCH>
struct RoiHead {
  uint4 y[PH < kRowsAhead ? PH : kRowsAhead];
  uint2 x[NCH];
  int roi;
};

template <int PH, int NCH>
__device__ __forceinline__ void load_roi_head(RoiHead<PH, NCH>& r, const uint2* __restrict__ ent,
                                              const int* __restrict__ order_j, int q, int last_col) {
  constexpr int kIters = (PH + 1) / 2;
  const uint4* ent_y = reinterpret_cast<const uint4*>(ent);
#pragma unroll
  for (int it = 0; it < (PH < kRowsAhead ? PH : kRowsAhead); ++it) r.y[it] = __ldg(ent_y + it);
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch) r.x[ch] = __ldg(ent + 4 * kIters + min(16 * ch + q, last_col));
  r.roi = __ldg(order_j);
}

// DENSE: the plane sits in shared memory as it sits in global memory (pitch == width, rows H + 1 and
// H + 2 zero, row H = row H - 1) and arrives through a handful of bulk copies (UBLKCP: no LSU work, no
// registers) instead of 13.6 K 16-byte cp.async per plane, whose issue alone took ~13 % of the kernel.
// What the padding columns did is done by the column records (see resident_fwd2_table_kernel).
template <int PH, int NCH, bool ACC, bool TYPED, bool DENSE = false>
__global__ void __launch_bounds__(kFwd2Threads, 1)
resident_fwd2r_kernel(const float* __restrict__ data, float* __restrict__ out, const int* __restrict__ order,
                      const int* __restrict__ img_start, const uint2* __restrict__ entries, int num_images,
                      int channels, int c_count, int height, int width, int pitch, int pooled_h, int pooled_w,
                      int vec16, int odt) {
  // DENSE: the first plane is on its way before the kernel waits for the table kernel (the planes and
  // the grouping by image -- written by the kernel before the table kernel -- do not depend on it)
  if (DENSE) ptx::chain_trigger();
  else ptx::chain_enter();
  bool waited = !DENSE;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* plane = reinterpret_cast<float*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int bins = pooled_h * pooled_w;
  constexpr int kIters = (PH + 1) / 2;
  constexpr int kPerRoi = 4 * kIters + 16 * NCH;      // uint2 records, as the table kernel lays them out
  constexpr int kHead = PH < kRowsAhead ? PH : kRowsAhead;
  const size_t plane_elems = (size_t)height * width;
  const int h = lane >> 4, q = lane & 15;
  const unsigned sbase = ptx::smem_addr(smem_raw);
  const unsigned row_bytes = (unsigned)pitch * 4u;

  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + (((size_t)(height + 3) * pitch * 4 + 16 + 7) & ~(size_t)7));
  unsigned parity = 0;
  if (DENSE) {
    // rows H + 1, H + 2 and the four floats behind them (high column of the last pixel)
    for (int x = tid; x < 2 * pitch + 4; x += blockDim.x) plane[(height + 1) * pitch + x] = 0.0f;
    if (tid == 0) {
      ptx::mbar_init(bar, 1);
      ptx::fence_mbar_init();
    }
  } else {
    zero_padding(plane, height, width, pitch);
  }
  __syncthreads();
  if (!ACC)
    zero_rows_of_invalid_rois(out, order, img_start, num_images, (long long)channels * bins,
                              (long long)c_count * bins, TYPED ? odt : kTopF32);

  bool stores[NCH];
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch) stores[ch] = (q & 1) == 0 && 8 * ch + (q >> 1) < pooled_w;

  long long u, u1;
  cta_share((long long)c_count * __ldg(img_start + num_images), u, u1);
  int n = 0;
  unsigned epoch = 0;
  while (u < u1) {
    const PlaneWork w = next_plane(img_start, c_count, u, u1, n);
    __syncthreads();   // every warp is done reading the previous plane
    if (DENSE) {
      if (tid == 0) {
        const float* src = data + ((size_t)w.n * channels + w.c) * plane_elems;
        const unsigned row_b = (unsigned)width * 4u;
        ptx::fence_proxy_async();
        ptx::mbar_arrive_expect_tx(bar, (unsigned)(height + 1) * row_b);
        // a few large copies: `vec16` rows each (DENSE: the argument carries the chunk height)
        for (int r = 0; r < height; r += vec16) {
          const int nr = min(vec16, height - r);
          ptx::bulk_g2s(plane + (size_t)r * width, src + (size_t)r * width, (unsigned)nr * row_b, bar);
        }
        ptx::bulk_g2s(plane + (size_t)height * width, src + (size_t)(height - 1) * width, row_b, bar);   // row H
      }
      if (!waited) {
        ptx::chain_wait();
        waited = true;
      }
      ptx::mbar_wait(bar, parity);
      parity ^= 1u;
    } else {
      load_padded_plane(plane, data + ((size_t)w.n * channels + w.c) * plane_elems, height, width, pitch,
                        vec16);
    }
    ++epoch;

    // pair p of this warp: ROIs jj and jj + 1; a half-warp without a ROI of its own repeats the
    // other one's loads (same addresses: broadcasts) and stores nothing
    auto mine = [&](int pair_jj) { return w.j0 + min(pair_jj + h, w.jj1 - 1); };
    int jj = w.jj0 + 2 * warp;
    RoiHead<PH, NCH> nxt;
    if (jj < w.jj1) load_roi_head(nxt, entries + (size_t)mine(jj) * kPerRoi, order + mine(jj), q, 2 * pooled_w - 1);
    for (; jj < w.jj1; jj += 2 * kFwd2Warps) {
      const bool have = jj + h < w.jj1;
      const uint4* ent_y = reinterpret_cast<const uint4*>(entries + (size_t)mine(jj) * kPerRoi);
      uint4 ys[PH];
#pragma unroll
      for (int it = 0; it < kHead; ++it) ys[it] = nxt.y[it];
#pragma unroll
      for (int it = kHead; it < PH; ++it) ys[it] = __ldg(ent_y + it);
      uint2 xs[NCH];
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) xs[ch] = nxt.x[ch];
      const int roi = nxt.roi;
      if (jj + 2 * kFwd2Warps < w.jj1)
        load_roi_head(nxt, entries + (size_t)mine(jj + 2 * kFwd2Warps) * kPerRoi, order + mine(jj + 2 * kFwd2Warps), q,
                      2 * pooled_w - 1);

      unsigned xa[NCH], xb[NCH];      // shared-memory address of the low column in row 0 / row 1
      unsigned xah[NCH], xbh[NCH];    // DENSE: of the high column
      unsigned live[NCH];             // DENSE: 0 = the sample column is rejected
      float2 hl2[NCH];                // (hx, lx)
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) {
        const unsigned xo = xs[ch].x;
        xa[ch] = sbase + (DENSE ? (xo & ~3u) : xo);
        xb[ch] = xa[ch] + row_bytes;
        const float lx = __uint_as_float(xs[ch].y), hx = __fsub_rn(1.0f, lx);
        if (DENSE) {
          const unsigned dcol = (xo & 2u) ? 0u : 4u;
          xah[ch] = xa[ch] + dcol;
          xbh[ch] = xb[ch] + dcol;
          live[ch] = (xo & 1u) ^ 1u;         // a rejected sample column loads nothing: its taps are 0
        }
        hl2[ch] = make_float2(hx, lx);
      }
      // this lane's first output: bin (0, q / 2) of chunk 0 of its half-warp's ROI
      float* optr = out + ((size_t)roi * channels + w.c) * bins + (q >> 1);
      Taps tp[NCH];
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) tp[ch].plo = tp[ch].phi = tp[ch].clo = tp[ch].chi = 0.0f;
#pragma unroll
      for (int it = 0; it < PH; ++it) {
        const uint4 ey = ys[it];
        const float2 ly2 = make_float2(__uint_as_float(ey.y), __uint_as_float(ey.w));
        const float2 hy2 = make_float2(__fsub_rn(1.0f, ly2.x), __fsub_rn(1.0f, ly2.y));
        const unsigned y0 = ey.x & ~kRowCodeMask, y1 = ey.z & ~kRowCodeMask;
        const unsigned code0 = it == 0 ? 0u : (ey.x & kRowCodeMask), code1 = ey.z & kRowCodeMask;
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
          // one sample at a time (its taps are consumed before the registers move on to the next
          // sample row); the packed multiplies pair the two columns of a row: (w1, w2) = hy * (hx, lx),
          // (w3, w4) = ly * (hx, lx), products (w1 v1, w2 v2) and (w3 v3, w4 v4) as in bilinear.h:50-53.
          // The sums stay scalar: ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 (it
          // honours .rn only for the scalar forms), which would change the rounding
          float2 sm;
          {
            if (DENSE) taps_advance_d(tp[ch], xa[ch] + y0, xah[ch] + y0, xb[ch] + y0, xbh[ch] + y0, code0, live[ch], epoch);
            else taps_advance(tp[ch], xa[ch] + y0, xb[ch] + y0, code0, epoch);
            const float2 wa = __fmul2_rn(make_float2(hy2.x, hy2.x), hl2[ch]), wb = __fmul2_rn(make_float2(ly2.x, ly2.x), hl2[ch]);
            const float2 pa = __fmul2_rn(wa, make_float2(tp[ch].plo, tp[ch].phi));
            const float2 pb = __fmul2_rn(wb, make_float2(tp[ch].clo, tp[ch].chi));
            sm.x = __fadd_rn(__fadd_rn(__fadd_rn(pa.x, pa.y), pb.x), pb.y);
          }
          {
            if (DENSE) taps_advance_d(tp[ch], xa[ch] + y1, xah[ch] + y1, xb[ch] + y1, xbh[ch] + y1, code1, live[ch], epoch);
            else taps_advance(tp[ch], xa[ch] + y1, xb[ch] + y1, code1, epoch);
            const float2 wa = __fmul2_rn(make_float2(hy2.y, hy2.y), hl2[ch]), wb = __fmul2_rn(make_float2(ly2.y, ly2.y), hl2[ch]);
            const float2 pa = __fmul2_rn(wa, make_float2(tp[ch].plo, tp[ch].phi));
            const float2 pb = __fmul2_rn(wb, make_float2(tp[ch].clo, tp[ch].chi));
            sm.y = __fadd_rn(__fadd_rn(__fadd_rn(pa.x, pa.y), pb.x), pb.y);
          }
          const float s0n = __shfl_down_sync(0xffffffffu, sm.x, 1);
          const float s1n = __shfl_down_sync(0xffffffffu, sm.y, 1);
          float t = __fadd_rn(0.0f, sm.x);
          t = __fadd_rn(t, s0n);
          t = __fadd_rn(t, sm.y);
          t = __fadd_rn(t, s1n);
          const float y = __fmul_rn(t, 0.25f);     // count = 4: an exact scaling
          if (stores[ch] && have) {
            float* o = optr + it * pooled_w + 8 * ch;
            if (ACC) *o = __fadd_rn(*o, y);
            else if (TYPED) store_top(out, (size_t)(o - out), y, odt);      // fp16 / bf16 storage
            else *o = y;
          }
        }
      }
    }
  }
}

// ---- the same forward over a channel PAIR (planes small enough that two padded planes fit) ----
// The plane holds two channels interleaved (float2 per pixel): every tap is one 8-byte load that
// serves both channels, a half-warp's load touches ONE feature-map row (8-byte accesses are
// served per half-warp), and the weights, addresses and table look-ups are shared by the pair.
__device__ __forceinline__ float2 lds2_at(unsigned addr, unsigned epoch) {
  float2 v;
  asm("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr), "r"(epoch));
  return v;
}
__device__ __forceinline__ float2 lds2_at8(unsigned addr, unsigned epoch) {
  float2 v;
  asm("ld.shared.v2.f32 {%0, %1}, [%2+8];" : "=f"(v.x), "=f"(v.y) : "r"(addr), "r"(epoch));
  return v;
}

// two dense planes (the second may be absent) -> one padded float2 plane in shared memory
__device__ __forceinline__ void load_padded_pair(float2* plane, const float* __restrict__ src0,
                                                 const float* __restrict__ src1, int height, int width,
                                                 int pitch) {
  if ((width & 3) == 0 && (((uintptr_t)src0 | (uintptr_t)src1) & 15u) == 0) {
    const int w4 = width >> 2, total = height * w4;
    int r = threadIdx.x / w4, c4 = threadIdx.x - r * w4;
    const int dr = blockDim.x / w4, dc = blockDim.x - dr * w4;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(src0) + i);
      const float4 b = src1 ? __ldg(reinterpret_cast<const float4*>(src1) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      float4* dst = reinterpret_cast<float4*>(plane + (size_t)r * pitch + 4 * c4);
      dst[0] = make_float4(a.x, b.x, a.y, b.y);
      dst[1] = make_float4(a.z, b.z, a.w, b.w);
      r += dr;
      c4 += dc;
      if (c4 >= w4) {
        c4 -= w4;
        ++r;
      }
    }
  } else {
    for (int i = threadIdx.x; i < height * width; i += blockDim.x) {
      const int r = i / width;
      plane[r * pitch + (i - r * width)] = make_float2(__ldg(src0 + i), src1 ? __ldg(src1 + i) : 0.0f);
    }
  }
  __syncthreads();
  for (int r = threadIdx.x; r < height; r += blockDim.x) plane[r * pitch + width] = plane[r * pitch + width - 1];
  __syncthreads();
  for (int x = threadIdx.x; x <= width; x += blockDim.x)
    plane[height * pitch + x] = plane[(height - 1) * pitch + x];
  __syncthreads();
}

template <int ITERS, int NCH, bool ACC, bool TYPED>
__global__ void __launch_bounds__(kFwd2Threads, 1)
resident_fwd2p_kernel(const float* __restrict__ data, float* __restrict__ out, const int* __restrict__ order,
                      const int* __restrict__ img_start, const uint2* __restrict__ entries, int num_images,
                      int channels, int c_count, int height, int width, int pitch, int pooled_h, int pooled_w, int odt) {
  ptx::chain_enter();
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float2* plane = reinterpret_cast<float2*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int bins = pooled_h * pooled_w;
  constexpr int kPerRoi = 4 * ITERS + 16 * NCH;       // uint2 records
  const size_t plane_elems = (size_t)height * width;
  const int h = lane >> 4, q = lane & 15;
  const unsigned sbase = ptx::smem_addr(smem_raw);
  const unsigned row_bytes = (unsigned)pitch * 8u;
  const int cpairs = (c_count + 1) >> 1;

  // the same padding as zero_padding(), on a plane of 2 * pitch floats per row
  zero_padding(reinterpret_cast<float*>(smem_raw), height, 2 * width + 1, 2 * pitch);
  __syncthreads();
  if (!ACC)
    zero_rows_of_invalid_rois(out, order, img_start, num_images, (long long)channels * bins,
                              (long long)c_count * bins, TYPED ? odt : kTopF32);

  bool stores[NCH];
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch) stores[ch] = (q & 1) == 0 && 8 * ch + (q >> 1) < pooled_w;

  long long u, u1;
  cta_share((long long)cpairs * __ldg(img_start + num_images), u, u1);
  int n = 0;
  unsigned epoch = 0;
  while (u < u1) {
    const PlaneWork w = next_plane(img_start, cpairs, u, u1, n);      // w.c: channel pair
    const int c = 2 * w.c;
    const bool two = c + 1 < c_count;
    __syncthreads();   // every warp is done reading the previous plane
    const float* src0 = data + ((size_t)w.n * channels + c) * plane_elems;
    load_padded_pair(plane, src0, two ? src0 + plane_elems : nullptr, height, width, pitch);
    ++epoch;

    int jj = w.jj0 + warp;
    Roi2Regs<ITERS, NCH> nxt;
    if (jj < w.jj1)
      load_roi2(nxt, entries + (size_t)(w.j0 + jj) * kPerRoi, order + w.j0 + jj, h, q, pooled_h - 1,
                2 * pooled_w - 1);
    for (; jj < w.jj1; jj += kFwd2Warps) {
      const Roi2Regs<ITERS, NCH> cur = nxt;
      if (jj + kFwd2Warps < w.jj1)
        load_roi2(nxt, entries + (size_t)(w.j0 + jj + kFwd2Warps) * kPerRoi, order + w.j0 + jj + kFwd2Warps, h, q,
                  pooled_h - 1, 2 * pooled_w - 1);

      unsigned xa[NCH], xb[NCH];      // shared-memory address of the low column in row 0 / row 1
      float lx[NCH], hx[NCH];
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) {
        xa[ch] = sbase + cur.x[ch].x;
        xb[ch] = xa[ch] + row_bytes;
        lx[ch] = __uint_as_float(cur.x[ch].y);
        hx[ch] = __fsub_rn(1.0f, lx[ch]);
      }
      // half-warp h: bin rows [ITERS * h, ITERS * h + ITERS), see resident_fwd2_kernel
      float* optr = out + ((size_t)cur.roi * channels + c) * bins + (ITERS * h) * pooled_w + (q >> 1);
      Taps2 tp[NCH];
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) tp[ch].plo = tp[ch].phi = tp[ch].clo = tp[ch].chi = make_float2(0.f, 0.f);
#pragma unroll
      for (int it = 0; it < ITERS; ++it) {
        const uint4 ey = cur.y[it];
        const float ly0 = __uint_as_float(ey.y), ly1 = __uint_as_float(ey.w);
        const float hy0 = __fsub_rn(1.0f, ly0), hy1 = __fsub_rn(1.0f, ly1);
        const bool row_ok = ITERS * h + it < pooled_h;
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
          // one bilinear sample of both channels (.x / .y); products packed over the pair, sums
          // scalar (see resident_fwd2_kernel), order as in bilinear.h:50-56
          auto sample = [&](unsigned row_code, bool first, float hy, float ly) {
            const unsigned row = row_code & ~kRowCodeMask;
            taps2_advance(tp[ch], xa[ch] + row, xb[ch] + row, first ? 0u : (row_code & kRowCodeMask), epoch);
            const float2 v1 = tp[ch].plo, v2 = tp[ch].phi, v3 = tp[ch].clo, v4 = tp[ch].chi;
            const float w1 = __fmul_rn(hy, hx[ch]), w2 = __fmul_rn(hy, lx[ch]);
            const float w3 = __fmul_rn(ly, hx[ch]), w4 = __fmul_rn(ly, lx[ch]);
            const float2 p1 = __fmul2_rn(make_float2(w1, w1), v1), p2 = __fmul2_rn(make_float2(w2, w2), v2);
            const float2 p3 = __fmul2_rn(make_float2(w3, w3), v3), p4 = __fmul2_rn(make_float2(w4, w4), v4);
            float2 s;
            s.x = __fadd_rn(__fadd_rn(__fadd_rn(p1.x, p2.x), p3.x), p4.x);
            s.y = __fadd_rn(__fadd_rn(__fadd_rn(p1.y, p2.y), p3.y), p4.y);
            return s;
          };
          const float2 s0 = sample(ey.x, it == 0, hy0, ly0), s1 = sample(ey.z, false, hy1, ly1);
          float2 t;
          t.x = __fadd_rn(0.0f, s0.x);
          t.y = __fadd_rn(0.0f, s0.y);
          t.x = __fadd_rn(t.x, __shfl_down_sync(0xffffffffu, s0.x, 1));
          t.y = __fadd_rn(t.y, __shfl_down_sync(0xffffffffu, s0.y, 1));
          t.x = __fadd_rn(t.x, s1.x);
          t.y = __fadd_rn(t.y, s1.y);
          t.x = __fadd_rn(t.x, __shfl_down_sync(0xffffffffu, s1.x, 1));
          t.y = __fadd_rn(t.y, __shfl_down_sync(0xffffffffu, s1.y, 1));
          if (stores[ch] && row_ok) {
            float* o = optr + it * pooled_w + 8 * ch;
            const float y0 = __fmul_rn(t.x, 0.25f), y1 = __fmul_rn(t.y, 0.25f);     // count = 4: exact
            if (ACC) {
              o[0] = __fadd_rn(o[0], y0);
              if (two) o[bins] = __fadd_rn(o[bins], y1);
            } else if (TYPED) {      // fp16 / bf16 storage
              store_top(out, (size_t)(o - out), y0, odt);
              if (two) store_top(out, (size_t)(o - out) + bins, y1, odt);
            } else {
              o[0] = y0;
              if (two) o[bins] = y1;
            }
          }
        }
      }
    }
  }
}

// The channel-pair forward for odd pooled heights: two ROIs per warp, as resident_fwd2r_kernel.
template <int PH, int NCH, bool ACC, bool TYPED>
__global__ void __launch_bounds__(kFwd2Threads, 1)
resident_fwd2pr_kernel(const float* __restrict__ data, float* __restrict__ out, const int* __restrict__ order,
                      const int* __restrict__ img_start, const uint2* __restrict__ entries, int num_images,
                      int channels, int c_count, int height, int width, int pitch, int pooled_h, int pooled_w, int odt) {
  ptx::chain_enter();
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float2* plane = reinterpret_cast<float2*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int bins = pooled_h * pooled_w;
  constexpr int kIters = (PH + 1) / 2;
  constexpr int kPerRoi = 4 * kIters + 16 * NCH;      // uint2 records, as the table kernel lays them out
  constexpr int kHead = PH < kRowsAhead ? PH : kRowsAhead;
  const size_t plane_elems = (size_t)height * width;
  const int h = lane >> 4, q = lane & 15;
  const unsigned sbase = ptx::smem_addr(smem_raw);
  const unsigned row_bytes = (unsigned)pitch * 8u;
  const int cpairs = (c_count + 1) >> 1;

  // the same padding as zero_padding(), on a plane of 2 * pitch floats per row
  zero_padding(reinterpret_cast<float*>(smem_raw), height, 2 * width + 1, 2 * pitch);
  __syncthreads();
  if (!ACC)
    zero_rows_of_invalid_rois(out, order, img_start, num_images, (long long)channels * bins,
                              (long long)c_count * bins, TYPED ? odt : kTopF32);

  bool stores[NCH];
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch) stores[ch] = (q & 1) == 0 && 8 * ch + (q >> 1) < pooled_w;

  long long u, u1;
  cta_share((long long)cpairs * __ldg(img_start + num_images), u, u1);
  int n = 0;
  unsigned epoch = 0;
  while (u < u1) {
    const PlaneWork w = next_plane(img_start, cpairs, u, u1, n);      // w.c: channel pair
    const int c = 2 * w.c;
    const bool two = c + 1 < c_count;
    __syncthreads();   // every warp is done reading the previous plane
    const float* src0 = data + ((size_t)w.n * channels + c) * plane_elems;
    load_padded_pair(plane, src0, two ? src0 + plane_elems : nullptr, height, width, pitch);
    ++epoch;

    // two ROIs per warp, half-warp h on its own ROI (see resident_fwd2r_kernel)
    auto mine = [&](int pair_jj) { return w.j0 + min(pair_jj + h, w.jj1 - 1); };
    int jj = w.jj0 + 2 * warp;
    RoiHead<PH, NCH> nxt;
    if (jj < w.jj1) load_roi_head(nxt, entries + (size_t)mine(jj) * kPerRoi, order + mine(jj), q, 2 * pooled_w - 1);
    for (; jj < w.jj1; jj += 2 * kFwd2Warps) {
      const bool have = jj + h < w.jj1;
      const uint4* ent_y = reinterpret_cast<const uint4*>(entries + (size_t)mine(jj) * kPerRoi);
      struct {
        uint4 y[PH];
        uint2 x[NCH];
        int roi;
      } cur;
#pragma unroll
      for (int it = 0; it < kHead; ++it) cur.y[it] = nxt.y[it];
#pragma unroll
      for (int it = kHead; it < PH; ++it) cur.y[it] = __ldg(ent_y + it);
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) cur.x[ch] = nxt.x[ch];
      cur.roi = nxt.roi;
      if (jj + 2 * kFwd2Warps < w.jj1)
        load_roi_head(nxt, entries + (size_t)mine(jj + 2 * kFwd2Warps) * kPerRoi, order + mine(jj + 2 * kFwd2Warps), q,
                      2 * pooled_w - 1);

      unsigned xa[NCH], xb[NCH];      // shared-memory address of the low column in row 0 / row 1
      float lx[NCH], hx[NCH];
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) {
        xa[ch] = sbase + cur.x[ch].x;
        xb[ch] = xa[ch] + row_bytes;
        lx[ch] = __uint_as_float(cur.x[ch].y);
        hx[ch] = __fsub_rn(1.0f, lx[ch]);
      }
      float* optr = out + ((size_t)cur.roi * channels + c) * bins + (q >> 1);
      Taps2 tp[NCH];
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) tp[ch].plo = tp[ch].phi = tp[ch].clo = tp[ch].chi = make_float2(0.f, 0.f);
#pragma unroll
      for (int it = 0; it < PH; ++it) {
        const uint4 ey = cur.y[it];
        const float ly0 = __uint_as_float(ey.y), ly1 = __uint_as_float(ey.w);
        const float hy0 = __fsub_rn(1.0f, ly0), hy1 = __fsub_rn(1.0f, ly1);
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
          // one bilinear sample of both channels (.x / .y); products packed over the pair, sums
          // scalar (see resident_fwd2_kernel), order as in bilinear.h:50-56
          auto sample = [&](unsigned row_code, bool first, float hy, float ly) {
            const unsigned row = row_code & ~kRowCodeMask;
            taps2_advance(tp[ch], xa[ch] + row, xb[ch] + row, first ? 0u : (row_code & kRowCodeMask), epoch);
            const float2 v1 = tp[ch].plo, v2 = tp[ch].phi, v3 = tp[ch].clo, v4 = tp[ch].chi;
            const float w1 = __fmul_rn(hy, hx[ch]), w2 = __fmul_rn(hy, lx[ch]);
            const float w3 = __fmul_rn(ly, hx[ch]), w4 = __fmul_rn(ly, lx[ch]);
            const float2 p1 = __fmul2_rn(make_float2(w1, w1), v1), p2 = __fmul2_rn(make_float2(w2, w2), v2);
            const float2 p3 = __fmul2_rn(make_float2(w3, w3), v3), p4 = __fmul2_rn(make_float2(w4, w4), v4);
            float2 s;
            s.x = __fadd_rn(__fadd_rn(__fadd_rn(p1.x, p2.x), p3.x), p4.x);
            s.y = __fadd_rn(__fadd_rn(__fadd_rn(p1.y, p2.y), p3.y), p4.y);
            return s;
          };
          const float2 s0 = sample(ey.x, it == 0, hy0, ly0), s1 = sample(ey.z, false, hy1, ly1);
          float2 t;
          t.x = __fadd_rn(0.0f, s0.x);
          t.y = __fadd_rn(0.0f, s0.y);
          t.x = __fadd_rn(t.x, __shfl_down_sync(0xffffffffu, s0.x, 1));
          t.y = __fadd_rn(t.y, __shfl_down_sync(0xffffffffu, s0.y, 1));
          t.x = __fadd_rn(t.x, s1.x);
          t.y = __fadd_rn(t.y, s1.y);
          t.x = __fadd_rn(t.x, __shfl_down_sync(0xffffffffu, s1.x, 1));
          t.y = __fadd_rn(t.y, __shfl_down_sync(0xffffffffu, s1.y, 1));
          if (stores[ch] && have) {
            float* o = optr + it * pooled_w + 8 * ch;
            const float y0 = __fmul_rn(t.x, 0.25f), y1 = __fmul_rn(t.y, 0.25f);     // count = 4: exact
            if (ACC) {
              o[0] = __fadd_rn(o[0], y0);
              if (two) o[bins] = __fadd_rn(o[bins], y1);
            } else if (TYPED) {      // fp16 / bf16 storage
              store_top(out, (size_t)(o - out), y0, odt);
              if (two) store_top(out, (size_t)(o - out) + bins, y1, odt);
            } else {
              o[0] = y0;
              if (two) o[bins] = y1;
            }
          }
        }
      }
    }
  }
}

template <typename K>
cudaError_t set_smem(K kernel, size_t bytes) {
  return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

template <typename K>
cudaError_t resident_grid(K kernel, int threads, size_t smem, int num_sms, long long work, unsigned* grid) {
  cudaError_t e = set_smem(kernel, smem);
  if (e != cudaSuccess) return e;
  int per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem);
  if (e != cudaSuccess) return e;
  if (per_sm < 1) return cudaErrorLaunchOutOfResources;
  long long g = (long long)num_sms * per_sm;
  if (g > work) g = work;
  *grid = (unsigned)g;
  return cudaSuccess;
}

// ==========================================================================================
// backward (sampling_ratio == 2), "atomic"-mode replacement without atomics
// ==========================================================================================
// A CTA owns a slab of consecutive rows of one (image, channel) plane of dX in shared memory;
// warp w of the CTA owns a band of rows of the slab and nobody else touches them, so the
// gradient is accumulated with plain shared-memory read-modify-writes and the finished slab
// goes to HBM once, with plain stores (the zero-fill of ROIAlign.py:33-34 is fused: every dX
// element is written exactly once and never read).
//
// The scatter of one ROI is separable: the 2*2*PW (column, weight) taps of its sample columns
// are merged per touched column once per call (table kernel), a lane owns one touched column
// and first folds the dy row of a bin row with its column weights,
//     T[col] = sum_pw Wx[col][pw] * dy[ph][pw] / count,
// then adds wy * T into the <= 4 feature-map rows the two sample rows of that bin row touch.
// Per bin row and touched row that is ONE conflict-free-ish shared-memory read-modify-write
// for all touched columns, instead of 2*2*PW atomics.  Contributions of one ROI to one pixel
// are summed before they are added, and the order of ROIs is fixed (ownership is static), so
// the result is reproducible run to run but not bit-identical to the reference's
// single-thread order (that is the deterministic mode, roialign_plane.cu).
#ifndef MOBULA_BWD_WARPS
#define MOBULA_BWD_WARPS 8
#endif
constexpr int kBwdWarps = MOBULA_BWD_WARPS;
constexpr int kBwdThreads = kBwdWarps * 32;
constexpr int kMaxSlotFloats = 256;   // dy tile of one (roi, channel): PH * PW <= 256 floats
constexpr int kMaxHitsPerCol = 16;    // = max pooled_w

// Column records of one ROI (see the table kernel below): `mine` = the sample column of this
// lane.  The even lane of a bin column merges the <= 4 taps of its two samples per touched
// column, the (column, bin column, weight) triples of the ROI are sorted, and a run of one
// column becomes a column record (two hits inline, the rest in the overflow list).
__device__ __forceinline__ void bwd2_column_records(const AxisTap& mine, int lane, int wib, int j, int pooled_w,
                                                    int maxc, uint4* __restrict__ cols, uint2* __restrict__ extra,
                                                    int2* __restrict__ crange, unsigned (*s_tkey)[64],
                                                    unsigned (*s_ckey)[64], float (*s_tw)[64], float (*s_cw)[64]) {
  {   // the columns this ROI touches: which column tiles have to visit it
    const int cmin = __reduce_min_sync(0xffffffffu, mine.valid ? mine.lo : 0x7fffffff);
    const int cmax = __reduce_max_sync(0xffffffffu, mine.valid ? mine.hi : -1);
    if (lane == 0) crange[j] = make_int2(cmin, cmax);
  }
  uint4* crow = cols + (size_t)j * maxc;
  uint2* erow = extra + (size_t)j * maxc;
  int tcol[4] = {0, 0, 0, 0}, ntap = 0;
  float tw[4] = {0.0f, 0.0f, 0.0f, 0.0f};
  {
    // the odd sample of the bin column comes over from the next lane
    const int o_valid = __shfl_down_sync(0xffffffffu, mine.valid ? 1 : 0, 1);
    const int o_lo = __shfl_down_sync(0xffffffffu, mine.lo, 1), o_hi = __shfl_down_sync(0xffffffffu, mine.hi, 1);
    const float o_l = __shfl_down_sync(0xffffffffu, mine.wl, 1);
    if ((lane & 1) == 0 && lane < 2 * pooled_w) {
      auto add = [&](int col, float w) {
        for (int k = 0; k < ntap; ++k)
          if (tcol[k] == col) {
            tw[k] += w;
            return;
          }
        tcol[ntap] = col;
        tw[ntap] = w;
        ++ntap;
      };
      // 1/count folded into the column weights: count = 4, an exact scaling
      auto sample = [&](int lo, int hi, float l) {
        add(lo, __fmul_rn(__fsub_rn(1.0f, l), 0.25f));
        if (hi != lo) add(hi, __fmul_rn(l, 0.25f));      // the clamped sample has no second column
      };
      if (mine.valid) sample(mine.lo, mine.hi, mine.wl);
      if (o_valid) sample(o_lo, o_hi, o_l);
    }
  }
  int tincl = ntap;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, tincl, d);
    if (lane >= d) tincl += v;
  }
  const int nt = __shfl_sync(0xffffffffu, tincl, 31);      // <= 4 * pooled_w <= 64 triples
  for (int k = 0; k < ntap; ++k) {
    s_tkey[wib][tincl - ntap + k] = (unsigned)tcol[k] * 16u + (unsigned)(lane >> 1);
    s_tw[wib][tincl - ntap + k] = tw[k];
  }
  __syncwarp();
  for (int e = lane; e < nt; e += 32) {                     // rank sort by (column, bin column): keys are unique
    const unsigned key = s_tkey[wib][e];
    int rank = 0;
    for (int k = 0; k < nt; ++k) rank += s_tkey[wib][k] < key ? 1 : 0;
    s_ckey[wib][rank] = key;
    s_cw[wib][rank] = s_tw[wib][e];
  }
  __syncwarp();
  int ncols = 0;
  {
    // triple i = lane (+ 32): head of its column's run?  position in the run?  overflow hits before it?
    unsigned heads[2], more[2];
    int run_pos[2];
#pragma unroll
    for (int hset = 0; hset < 2; ++hset) {
      const int i = lane + 32 * hset;
      const bool in = i < nt;
      const bool head = in && (i == 0 || (s_ckey[wib][i - 1] >> 4) != (s_ckey[wib][i] >> 4));
      heads[hset] = __ballot_sync(0xffffffffu, head);
    }
    const unsigned le = 0xffffffffu >> (31 - lane);          // lanes <= this one
#pragma unroll
    for (int hset = 0; hset < 2; ++hset) {
      const int i = lane + 32 * hset;
      int h;                                                  // index of the run's head
      if (hset == 1 && (heads[1] & le)) h = 32 + 31 - __clz(heads[1] & le);
      else h = 31 - __clz(hset == 1 ? heads[0] : (heads[0] & le));
      run_pos[hset] = i - h;
      more[hset] = __ballot_sync(0xffffffffu, i < nt && run_pos[hset] >= 2);
    }
    ncols = __popc(heads[0]) + __popc(heads[1]);
    const unsigned lt = le >> 1;                              // lanes < this one
#pragma unroll
    for (int hset = 0; hset < 2; ++hset) {
      const int i = lane + 32 * hset;
      if (i >= nt) continue;
      const unsigned key = s_ckey[wib][i];
      const int before = hset == 0 ? __popc(more[0] & lt) : __popc(more[0]) + __popc(more[1] & lt);
      if (run_pos[hset] >= 2) {
        erow[before] = make_uint2(key & 15u, __float_as_uint(s_cw[wib][i]));
      } else if (run_pos[hset] == 0) {
        // length of the run: up to the next head
        unsigned above = hset == 0 ? (heads[0] & ~le) : 0u, above1 = hset == 0 ? heads[1] : (heads[1] & ~le);
        int next = nt;
        if (above) next = __ffs(above) - 1;
        else if (above1) next = 32 + __ffs(above1) - 1;
        const int nh = next - i;
        const int slot = hset == 0 ? __popc(heads[0] & lt) : __popc(heads[0]) + __popc(heads[1] & lt);
        uint4 r;
        r.x = (key >> 4) * 8u | ((unsigned)nh << 16) | ((key & 15u) << 24) |
              ((nh > 1 ? s_ckey[wib][i + 1] & 15u : 0u) << 28);
        r.y = __float_as_uint(s_cw[wib][i]);
        r.z = __float_as_uint(nh > 1 ? s_cw[wib][i + 1] : 0.0f);
        r.w = (unsigned)before;                               // its overflow hits start here
        crow[slot] = r;
      }
    }
  }
  for (int k = ncols + lane; k < maxc; k += 32) crow[k] = make_uint4(0u, 0u, 0u, 0u);   // unused slots
}

// Tables of the backward, per ROI row j (channel independent):
//   cols[j][maxc]   column records, one per touched column: x = col_off | nhits << 16 |
//                   pw0 << 24 | pw1 << 28 (nhits = 0: unused slot), y = w0, z = w1, w = index of
//                   the third hit in extra[j][]
//   rows[j][maxe]   bin-row records, 32 bytes each, sorted by (band, bin row): the <= 4 rows
//                   of bin row ph that lie in one band.  First uint4: x = ph | rows << 4,
//                   y = off0 | off1 << 16, z = off2 | off3 << 16 (byte offsets of the rows from
//                   the band's first row); second uint4: the merged row weights
//   bstart[b][j]    index of the first record of band b (rows [b * band_rows, ...)), for
//                   b = 0 .. nbands; bstart[nbands][j] = number of records
// and per (image, band) the ordered list of visits (roialign_bwd2_visits_kernel).
// Two warps per ROI, every step across the lanes: one warp builds the bin-row records, the other
// one the column records.
__global__ void __launch_bounds__(256)
resident_bwd2_table_kernel(const float* __restrict__ rois, int* __restrict__ order,
                           int* __restrict__ img_start, int num_images, int num_rois, int height,
                           int width, int row_bytes, int col_bytes, int pooled_h, int pooled_w, float scale,
                           int maxe, int maxc, int band_rows, int nbands, uint4* __restrict__ rows_out,
                           unsigned char* __restrict__ bstart, uint4* __restrict__ cols,
                           uint2* __restrict__ extra, int2* __restrict__ crange,
                           int* __restrict__ counter) {
  ptx::chain_enter();
  __shared__ unsigned s_key[4][64], s_sband[4][64];
  __shared__ uint4 s_ra[4][64], s_rb[4][64];
  __shared__ unsigned s_tkey[4][64], s_ckey[4][64];
  __shared__ float s_tw[4][64], s_cw[4][64];
  if (blockIdx.x == 0 && threadIdx.x == 0) counter[0] = counter[1] = counter[2] = 0;   // unit queue, visit blocks done, heaviest pair
  // the first num_images + 1 blocks group the ROI rows by image (nobody in this launch reads
  // the result), the others build the tables, indexed by ROI ROW
  if ((int)blockIdx.x <= num_images) {
    __shared__ int s_sums[8];
    roi_group_block<256>(rois, num_rois, num_images, blockIdx.x, img_start, order, s_sums);
    return;
  }
  const int wib = threadIdx.x >> 6, lane = threadIdx.x & 31;
  const bool column_warp = (threadIdx.x >> 5) & 1;
  const int j = ((int)blockIdx.x - num_images - 1) * 4 + wib;          // the ROI row
  if (j >= num_rois || segment_of(rois, j, num_images) == num_images) return;
  const int roi = j;
  const RoiGeom g = roi_decode(rois + (size_t)roi * 5, scale, pooled_h, pooled_w, 2);
  AxisTap mine;
  mine.valid = false;
  mine.lo = mine.hi = 0;
  mine.wl = mine.wh = 0.0f;
  if (lane < 2 * pooled_w) mine = axis_decode(sample_pos(g.start_w, lane >> 1, g.bin_w, lane & 1, 2), width);
  if (column_warp) {
    bwd2_column_records(mine, lane, wib, j, pooled_w, maxc, cols, extra, crange, s_tkey, s_ckey, s_tw, s_cw);
    return;
  }
  // no column inside the plane: the ROI contributes nothing
  const bool has_columns = __any_sync(0xffffffffu, mine.valid);

  // ---- rows: bin row `lane` touches <= 4 distinct rows (ascending), weights merged ----
  int nrow = 0, rrow[4] = {0, 0, 0, 0};
  float rw[4] = {0.0f, 0.0f, 0.0f, 0.0f};
  if (lane < pooled_h) {
    auto add = [&](int row, float w) {
      for (int k = 0; k < nrow; ++k)
        if (rrow[k] == row) {
          rw[k] += w;
          return;
        }
      rrow[nrow] = row;
      rw[nrow] = w;
      ++nrow;
    };
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const AxisTap t = axis_decode(sample_pos(g.start_h, lane, g.bin_h, i, 2), height);
      if (t.valid) {
        add(t.lo, t.wh);
        if (t.hi != t.lo) add(t.hi, t.wl);     // the clamped sample has no second row
      }
    }
  }
  // records of this bin row: runs of its (ascending) rows inside one band
  int nrec = 0, rec_first[4] = {0, 0, 0, 0}, rec_n[4] = {0, 0, 0, 0};
  for (int k = 0; k < nrow; ++k) {
    if (nrec == 0 || rrow[k] / band_rows != rrow[rec_first[nrec - 1]] / band_rows) {
      rec_first[nrec] = k;
      rec_n[nrec] = 1;
      ++nrec;
    } else {
      ++rec_n[nrec - 1];
    }
  }
  // exclusive prefix over the lanes -> position of this bin row's records in the ROI's list
  int incl = nrec;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += v;
  }
  int ne = __shfl_sync(0xffffffffu, incl, 31);
  const int base = incl - nrec;

  if (!has_columns) {
    ne = 0;
    nrec = 0;
  }

  // ---- bin-row records, sorted by (band, bin row) so that the records of a band are contiguous ----
  for (int q = 0; q < nrec; ++q) {
    const int k0 = rec_first[q], band = rrow[k0] / band_rows, r0 = band * band_rows;
    unsigned off[4] = {0u, 0u, 0u, 0u};
    float w[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    for (int i = 0; i < rec_n[q]; ++i) {
      off[i] = (unsigned)((rrow[k0 + i] - r0) * row_bytes);
      w[i] = rw[k0 + i];
    }
    s_key[wib][base + q] = (unsigned)(band * 16 + lane);
    s_ra[wib][base + q] = make_uint4((unsigned)lane | ((unsigned)rec_n[q] << 4), off[0] | (off[1] << 16),
                                     off[2] | (off[3] << 16), 0u);
    s_rb[wib][base + q] = make_uint4(__float_as_uint(w[0]), __float_as_uint(w[1]), __float_as_uint(w[2]),
                                     __float_as_uint(w[3]));
  }
  __syncwarp();
  uint4* rout = rows_out + (size_t)j * maxe * 2;
  for (int e = lane; e < ne; e += 32) {
    const unsigned key = s_key[wib][e];
    int rank = 0;
    for (int k = 0; k < ne; ++k) rank += s_key[wib][k] < key ? 1 : 0;
    rout[2 * rank] = s_ra[wib][e];
    rout[2 * rank + 1] = s_rb[wib][e];
    s_sband[wib][rank] = key >> 4;
  }
  __syncwarp();
  // first record of every band (bands ascend with the rank)
  for (int bnd = lane; bnd <= nbands; bnd += 32) {
    int lo = 0, hi = ne;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (s_sband[wib][mid] < (unsigned)bnd) lo = mid + 1;
      else hi = mid;
    }
    bstart[(size_t)bnd * num_rois + j] = (unsigned char)(bnd == nbands ? ne : lo);
  }
}

// Tables of the DETERMINISTIC backward: the same shape as above, but nothing is merged --
// every record keeps the individual bilinear taps so that the kernel can add the contributions
// to a pixel one by one, in the reference's single-thread order (roi, ph, pw, iy, ix, tap):
//   rows[j][maxe]    uint4: x = row * row_bytes | ph, y = flags (bit iy: sample row iy of bin row
//                    ph touches this row), z / w = its row weight for iy = 0 / 1
//   colh[j][maxc]    uint4: x = col_off | nhits << 16, y = index of the third hit in extra[j][],
//                    z = pwA | flagsA << 8 | pwB << 16 | flagsB << 24 (bit ix: sample column ix
//                    of bin column pw touches this column)
//   colw[j][maxc]    uint4: column weights of hit A for ix = 0 / 1, of hit B for ix = 0 / 1
//   extra[j][maxc]   uint4: x = pw | flags << 8, z / w = weights for ix = 0 / 1
// A clamped sample (low = high tap, bilinear.h:32-44) keeps only its low tap: the high one has
// weight 0 and adding +-0 never changes a sum that started at +0.
__global__ void __launch_bounds__(128)
resident_bwd2d_table_kernel(const float* __restrict__ rois, int* __restrict__ order,
                            int* __restrict__ img_start, int num_images, int num_rois, int height,
                            int width, int row_bytes, int col_bytes, int pooled_h, int pooled_w, float scale,
                            int maxe, int maxc, int band_rows, int nbands, uint4* __restrict__ rows_out,
                            unsigned char* __restrict__ bstart, uint4* __restrict__ colh,
                            uint4* __restrict__ colw, uint4* __restrict__ extra, int2* __restrict__ crange,
                            int* __restrict__ counter) {
  ptx::chain_enter();
  __shared__ unsigned s_key[4][64], s_srow[4][64], s_rfl[4][64];
  __shared__ float s_w0[4][64], s_w1[4][64];
  __shared__ int s_lo[4][32], s_hi[4][32];
  __shared__ float s_fl[4][32];
  // open columns A / B of the merge: per hit (pw, flags, weight ix=0, weight ix=1)
  __shared__ int s_pw[4][2][kMaxHitsPerCol], s_cf[4][2][kMaxHitsPerCol];
  __shared__ float s_c0[4][2][kMaxHitsPerCol], s_c1[4][2][kMaxHitsPerCol];
  if (blockIdx.x == 0 && threadIdx.x == 0) counter[0] = counter[1] = counter[2] = 0;   // unit queue, visit blocks done, heaviest pair
  if ((int)blockIdx.x <= num_images) {       // grouping blocks, see resident_bwd2_table_kernel
    __shared__ int s_sums[4];
    roi_group_block<128>(rois, num_rois, num_images, blockIdx.x, img_start, order, s_sums);
    return;
  }
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = ((int)blockIdx.x - num_images - 1) * 4 + wib;          // the ROI row
  if (j >= num_rois || segment_of(rois, j, num_images) == num_images) return;
  const int roi = j;
  const RoiGeom g = roi_decode(rois + (size_t)roi * 5, scale, pooled_h, pooled_w, 2);

  // ---- rows ----
  int nrow = 0, rrow[4] = {0, 0, 0, 0};
  unsigned rfl[4] = {0u, 0u, 0u, 0u};
  float rw0[4] = {0.0f, 0.0f, 0.0f, 0.0f}, rw1[4] = {0.0f, 0.0f, 0.0f, 0.0f};
  if (lane < pooled_h) {
    auto add = [&](int row, int iy, float w) {
      int k = 0;
      while (k < nrow && rrow[k] != row) ++k;
      if (k == nrow) {
        rrow[k] = row;
        ++nrow;
      }
      rfl[k] |= 1u << iy;
      if (iy) rw1[k] = w;
      else rw0[k] = w;
    };
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const AxisTap t = axis_decode(sample_pos(g.start_h, lane, g.bin_h, i, 2), height);
      if (t.valid) {
        add(t.lo, i, t.wh);
        if (t.hi != t.lo) add(t.hi, i, t.wl);
      }
    }
  }
  int incl = nrow;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += v;
  }
  int ne = __shfl_sync(0xffffffffu, incl, 31);
  const int base = incl - nrow;

  // ---- columns ----
  AxisTap mine;
  mine.valid = false;
  mine.lo = mine.hi = 0;
  mine.wl = mine.wh = 0.0f;
  if (lane < 2 * pooled_w) mine = axis_decode(sample_pos(g.start_w, lane >> 1, g.bin_w, lane & 1, 2), width);
  {   // the columns this ROI touches: which column tiles have to visit it
    const int cmin = __reduce_min_sync(0xffffffffu, mine.valid ? mine.lo : 0x7fffffff);
    const int cmax = __reduce_max_sync(0xffffffffu, mine.valid ? mine.hi : -1);
    if (lane == 0) crange[j] = make_int2(cmin, cmax);
  }
  uint4* hrow = colh + (size_t)j * maxc;
  uint4* wrow = colw + (size_t)j * maxc;
  uint4* erow = extra + (size_t)j * maxc;
  int ncols = 0;
  s_lo[wib][lane] = mine.valid ? mine.lo : -1;
  s_hi[wib][lane] = mine.hi;
  s_fl[wib][lane] = mine.wl;
  __syncwarp();
  if (lane == 0) {
    int colA = -1, nA = 0, nB = 0, nextra = 0, A = 0;      // A: which of the two slots is column A
    auto hit = [&](int slot, int& n, int pw, int ix, float w) {
      if (n == 0 || s_pw[wib][slot][n - 1] != pw) {
        s_pw[wib][slot][n] = pw;
        s_cf[wib][slot][n] = 0;
        s_c0[wib][slot][n] = s_c1[wib][slot][n] = 0.0f;
        ++n;
      }
      s_cf[wib][slot][n - 1] |= 1 << ix;
      if (ix) s_c1[wib][slot][n - 1] = w;
      else s_c0[wib][slot][n - 1] = w;
    };
    auto flushA = [&]() {
      if (nA > 0 && colA < width) {
        const int* pw = s_pw[wib][A];
        const int* cf = s_cf[wib][A];
        const float* c0 = s_c0[wib][A];
        const float* c1 = s_c1[wib][A];
        hrow[ncols] = make_uint4((unsigned)(colA * col_bytes) | ((unsigned)nA << 16), (unsigned)nextra,
                                 (unsigned)pw[0] | ((unsigned)cf[0] << 8) |
                                     ((unsigned)(nA > 1 ? pw[1] : 0) << 16) | ((unsigned)(nA > 1 ? cf[1] : 0) << 24),
                                 0u);
        wrow[ncols] = make_uint4(__float_as_uint(c0[0]), __float_as_uint(c1[0]),
                                 __float_as_uint(nA > 1 ? c0[1] : 0.0f), __float_as_uint(nA > 1 ? c1[1] : 0.0f));
        for (int k = 2; k < nA; ++k)
          erow[nextra++] = make_uint4((unsigned)pw[k] | ((unsigned)cf[k] << 8), 0u, __float_as_uint(c0[k]),
                                      __float_as_uint(c1[k]));
        ++ncols;
      }
      colA += 1;
      A ^= 1;             // B becomes A
      nA = nB;
      nB = 0;
    };
    for (int s = 0; s < 2 * pooled_w; ++s) {
      const int lo = s_lo[wib][s];
      if (lo < 0) continue;
      const int hi = s_hi[wib][s];
      const float l = s_fl[wib][s], h = __fsub_rn(1.0f, l);
      if (colA < 0) colA = lo;
      while (colA < lo) {
        if (nA == 0 && nB == 0) {
          colA = lo;
          break;
        }
        flushA();
      }
      hit(A, nA, s >> 1, s & 1, h);
      if (hi != lo) hit(A ^ 1, nB, s >> 1, s & 1, l);
    }
    flushA();
    flushA();
  }
  ncols = __shfl_sync(0xffffffffu, ncols, 0);
  for (int k = ncols + lane; k < maxc; k += 32) hrow[k] = make_uint4(0u, 0u, 0u, 0u);   // unused slots
  if (ncols == 0) {
    ne = 0;
    nrow = 0;
  }

  // ---- row entries sorted by (band, bin row, row) ----
  for (int k = 0; k < nrow; ++k) {
    s_key[wib][base + k] = (unsigned)((rrow[k] / band_rows) * 16 + lane) * 65536u + (unsigned)rrow[k];
    s_rfl[wib][base + k] = rfl[k];
    s_w0[wib][base + k] = rw0[k];
    s_w1[wib][base + k] = rw1[k];
  }
  __syncwarp();
  uint4* rout = rows_out + (size_t)j * maxe;
  for (int e = lane; e < ne; e += 32) {
    const unsigned key = s_key[wib][e];
    int rank = 0;
    for (int k = 0; k < ne; ++k) rank += s_key[wib][k] < key ? 1 : 0;
    const unsigned row = key & 0xffffu, ph = (key >> 16) & 15u;
    rout[rank] = make_uint4(row * (unsigned)row_bytes | ph, s_rfl[wib][e], __float_as_uint(s_w0[wib][e]),
                            __float_as_uint(s_w1[wib][e]));
    s_srow[wib][rank] = row;
  }
  __syncwarp();
  for (int bnd = lane; bnd <= nbands; bnd += 32) {
    const unsigned limit = (unsigned)(bnd * band_rows);
    int lo = 0, hi = ne;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (s_srow[wib][mid] < limit) lo = mid + 1;
      else hi = mid;
    }
    bstart[(size_t)bnd * num_rois + j] = (unsigned char)(bnd == nbands ? ne : lo);
  }
}

// One block per (image, band): the visits of the band, in ROI order.  Thread t owns a
// contiguous run of the image's ROIs; a block-wide prefix sum places its visits.
// visit = x: roi row | first entry << 20 | (entries - 1) << 26; y: roi row
// The block that finishes last then orders the (image, band) pairs by work, heaviest first
// (`uorder`): the main kernel's warps take units from a queue, and a unit is as long as its
// band's visit list, so the long ones must start first for the warps to finish together.
constexpr int kVisitThreads = 512;
constexpr int kWorkBuckets = 64;
__global__ void __launch_bounds__(kVisitThreads)
resident_bwd2_visits_kernel(const int* __restrict__ order, const int* __restrict__ img_start,
                            const unsigned char* __restrict__ bstart, const int2* __restrict__ crange,
                            int num_images, int num_rois, int nbands, int ntx, int tile_cols,
                            uint2* __restrict__ visits, int* __restrict__ vcount, int* __restrict__ uwork,
                            int* __restrict__ uorder, int* __restrict__ counter) {
  ptx::chain_enter();
  __shared__ int s_warp[kVisitThreads / 32];
  __shared__ int s_entries, s_last, s_hist[kWorkBuckets];
  if (threadIdx.x == 0) s_entries = 0;
  const int w = blockIdx.x, t = threadIdx.x, lane = t & 31, wid = t >> 5;
  // tile w = ((image, band), column tile): rows [bnd * band_rows, ...), columns [c0, c1)
  const int tiles = nbands * ntx;
  const int n = w / tiles, bt = w - n * tiles, bnd = bt / ntx, tx = bt - bnd * ntx;
  const int c0 = tx * tile_cols, c1 = c0 + tile_cols;
  const int j0 = __ldg(img_start + n), nroi = __ldg(img_start + n + 1) - j0;
  uint2* list = visits + ((size_t)j0 * tiles + (size_t)bt * nroi);
  const unsigned char* s0 = bstart + (size_t)bnd * num_rois;
  const unsigned char* s1 = s0 + num_rois;
  const int per = (nroi + kVisitThreads - 1) / kVisitThreads;
  const int jb = min(nroi, t * per), je = min(nroi, jb + per);
  // (the tables are indexed by ROI row r = order[j])
  auto meets = [&](int r) {
    if (s1[r] <= s0[r]) return false;
    if (ntx == 1) return true;
    const int2 cr = crange[r];
    return cr.y >= c0 && cr.x < c1;
  };
  int mine = 0;
  for (int jj = jb; jj < je; ++jj) mine += meets(__ldg(order + j0 + jj)) ? 1 : 0;
  int incl = mine;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += v;
  }
  if (lane == 31) s_warp[wid] = incl;
  __syncthreads();
  int before = incl - mine, total = 0;
  for (int k = 0; k < kVisitThreads / 32; ++k) {
    const int v = s_warp[k];
    if (k < wid) before += v;
    total += v;
  }
  int entries = 0;
  for (int jj = jb; jj < je; ++jj) {
    const int r = __ldg(order + j0 + jj);
    const int first = s0[r], cnt = s1[r] - first;
    if (meets(r)) {
      list[before++] = make_uint2((unsigned)r | ((unsigned)first << 20) | ((unsigned)(cnt - 1) << 26),
                                  (unsigned)r);
      entries += cnt;
    }
  }
  if (entries) atomicAdd(&s_entries, entries);
  __syncthreads();
  if (t == 0) {
    vcount[w] = total;
    const int mywork = 3 * total + s_entries;      // a visit costs about three row entries
    uwork[w] = mywork;
    atomicMax(counter + 2, mywork);                // the heaviest pair, for the classes below
    __threadfence();
    s_last = atomicAdd(counter + 1, 1) == (int)gridDim.x - 1 ? 1 : 0;
  }
  for (int k = t; k < kWorkBuckets; k += kVisitThreads) s_hist[k] = 0;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // counting sort of the pairs by work, in kWorkBuckets classes, heaviest class first
  const int pairs = (int)gridDim.x;
  const volatile int* work = uwork;
  const long long top = max(1, *(const volatile int*)(counter + 2));
  auto bucket = [&](int v) { return (kWorkBuckets - 1) - (int)((long long)v * (kWorkBuckets - 1) / top); };
  for (int k = t; k < pairs; k += kVisitThreads) atomicAdd(&s_hist[bucket(work[k])], 1);
  __syncthreads();
  if (t < 32) {                                   // exclusive prefix over the classes, one warp
    static_assert(kWorkBuckets == 64, "two classes per lane");
    const int a = s_hist[2 * t], b = s_hist[2 * t + 1];
    int incl = a + b;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, d);
      if (t >= d) incl += v;
    }
    s_hist[2 * t] = incl - a - b;
    s_hist[2 * t + 1] = incl - b;
  }
  __syncthreads();
  for (int k = t; k < pairs; k += kVisitThreads) uorder[atomicAdd(&s_hist[bucket(work[k])], 1)] = k;
}

// 8-byte shared-memory accesses on a 32-bit shared address; volatile: program order is kept
__device__ __forceinline__ float2 lds64(unsigned addr) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sts64(unsigned addr, float2 v) {
  asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(v.x), "f"(v.y) : "memory");
}
__device__ __forceinline__ float lds_vol(unsigned addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ uint4 lds128(unsigned addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "r"(addr)
               : "memory");
  return v;
}
__device__ __forceinline__ void sts128(unsigned addr, uint4 v) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}
__device__ __forceinline__ uint2 lds64u(unsigned addr) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr) : "memory");
  return v;
}

// Everything a warp needs for one visit, fetched while the previous visit is computed: NT floats
// of the dy tile pair per lane, the band's records of the ROI (merged mode: one 32-byte bin-row
// record per lane, at most 16; deterministic mode: NE 16-byte row entries), NCK column records.
constexpr int kMaxVisitRecords = 16;     // = max pooled_h: a bin row has one record per band
template <int NCK, bool DET>
struct VisitCols {
  uint4 col[NCK];
  uint4 colw[DET ? NCK : 1];
};

template <int NT, int NCK, int NE, bool DET>
struct VisitRegs {
  float g[NT];
  uint4 e[DET ? NE : 2];
  uint4 col[NCK];
  uint4 colw[DET ? NCK : 1];      // deterministic: the column weights
};

// Every warp is on its own: it takes (image, channel pair, band) units from a queue, keeps the
// band -- `band_rows` rows x two channels, float2 per pixel -- in its private shared memory,
// walks the band's visit list and stores the finished band to the two dX planes.
// DET: contributions are added one by one in the reference's single-thread order (bit-exact dX)
// instead of merged per ROI.
template <int NT, int NCK, int NE, bool ACC, bool DET>
__global__ void __launch_bounds__(kBwdThreads, 2)
resident_bwd2_kernel(const float* __restrict__ dy, float* __restrict__ dx, const uint2* __restrict__ visits,
                     const int* __restrict__ vcount, const int* __restrict__ uorder,
                     const void* __restrict__ rows_v,
                     const uint4* __restrict__ cols, const uint4* __restrict__ colw,
                     const void* __restrict__ extra_v,
                     const int* __restrict__ img_start, int* __restrict__ counter, int num_images,
                     int channels, int c_count, int height, int width, int pitch2, int pooled_h, int pooled_w,
                     int maxe, int maxc, int nbands, int band_rows, int ntx, int tile_cols) {
  ptx::chain_enter();
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int bins = pooled_h * pooled_w;
  const unsigned rb = (unsigned)pitch2 * 8u;                 // bytes per band row
  // per warp: the band, then the dy tiles of the channel pair (NT * 32 floats) and the row
  // entries of the visit (NE * 32 * 8 bytes)
  constexpr unsigned kMaxE = NE * 32u, kMaxC = NCK * 32u;      // = maxe, maxc of the tables
  (void)maxe;
  (void)maxc;
  const uint4* rows_in = static_cast<const uint4*>(rows_v);
  const unsigned slot_bytes =
      (unsigned)(NT * 32) * 4u + (DET ? (NE * 32u + 2u) * 16u : (unsigned)kMaxVisitRecords * 32u);
  const unsigned warp_bytes = (unsigned)band_rows * rb + slot_bytes;
  float2* band = reinterpret_cast<float2*>(smem_raw + (size_t)warp * warp_bytes);
  const unsigned sbase = ptx::smem_addr(band);
  const unsigned slot_s = sbase + (unsigned)band_rows * rb;
  const unsigned eslot_s = slot_s + NT * 128u;
  const size_t plane_elems = (size_t)height * width;
  const int cpairs = (c_count + 1) >> 1;
  const int tiles = nbands * ntx;                       // per plane: row bands x column tiles
  const int units = num_images * cpairs * tiles;
  const int w4 = width >> 2;

  const unsigned tile_s = slot_s + (unsigned)lane * 4u;       // this lane's first tile float
  const unsigned ch1 = (unsigned)bins * 4u;                   // second channel's tile follows the first
  const unsigned cb = (unsigned)channels * (unsigned)bins;

  int unit_next = 0;
  if (lane == 0) unit_next = atomicAdd(counter, 1);
  for (;;) {
    const int unit = __shfl_sync(0xffffffffu, unit_next, 0);
    if (unit >= units) break;
    if (lane == 0) unit_next = atomicAdd(counter, 1);          // in flight while this unit runs
    // units in queue order: (image, tile) pairs by decreasing work, then the channel pairs
    const int rank = unit / cpairs, c = 2 * (unit - rank * cpairs);
    const int pair = __ldg(uorder + rank), n = pair / tiles, bt = pair - n * tiles;
    const int bnd = bt / ntx, tx = bt - bnd * ntx;
    const int c0 = tx * tile_cols, cw = min(width, c0 + tile_cols) - c0;     // the tile's columns
    const bool two = c + 1 < c_count;
    const int tile_floats = two ? 2 * bins : bins;
    const int r0 = bnd * band_rows, r1 = min(height, r0 + band_rows);
    float* gplane = dx + ((size_t)n * channels + c) * plane_elems;
    const int j0 = __ldg(img_start + n), nroi = __ldg(img_start + n + 1) - j0;
    const int nvis = __ldg(vcount + pair);
    const uint2* vlist = visits + ((size_t)j0 * tiles + (size_t)bt * nroi);
    // visit records: every lane reads the same 8 bytes, two visits before they are staged
    uint2 rec0 = nvis > 0 ? __ldg(vlist) : make_uint2(0u, 0u);
    uint2 rec1 = nvis > 1 ? __ldg(vlist + 1) : make_uint2(0u, 0u);

    // ---- band := 0 (or the current dX rows when accumulating) ----
    __syncwarp();
    if (ACC) {
      for (int i = lane; i < (r1 - r0) * cw; i += 32) {
        const int r = i / cw, x = i - r * cw;
        const size_t gi = (size_t)(r0 + r) * width + c0 + x;
        band[r * pitch2 + x] = make_float2(gplane[gi], two ? gplane[plane_elems + gi] : 0.0f);
      }
    } else {
      float4* s4 = reinterpret_cast<float4*>(band);
      const int n4 = ((r1 - r0) * pitch2) >> 1;
      for (int i = lane; i < n4; i += 32) s4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncwarp();

    const unsigned col_base = sbase - (unsigned)r0 * rb - (unsigned)c0 * 8u;   // + row offset + column offset
    const unsigned tile_lo = (unsigned)c0 * 8u, tile_span = (unsigned)cw * 8u;   // column offsets of the tile
    const float* dy_c = dy + (size_t)c * bins;
    // float lane + 32 * i of the tile pair exists (one address per visit, constant offsets)
    bool tile_ok[NT];
#pragma unroll
    for (int i = 0; i < NT; ++i) tile_ok[i] = lane + 32 * i < tile_floats;

    // visit k with its record `rec`
#define MOBULA_LOAD_VISIT(v, rec)                                                                  \
  do {                                                                                            \
    const unsigned vx_ = (rec).x, roi_ = (rec).y;                                                 \
    const unsigned jv_ = vx_ & 0xfffffu, first_ = (vx_ >> 20) & 63u, cnt_ = (vx_ >> 26) + 1u;       \
    const float* tile_ = dy_c + ((size_t)roi_ * cb + (unsigned)lane);                             \
    _Pragma("unroll") for (int i = 0; i < NT; ++i)                                                \
        if (tile_ok[i])(v).g[i] = __ldg(tile_ + 32 * i);                                          \
    if (DET) {                                                                                    \
      const uint4* er_ = rows_in + (jv_ * kMaxE + first_ + (unsigned)lane);                       \
      _Pragma("unroll") for (int i = 0; i < (DET ? NE : 0); ++i)                                  \
          if ((unsigned)lane + 32u * i < cnt_)(v).e[i] = __ldg(er_ + 32 * i);                     \
    } else if ((unsigned)lane < cnt_) {                                                           \
      const uint4* er_ = rows_in + (size_t)(jv_ * kMaxE + first_ + (unsigned)lane) * 2;           \
      (v).e[0] = __ldg(er_);                                                                      \
      (v).e[1] = __ldg(er_ + 1);                                                                  \
    }                                                                                             \
    const uint4* cr_ = cols + (jv_ * kMaxC + (unsigned)lane);                            \
    _Pragma("unroll") for (int i = 0; i < NCK; ++i)(v).col[i] = __ldg(cr_ + 32 * i);              \
    if (DET) {                                                                                    \
      const uint4* cw_ = colw + (jv_ * kMaxC + (unsigned)lane);                          \
      _Pragma("unroll") for (int i = 0; i < NCK; ++i)(v).colw[i] = __ldg(cw_ + 32 * i);           \
    }                                                                                             \
  } while (0)

    // one visit, first half: the ROI's dy tiles and the band's row entries go to the warp's slot
    auto publish = [&](const VisitRegs<NT, NCK, NE, DET>& cur, unsigned vx) {
      const unsigned cnt = (vx >> 26) + 1u;
      __syncwarp();
#pragma unroll
      for (int i = 0; i < NT; ++i) sts_at(tile_s + 128u * i, cur.g[i]);
      if constexpr (DET) {
#pragma unroll
        for (int i = 0; i < NE; ++i)
          if ((unsigned)lane + 32u * i < cnt) sts128(eslot_s + (lane + 32u * i) * 16u, cur.e[i]);
      } else if ((unsigned)lane < cnt) {
        sts128(eslot_s + lane * 32u, cur.e[0]);
        sts128(eslot_s + lane * 32u + 16u, cur.e[1]);
      }
      __syncwarp();
    };
    // second half: every touched column (lane) adds its share to the band's rows
    auto apply = [&](const VisitCols<NCK, DET>& cur, unsigned vx) {
      const unsigned jcur = vx & 0xfffffu, cnt = (vx >> 26) + 1u;
      if constexpr (DET) {
        const uint4* extra = static_cast<const uint4*>(extra_v);
#pragma unroll
        for (int ck = 0; ck < NCK; ++ck) {
          const uint4 ch = cur.col[ck], cw = cur.colw[ck];
          const int nh = (int)((ch.x >> 16) & 0xffu);
          // this lane has no column in this chunk, or its column belongs to another tile
          if (nh == 0 || (ch.x & 0xffffu) - tile_lo >= tile_span) continue;
          const unsigned caddr = col_base + (ch.x & 0xffffu);
          for (unsigned e = 0; e < cnt; ++e) {
            const uint4 re = lds128(eslot_s + e * 16u);           // broadcast
            const unsigned ph = re.x & 15u, rfl = re.y;
            const float wy0 = __uint_as_float(re.z), wy1 = __uint_as_float(re.w);
            const unsigned a = caddr + (re.x & ~15u);
            const unsigned grow = slot_s + ph * (unsigned)pooled_w * 4u;
            float2 v = lds64(a);
            // the reference's order at this pixel: pw, then iy, then ix (ROIAlign.cpp:126-157);
            // g = (dtop * (wy * wx)) / count, count = 4
            auto hit = [&](unsigned pw, unsigned cfl, float wx0, float wx1) {
              const float d0 = lds_vol(grow + pw * 4u), d1 = lds_vol(grow + pw * 4u + ch1);
              auto add = [&](float wy, float wx) {
                const float w = __fmul_rn(wy, wx);
                v.x = __fadd_rn(v.x, __fmul_rn(__fmul_rn(d0, w), 0.25f));
                v.y = __fadd_rn(v.y, __fmul_rn(__fmul_rn(d1, w), 0.25f));
              };
              if (rfl & 1u) {
                if (cfl & 1u) add(wy0, wx0);
                if (cfl & 2u) add(wy0, wx1);
              }
              if (rfl & 2u) {
                if (cfl & 1u) add(wy1, wx0);
                if (cfl & 2u) add(wy1, wx1);
              }
            };
            hit(ch.z & 0xffu, (ch.z >> 8) & 0xffu, __uint_as_float(cw.x), __uint_as_float(cw.y));
            if (nh > 1) {
              hit((ch.z >> 16) & 0xffu, ch.z >> 24, __uint_as_float(cw.z), __uint_as_float(cw.w));
              for (int x = 2; x < nh; ++x) {
                const uint4 hx = __ldg(extra + (size_t)jcur * kMaxC + ch.y + (x - 2));
                hit(hx.x & 0xffu, (hx.x >> 8) & 0xffu, __uint_as_float(hx.z), __uint_as_float(hx.w));
              }
            }
            sts64(a, v);
          }
        }
        return;
      }
      const uint2* extra = static_cast<const uint2*>(extra_v);
#pragma unroll
      for (int ck = 0; ck < NCK; ++ck) {
        const uint4 cr = cur.col[ck];
        const int nh = (int)((cr.x >> 16) & 0xffu);
        // this lane has no column in this chunk, or its column belongs to another tile
        if (nh == 0 || (cr.x & 0xffffu) - tile_lo >= tile_span) continue;
        const unsigned caddr = sbase + ((cr.x & 0xffffu) - tile_lo);     // + row offset inside the band
        const unsigned g0 = slot_s + ((cr.x >> 24) & 0xfu) * 4u, g1 = slot_s + (cr.x >> 28) * 4u;
        const float w0 = __uint_as_float(cr.y), w1 = __uint_as_float(cr.z);
        // one bin row at a time: T = sum_pw Wx[col][pw] * dy[ph][pw] for the channel pair, then
        // its <= 4 rows in this band, all distinct: the read-modify-writes overlap
        for (unsigned e = 0; e < cnt; ++e) {
          const uint4 ra = lds128(eslot_s + e * 32u), rw = lds128(eslot_s + e * 32u + 16u);   // broadcasts
          const unsigned grow = (ra.x & 15u) * (unsigned)pooled_w * 4u, nr = ra.x >> 4;
          float2 t;
          t.x = w0 * lds_vol(g0 + grow);
          t.y = w0 * lds_vol(g0 + grow + ch1);
          if (nh > 1) {
            t.x = fmaf(w1, lds_vol(g1 + grow), t.x);
            t.y = fmaf(w1, lds_vol(g1 + grow + ch1), t.y);
            for (int x = 2; x < nh; ++x) {
              const uint2 hx = __ldg(extra + (size_t)jcur * kMaxC + cr.w + (x - 2));
              const unsigned ge = slot_s + hx.x * 4u + grow;
              t.x = fmaf(__uint_as_float(hx.y), lds_vol(ge), t.x);
              t.y = fmaf(__uint_as_float(hx.y), lds_vol(ge + ch1), t.y);
            }
          }
          const unsigned a0 = caddr + (ra.y & 0xffffu), a1 = caddr + (ra.y >> 16);
          const unsigned a2 = caddr + (ra.z & 0xffffu), a3 = caddr + (ra.z >> 16);
          const float y0 = __uint_as_float(rw.x), y1 = __uint_as_float(rw.y);
          const float y2 = __uint_as_float(rw.z), y3 = __uint_as_float(rw.w);
          if (nr <= 2u) {                                   // warp-uniform
            const float2 v0 = lds64(a0);
            float2 v1 = make_float2(0.0f, 0.0f);
            if (nr == 2u) v1 = lds64(a1);
            sts64(a0, make_float2(fmaf(y0, t.x, v0.x), fmaf(y0, t.y, v0.y)));
            if (nr == 2u) sts64(a1, make_float2(fmaf(y1, t.x, v1.x), fmaf(y1, t.y, v1.y)));
          } else {
            const float2 v0 = lds64(a0), v1 = lds64(a1), v2 = lds64(a2);
            float2 v3 = make_float2(0.0f, 0.0f);
            if (nr == 4u) v3 = lds64(a3);
            sts64(a0, make_float2(fmaf(y0, t.x, v0.x), fmaf(y0, t.y, v0.y)));
            sts64(a1, make_float2(fmaf(y1, t.x, v1.x), fmaf(y1, t.y, v1.y)));
            sts64(a2, make_float2(fmaf(y2, t.x, v2.x), fmaf(y2, t.y, v2.y)));
            if (nr == 4u) sts64(a3, make_float2(fmaf(y3, t.x, v3.x), fmaf(y3, t.y, v3.y)));
          }
        }
      }
    };

    // Two register sets in turn.  A visit first publishes its set, then starts the loads of the
    // visit two ahead into the registers just freed, then computes: the scoreboard wait of the
    // publishing stores (every global load of this kernel shares one scoreboard) therefore
    // covers only loads that have been in flight for a whole visit, never the ones just issued.
    // STAGE(set, x, k): visit k (k < nvis) -> register set; the record of visit k + 2 is requested
#define MOBULA_STAGE(v, vx, k)                                                                    \
  do {                                                                                            \
    if ((k) < nvis) {                                                                             \
      const uint2 rec_ = rec0;                                                                    \
      rec0 = rec1;                                                                                \
      if ((k) + 2 < nvis) rec1 = __ldg(vlist + (k) + 2);                                          \
      vx = rec_.x;                                                                                \
      MOBULA_LOAD_VISIT(v, rec_);                                                                 \
    }                                                                                             \
  } while (0)
#define MOBULA_VISIT(v, vx, knext)                                                                \
  do {                                                                                            \
    VisitCols<NCK, DET> cols_;                                                                    \
    _Pragma("unroll") for (int i = 0; i < NCK; ++i) {                                             \
      cols_.col[i] = (v).col[i];                                                                  \
      if (DET) cols_.colw[i] = (v).colw[i];                                                       \
    }                                                                                             \
    const unsigned x_ = vx;                                                                       \
    publish(v, x_);                                                                               \
    MOBULA_STAGE(v, vx, knext);                                                                   \
    apply(cols_, x_);                                                                             \
  } while (0)
    {
      VisitRegs<NT, NCK, NE, DET> va, vb;
      unsigned xa = 0u, xb = 0u;
      MOBULA_STAGE(va, xa, 0);
      MOBULA_STAGE(vb, xb, 1);
      for (int k = 0; k < nvis; k += 2) {
        MOBULA_VISIT(va, xa, k + 2);
        if (k + 1 >= nvis) break;
        MOBULA_VISIT(vb, xb, k + 3);
      }
    }
#undef MOBULA_VISIT
#undef MOBULA_STAGE
#undef MOBULA_LOAD_VISIT

    __syncwarp();
    // ---- the finished band -> the two planes of dX, each element written once ----
    float* gplane1 = gplane + plane_elems;
    if (((width | c0 | cw) & 3) == 0 && ((uintptr_t)dx & 15u) == 0) {
      const int t4 = cw >> 2, total = (r1 - r0) * t4;          // the tile's row segments, 16 bytes at a time
      int r = lane / t4, c4 = lane - r * t4;
      const int dr = 32 / t4, dc = 32 - dr * t4;
      float4* ga = reinterpret_cast<float4*>(gplane + (size_t)r0 * width + c0);
      float4* gb = reinterpret_cast<float4*>(gplane1 + (size_t)r0 * width + c0);
      for (int i = lane; i < total; i += 32) {
        const float4* src = reinterpret_cast<const float4*>(band + (size_t)r * pitch2 + 4 * c4);
        const float4 lo = src[0], hi = src[1];       // (a0 b0 a1 b1) (a2 b2 a3 b3)
        const size_t gi = (size_t)r * w4 + c4;
        ga[gi] = make_float4(lo.x, lo.z, hi.x, hi.z);
        if (two) gb[gi] = make_float4(lo.y, lo.w, hi.y, hi.w);
        r += dr;
        c4 += dc;
        if (c4 >= t4) {
          c4 -= t4;
          ++r;
        }
      }
    } else {
      for (int i = lane; i < (r1 - r0) * cw; i += 32) {
        const int r = i / cw, x = i - r * cw;
        const float2 v = band[r * pitch2 + x];
        const size_t gi = (size_t)(r0 + r) * width + c0 + x;
        gplane[gi] = v.x;
        if (two) gplane1[gi] = v.y;
      }
    }
  }
}

struct BwdLayout {
  size_t img_start, order, rows, bstart, cols, colw, extra, crange, visits, vcount, uwork, uorder, counter, total;
};

BwdLayout bwd_layout(const RoiAlignArgs& a, int maxe, int maxc, int nbands, int ntx, bool det) {
  BwdLayout l;
  size_t p = 0;
  const size_t R = (size_t)a.num_rois, N = (size_t)(a.num_images > 0 ? a.num_images : 0);
  l.img_start = p;  p += align_up((N + 2) * sizeof(int), 256);
  l.order = p;      p += align_up(R * sizeof(int), 256);
  l.rows = p;       p += align_up(R * maxe * (det ? sizeof(uint4) : 2 * sizeof(uint4)), 256);   // merged: 32-byte records
  l.bstart = p;     p += align_up(R * (size_t)(nbands + 1), 256);
  l.cols = p;       p += align_up(R * maxc * sizeof(uint4), 256);
  l.colw = p;       p += det ? align_up(R * maxc * sizeof(uint4), 256) : 0;
  l.extra = p;      p += align_up(R * maxc * (det ? sizeof(uint4) : sizeof(uint2)), 256);
  const size_t tiles = (size_t)nbands * ntx;
  l.crange = p;     p += align_up(R * sizeof(int2), 256);
  l.visits = p;     p += align_up(R * tiles * sizeof(uint2), 256);
  l.vcount = p;     p += align_up(N * tiles * sizeof(int), 256);
  l.uwork = p;      p += align_up(N * tiles * sizeof(int), 256);
  l.uorder = p;     p += align_up(N * tiles * sizeof(int), 256);
  l.counter = p;    p += 256;
  l.total = p;
  return l;
}

struct FwdLayout {
  size_t img_start, order, entries, total;
};

FwdLayout fwd_layout(const RoiAlignArgs& a, size_t entry_bytes_per_roi) {
  FwdLayout l;
  size_t p = 0;
  l.img_start = p;  p += align_up((size_t)(a.num_images + 2) * sizeof(int), 256);
  l.order = p;      p += align_up((size_t)a.num_rois * sizeof(int), 256);
  l.entries = p;    p += align_up((size_t)a.num_rois * entry_bytes_per_roi, 256);
  l.total = p;
  return l;
}

}  // namespace

bool resident_supported(const RoiAlignArgs& a, const void* plane_ptr, int smem_optin, ResidentPlan* plan) {
  if (a.num_images < 0 || a.num_images > 4096) return false;
  if (a.sampling_ratio < 1 || a.sampling_ratio > 4) return false;
  const int per_roi = (a.pooled_h + a.pooled_w) * a.sampling_ratio;
  // row pitch: room for the repeated and the zero columns, rows 16-byte aligned for the bulk
  // copies, and 4 mod 8 floats so that consecutive rows start in different banks
  int pitch = (a.width + 3 + 3) / 4 * 4;
  if ((pitch / 4) % 2 == 0) pitch += 4;
  const size_t plane_bytes = (size_t)(a.height + 3) * pitch * sizeof(float);
  // the sampling_ratio == 2 kernel keeps its records in registers: no per-warp slots
  const bool two = a.sampling_ratio == 2 && a.pooled_w <= 16 && a.pooled_h <= 16;
  if (!two && per_roi > kMaxEntries) return false;
  const size_t fwd_smem = plane_bytes + (two ? 0 : (size_t)kFwdWarps * per_roi * sizeof(uint4));
  if (fwd_smem + 64 > (size_t)smem_optin) return false;
  plan->grid = a.sampling_ratio;
  plan->pitch = pitch;
  plan->per_roi = per_roi;
  plan->fwd_two = two ? 1 : 0;
  // channel-pair variant of the sampling_ratio == 2 forward: two padded planes must fit
  plan->fwd_pair = (two && 2 * plane_bytes + 64 <= (size_t)smem_optin) ? 1 : 0;
  plan->ypairs = 2 * ((a.pooled_h + 1) / 2);
  plan->xslots = 16 * ((2 * a.pooled_w + 15) / 16);
  plan->plane_bytes = plane_bytes;
  plan->fwd_smem = plan->fwd_pair ? 2 * plane_bytes : fwd_smem;
  plan->vec16 = (a.width % 4 == 0) && ((uintptr_t)plane_ptr % 16 == 0) ? 1 : 0;
  // backward: sampling_ratio 2, the dy tile of a (roi, channel) in a 256-float slot; the plane
  // is cut into as few slabs of rows as let two CTAs share an SM
  plan->bwd_ok = 0;
  if (a.sampling_ratio == 2 && a.pooled_w <= kMaxHitsPerCol && a.pooled_h <= 16 &&
      a.pooled_h * a.pooled_w <= kMaxSlotFloats && a.height < 65536 && a.width * 8 < 65536) {
    const int nt_raw = (2 * a.pooled_h * a.pooled_w + 31) / 32;
    const int nt = nt_raw <= 1 ? 1 : nt_raw <= 2 ? 2 : nt_raw <= 4 ? 4 : nt_raw <= 8 ? 8 : 16;
    int maxc = (4 * a.pooled_w + 31) / 32 * 32, maxe = (4 * a.pooled_h + 31) / 32 * 32;
    const bool small = maxc == 32 && maxe == 32;
    if (!small) maxc = maxe = 64;      // the kernel is instantiated for 32 / 32 and 64 / 64
    // per warp: dy tiles of the channel pair + the row entries of a visit
    // (sized for the 16-byte records of the deterministic variant)
    const size_t slot = (size_t)(nt * 32) * sizeof(float) + ((small ? 1 : 2) * 32 + 2) * sizeof(uint4);
    // band rows hold two channels (float2 per pixel); a warp instruction touches one row, so
    // the pitch only has to keep rows 16-byte aligned
    // every warp owns a private tile of the plane; two CTAs of kBwdWarps warps share an SM
    const size_t per_warp = (((size_t)smem_optin + 1024) / 2 - 1024 - 64) / kBwdWarps / 16 * 16;
    // A visit has a fixed cost (~45 % of the kernel's instructions at C2), and a ROI is visited
    // by every tile it meets: full-width bands of a wide plane are only a few rows high and a
    // typical ROI meets many of them.  Two column tiles of twice the height meet it less often
    // (C2: 5.9 -> 3.8 visits per ROI); with bands of 8 rows and more the row work dominates
    // and splitting the columns (= the lanes) only costs.
    int ntx = 1, tile_cols = (a.width + 3) & ~3;
    if (per_warp > slot && (per_warp - slot) / ((size_t)tile_cols * sizeof(float2)) < 8 && a.width >= 64) {
      ntx = 2;
      tile_cols = ((a.width + 1) / 2 + 3) & ~3;
    }
    const int pitch2 = tile_cols;
    const size_t row_bytes = (size_t)pitch2 * sizeof(float2);
    if (per_warp > slot + row_bytes && a.num_rois <= (1 << 20)) {
      int rows = (int)((per_warp - slot) / row_bytes);
      if (rows > a.height) rows = a.height;
      plan->bwd_ok = (a.height + rows - 1) / rows <= 4096 ? 1 : 0;   // band index: 12 bits of the sort key
      plan->bwd_slabs = (a.height + rows - 1) / rows;
      plan->bwd_slab_rows = rows;
      plan->bwd_smem = (size_t)kBwdWarps * (rows * row_bytes + slot);
      plan->bwd_pitch2 = pitch2;
      plan->bwd_ntx = ntx;
      plan->bwd_tile_cols = tile_cols;
      plan->bwd_maxc = maxc;
      plan->bwd_maxe = maxe;
    }
  }
  return true;
}

size_t resident_fwd_workspace(const RoiAlignArgs& a, const ResidentPlan& plan) {
  const size_t per_roi = plan.fwd_two ? (size_t)(2 * plan.ypairs + plan.xslots) * sizeof(uint2)
                                      : (size_t)plan.per_roi * sizeof(uint4);
  return fwd_layout(a, per_roi).total;
}

cudaError_t launch_fwd_resident(const RoiAlignArgs& a, const ResidentPlan& plan, void* workspace,
                                const float* data, float* out, int accumulate, int num_sms,
                                cudaStream_t stream) {
  if (a.num_rois == 0 || a.channels == 0) return cudaSuccess;
  const size_t per_roi_bytes = plan.fwd_two ? (size_t)(2 * plan.ypairs + plan.xslots) * sizeof(uint2)
                                            : (size_t)plan.per_roi * sizeof(uint4);
  const FwdLayout l = fwd_layout(a, per_roi_bytes);
  char* base = (char*)workspace;
  int* img_start = (int*)(base + l.img_start);
  int* order = (int*)(base + l.order);
  cudaError_t e = launch_roi_group(a, img_start, order, stream);
  if (e != cudaSuccess) return e;
  const int cc = a.c_count > 0 ? a.c_count : a.channels;      // channels worked on (stride: a.channels)
  const long long work = (long long)a.num_rois * cc;
  unsigned grid = 0;
  // dense plane + bulk copies: the box-head kernel (7 x 7, fp32, plain store) on planes whose rows are
  // 16-byte multiples; MOBULA_FWD_DENSE=0 / 1 overrides the default
  const int dense_env = [] { const char* e = getenv("MOBULA_FWD_DENSE"); return e ? atoi(e) : MOBULA_FWD_DENSE_DEFAULT; }();
  // rows per bulk copy (MOBULA_FWD_DENSE_ROWS)
  static const int dense_rows_env = [] { const char* e = getenv("MOBULA_FWD_DENSE_ROWS"); return e ? atoi(e) : 32; }();
  const int dense_rows = dense_rows_env > 0 ? dense_rows_env : 32;
  const size_t dense_smem = (size_t)(a.height + 3) * a.width * sizeof(float) + 64;
  const bool dense = dense_env != 0 && plan.fwd_two && !plan.fwd_pair && plan.vec16 && a.pooled_h == 7 &&
                     plan.xslots == 16 && !accumulate && a.top_dtype == 0 && dense_smem <= plan.fwd_smem + 64 &&
                     (size_t)(a.height + 1) * a.width * sizeof(float) < (1u << 20);   // one barrier phase counts < 2^20 bytes
  if (plan.fwd_two) {
    uint2* entries = (uint2*)(base + l.entries);
    const long long threads = (long long)a.num_rois * (2 * plan.ypairs + plan.xslots);
    chained_launch(resident_fwd2_table_kernel, (unsigned)((threads + 255) / 256), 256, 0, stream, 
        a.rois, order, img_start, a.num_images, a.height, a.width, dense ? a.width : plan.pitch, plan.fwd_pair ? 8 : 4,
        a.pooled_h, a.pooled_w, a.scale, plan.ypairs, plan.xslots, entries, dense ? 1 : 0);
    count_launch();
#define MOBULA_FWD2(ITERS, NCH, ACC, TYPED)                                                                 \
  do {                                                                                               \
    if (plan.fwd_pair) {                                                                             \
      e = resident_grid(resident_fwd2p_kernel<ITERS, NCH, ACC, TYPED>, kFwd2Threads, plan.fwd_smem, num_sms, \
                        (work + 1) / 2, &grid);                                                      \
      if (e != cudaSuccess) return e;                                                                \
      chained_launch(resident_fwd2p_kernel<ITERS, NCH, ACC, TYPED>, grid, kFwd2Threads, plan.fwd_smem, stream,          \
          data, out, order, img_start, entries, a.num_images, a.channels, cc, a.height, a.width,     \
          plan.pitch, a.pooled_h, a.pooled_w, a.top_dtype);                                          \
      break;                                                                                         \
    }                                                                                                \
    e = resident_grid(resident_fwd2_kernel<ITERS, NCH, ACC, TYPED>, kFwd2Threads, plan.fwd_smem, num_sms,   \
                      work, &grid);                                                                  \
    if (e != cudaSuccess) return e;                                                                  \
    chained_launch(resident_fwd2_kernel<ITERS, NCH, ACC, TYPED>, grid, kFwd2Threads, plan.fwd_smem, stream,             \
        data, out, order, img_start, entries, a.num_images, a.channels, cc, a.height, a.width,       \
        plan.pitch, a.pooled_h, a.pooled_w, plan.vec16, a.top_dtype);                             \
  } while (0)
#define MOBULA_FWD2_NCH(ITERS, NCH)                               \
  do {                                                            \
    if (accumulate) MOBULA_FWD2(ITERS, NCH, true, false);         \
    else if (a.top_dtype != 0) MOBULA_FWD2(ITERS, NCH, false, true); \
    else MOBULA_FWD2(ITERS, NCH, false, false);                   \
  } while (0)
#define MOBULA_FWD2_ITERS(ITERS)                             \
  do {                                                       \
    if (plan.xslots == 16) MOBULA_FWD2_NCH(ITERS, 1);        \
    else MOBULA_FWD2_NCH(ITERS, 2);                          \
  } while (0)
    const bool pairs_of_rois = (a.pooled_h & 1) && a.pooled_h <= 7;
#define MOBULA_FWD2R(PH, NCH, ACC, TYPED)                                                                    \
  do {                                                                                               \
    if (plan.fwd_pair) {                                                                             \
      e = resident_grid(resident_fwd2pr_kernel<PH, NCH, ACC, TYPED>, kFwd2Threads, plan.fwd_smem, num_sms,  \
                        (work + 1) / 2, &grid);                                                      \
      if (e != cudaSuccess) return e;                                                                \
      chained_launch(resident_fwd2pr_kernel<PH, NCH, ACC, TYPED>, grid, kFwd2Threads, plan.fwd_smem, stream, \
          data, out, order, img_start, entries, a.num_images, a.channels, cc, a.height, a.width,     \
          plan.pitch, a.pooled_h, a.pooled_w, a.top_dtype);                                          \
      break;                                                                                         \
    }                                                                                                \
    if (dense && PH == 7 && NCH == 1 && !ACC && !TYPED) {                                            \
      auto kd = resident_fwd2r_kernel<7, 1, false, false, true>;                                     \
      e = resident_grid(kd, kFwd2Threads, dense_smem, num_sms, work, &grid);                         \
      if (e != cudaSuccess) return e;                                                                \
      chained_launch(kd, grid, kFwd2Threads, dense_smem, stream, data, out, order, img_start, entries, \
          a.num_images, a.channels, cc, a.height, a.width, a.width, a.pooled_h, a.pooled_w, dense_rows, a.top_dtype); \
      break;                                                                                         \
    }                                                                                                \
    e = resident_grid(resident_fwd2r_kernel<PH, NCH, ACC, TYPED>, kFwd2Threads, plan.fwd_smem, num_sms, work, &grid); \
    if (e != cudaSuccess) return e;                                                                  \
    chained_launch(resident_fwd2r_kernel<PH, NCH, ACC, TYPED>, grid, kFwd2Threads, plan.fwd_smem, stream,   \
        data, out, order, img_start, entries, a.num_images, a.channels, cc, a.height, a.width,       \
        plan.pitch, a.pooled_h, a.pooled_w, plan.vec16, a.top_dtype);                                \
  } while (0)
#define MOBULA_FWD2R_PH(PH)                                          \
  do {                                                               \
    if (plan.xslots == 16) {                                         \
      if (accumulate) MOBULA_FWD2R(PH, 1, true, false);              \
      else if (a.top_dtype != 0) MOBULA_FWD2R(PH, 1, false, true);   \
      else MOBULA_FWD2R(PH, 1, false, false);                        \
    } else {                                                         \
      if (accumulate) MOBULA_FWD2R(PH, 2, true, false);              \
      else if (a.top_dtype != 0) MOBULA_FWD2R(PH, 2, false, true);   \
      else MOBULA_FWD2R(PH, 2, false, false);                        \
    }                                                                \
  } while (0)
    if (pairs_of_rois) {
      switch (a.pooled_h) {
        case 1: MOBULA_FWD2R_PH(1); break;
        case 3: MOBULA_FWD2R_PH(3); break;
        case 5: MOBULA_FWD2R_PH(5); break;
        default: MOBULA_FWD2R_PH(7); break;
      }
    } else
    switch (plan.ypairs / 2) {
      case 1: MOBULA_FWD2_ITERS(1); break;
      case 2: MOBULA_FWD2_ITERS(2); break;
      case 3: MOBULA_FWD2_ITERS(3); break;
      case 4: MOBULA_FWD2_ITERS(4); break;
      case 5: MOBULA_FWD2_ITERS(5); break;
      case 6: MOBULA_FWD2_ITERS(6); break;
      case 7: MOBULA_FWD2_ITERS(7); break;
      default: MOBULA_FWD2_ITERS(8); break;
    }
#undef MOBULA_FWD2R_PH
#undef MOBULA_FWD2R
#undef MOBULA_FWD2_ITERS
#undef MOBULA_FWD2_NCH
#undef MOBULA_FWD2
    count_launch();
    return cudaGetLastError();
  }
  uint4* entries = (uint4*)(base + l.entries);
  const long long threads = (long long)a.num_rois * plan.per_roi;
  chained_launch(resident_fwd_table_kernel, (unsigned)((threads + 255) / 256), 256, 0, stream, 
      a.rois, order, img_start, a.num_images, a.height, a.width, plan.pitch, a.pooled_h, a.pooled_w,
      a.scale, plan.grid, entries);
  count_launch();
#define MOBULA_RESIDENT_FWD(G)                                                                      \
  do {                                                                                              \
    e = resident_grid(resident_fwd_kernel<G>, kFwdThreads, plan.fwd_smem, num_sms, work, &grid);    \
    if (e != cudaSuccess) return e;                                                                 \
    chained_launch(resident_fwd_kernel<G>, grid, kFwdThreads, plan.fwd_smem, stream,                            \
        data, out, order, img_start, entries, a.num_images, a.channels, cc, a.height, a.width,      \
        plan.pitch, a.pooled_h, a.pooled_w, accumulate, plan.vec16);                             \
  } while (0)
  switch (plan.grid) {
    case 1: MOBULA_RESIDENT_FWD(1); break;
    case 2: MOBULA_RESIDENT_FWD(2); break;
    case 3: MOBULA_RESIDENT_FWD(3); break;
    default: MOBULA_RESIDENT_FWD(4); break;
  }
#undef MOBULA_RESIDENT_FWD
  count_launch();
  return cudaGetLastError();
}

size_t resident_bwd_workspace(const RoiAlignArgs& a, const ResidentPlan& plan, int deterministic) {
  return bwd_layout(a, plan.bwd_maxe, plan.bwd_maxc, plan.bwd_slabs, plan.bwd_ntx, deterministic != 0).total;
}

cudaError_t launch_bwd_resident(const RoiAlignArgs& a, const ResidentPlan& plan, void* workspace,
                                const float* dy, float* dx, int accumulate, int deterministic, int num_sms,
                                cudaStream_t stream) {
  if (a.num_images == 0 || a.channels == 0) return cudaSuccess;
  const int nbands = plan.bwd_slabs;
  const bool det = deterministic != 0;
  const int ntx = plan.bwd_ntx;
  const BwdLayout l = bwd_layout(a, plan.bwd_maxe, plan.bwd_maxc, nbands, ntx, det);
  char* base = (char*)workspace;
  int* img_start = (int*)(base + l.img_start);
  int* order = (int*)(base + l.order);
  void* rows = base + l.rows;
  unsigned char* bstart = (unsigned char*)(base + l.bstart);
  uint4* cols = (uint4*)(base + l.cols);
  uint4* colw = (uint4*)(base + l.colw);
  void* extra = base + l.extra;
  int2* crange = (int2*)(base + l.crange);
  uint2* visits = (uint2*)(base + l.visits);
  int* vcount = (int*)(base + l.vcount);
  int* uwork = (int*)(base + l.uwork);
  int* uorder = (int*)(base + l.uorder);
  int* counter = (int*)(base + l.counter);
  cudaError_t e = cudaSuccess;
  // grouping blocks (one per image + one for the rows with an invalid batch index), then four ROIs per block
  const int tb = a.num_images + 1 + (a.num_rois + 3) / 4;
  if (det)
    chained_launch(resident_bwd2d_table_kernel, tb, 128, 0, stream, 
        a.rois, order, img_start, a.num_images, a.num_rois, a.height, a.width, plan.bwd_pitch2 * 8, 8,
        a.pooled_h, a.pooled_w, a.scale, plan.bwd_maxe, plan.bwd_maxc, plan.bwd_slab_rows, nbands, (uint4*)rows,
        bstart, cols, colw, (uint4*)extra, crange, counter);
  else
    chained_launch(resident_bwd2_table_kernel, tb, 256, 0, stream, 
        a.rois, order, img_start, a.num_images, a.num_rois, a.height, a.width, plan.bwd_pitch2 * 8, 8,
        a.pooled_h, a.pooled_w, a.scale, plan.bwd_maxe, plan.bwd_maxc, plan.bwd_slab_rows, nbands, (uint4*)rows,
        bstart, cols, (uint2*)extra, crange, counter);
  count_launch();
  chained_launch(resident_bwd2_visits_kernel, a.num_images * nbands * ntx, kVisitThreads, 0, stream, 
      order, img_start, bstart, crange, a.num_images, a.num_rois, nbands, ntx, plan.bwd_tile_cols, visits, vcount,
      uwork, uorder, counter);
  count_launch();
  const int cc = a.c_count > 0 ? a.c_count : a.channels;      // channels worked on (stride: a.channels)
  const long long units = (long long)a.num_images * ((cc + 1) / 2) * nbands * ntx;
  const long long cta_units = (units + kBwdWarps - 1) / kBwdWarps;
  unsigned grid = 0;
#define MOBULA_BWD2(NT, NCK, NE, ACC, DET)                                                            \
  do {                                                                                                \
    e = resident_grid(resident_bwd2_kernel<NT, NCK, NE, ACC, DET>, kBwdThreads, plan.bwd_smem,        \
                      num_sms, cta_units, &grid);                                                     \
    if (e != cudaSuccess) return e;                                                                   \
    chained_launch(resident_bwd2_kernel<NT, NCK, NE, ACC, DET>, grid, kBwdThreads, plan.bwd_smem, stream,         \
        dy, dx, visits, vcount, uorder, rows, cols, colw, extra, img_start, counter, a.num_images,    \
        a.channels, cc, a.height, a.width, plan.bwd_pitch2, a.pooled_h, a.pooled_w, plan.bwd_maxe,    \
        plan.bwd_maxc, nbands, plan.bwd_slab_rows, ntx, plan.bwd_tile_cols);                          \
  } while (0)
#define MOBULA_BWD2_ACC(NT, NCK, NE)                                       \
  do {                                                                     \
    if (det) {                                                             \
      if (accumulate) MOBULA_BWD2(NT, NCK, NE, true, true);                \
      else MOBULA_BWD2(NT, NCK, NE, false, true);                          \
    } else {                                                               \
      if (accumulate) MOBULA_BWD2(NT, NCK, NE, true, false);               \
      else MOBULA_BWD2(NT, NCK, NE, false, false);                         \
    }                                                                      \
  } while (0)
#define MOBULA_BWD2_NT(NT)                                                  \
  do {                                                                      \
    if (plan.bwd_maxc == 32 && plan.bwd_maxe == 32) MOBULA_BWD2_ACC(NT, 1, 1); \
    else MOBULA_BWD2_ACC(NT, 2, 2);                                         \
  } while (0)
  const int nt = (2 * a.pooled_h * a.pooled_w + 31) / 32;
  if (nt <= 1) MOBULA_BWD2_NT(1);
  else if (nt <= 2) MOBULA_BWD2_NT(2);
  else if (nt <= 4) MOBULA_BWD2_NT(4);
  else if (nt <= 8) MOBULA_BWD2_NT(8);
  else MOBULA_BWD2_NT(16);
#undef MOBULA_BWD2_NT
#undef MOBULA_BWD2_ACC
#undef MOBULA_BWD2
  count_launch();
  return cudaGetLastError();
}

}  // namespace mobula_b200
