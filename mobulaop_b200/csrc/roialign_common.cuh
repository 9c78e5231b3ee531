// ROI / sample decoding shared by every ROIAlign kernel.
//
// The integer results of this arithmetic (grid sizes, tap indices, the
// inside/outside decision) are discontinuous in the inputs, so it has to round
// exactly like the reference's CPU build, which has no FMA contraction:
// every operation below is an explicit round-to-nearest intrinsic and the
// translation unit is compiled with -fmad=false as a second line of defence.
//
// Reference arithmetic: /root/reference/opzoo/ROIAlign/ROIAlign.cpp:24-65 and
// /root/reference/opzoo/ROIAlign/bilinear.h:15-54.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace mobula_b200 {

constexpr float kMaxAdaptiveGrid = 1024.0f;

struct RoiGeom {
  int batch;            // ROIAlign.cpp:25
  float start_w, start_h;
  float bin_w, bin_h;   // ROIAlign.cpp:40-41
  int grid_h, grid_w;   // ROIAlign.cpp:47-51
  float count;          // ROIAlign.cpp:54
};

// aligned = the half-pixel variant of torchvision / detectron2 (roi_align(aligned=True)): box
// coordinates are shifted by -0.5 after scaling and the box is NOT forced to be at least one
// pixel wide (torchvision/csrc/ops/cpu/roi_align_kernel.cpp); false = the reference operator.
__device__ __forceinline__ RoiGeom roi_decode(const float* __restrict__ roi, float scale,
                                              int pooled_h, int pooled_w, int sampling_ratio,
                                              bool aligned = false) {
  RoiGeom g;
  const float r0 = __ldg(roi + 0), r1 = __ldg(roi + 1), r2 = __ldg(roi + 2);
  const float r3 = __ldg(roi + 3), r4 = __ldg(roi + 4);
  g.batch = __float2int_rz(r0);
  g.start_w = __fmul_rn(r1, scale);
  g.start_h = __fmul_rn(r2, scale);
  float end_w = __fmul_rn(r3, scale);
  float end_h = __fmul_rn(r4, scale);
  float roi_w, roi_h;
  if (aligned) {
    g.start_w = __fsub_rn(g.start_w, 0.5f);
    g.start_h = __fsub_rn(g.start_h, 0.5f);
    end_w = __fsub_rn(end_w, 0.5f);
    end_h = __fsub_rn(end_h, 0.5f);
    roi_w = __fsub_rn(end_w, g.start_w);
    roi_h = __fsub_rn(end_h, g.start_h);
  } else {
    roi_w = fmaxf(__fsub_rn(end_w, g.start_w), 1.0f);
    roi_h = fmaxf(__fsub_rn(end_h, g.start_h), 1.0f);
  }
  g.bin_h = __fdiv_rn(roi_h, (float)pooled_h);
  g.bin_w = __fdiv_rn(roi_w, (float)pooled_w);
  float gh = (float)sampling_ratio, gw = (float)sampling_ratio;
  if (sampling_ratio <= 0) {
    gh = ceilf(g.bin_h);
    gw = ceilf(g.bin_w);
  }
  // A ROI with a non-finite coordinate, or an adaptive grid beyond kMaxAdaptiveGrid samples per
  // bin axis (a box thousands of pixels outside the map), would overflow the int grid sizes and
  // keep a thread looping for minutes; the reference itself is undefined there ((int)inf,
  // ROIAlign.cpp:47-51).  Such a ROI is moved outside the plane: every sample is rejected
  // (bilinear.h:15-18), its output rows are zero and it scatters nothing.
  const float probe = (g.start_w + g.start_h) + (g.bin_w + g.bin_h);
  // (aligned boxes may be empty: an adaptive grid of zero samples pools to 0 / max(count, 1) = 0)
  // (an aligned box with x2 < x1 or y2 < y1 would be sampled backwards by torchvision; here it pools to 0)
  if (!(gh <= kMaxAdaptiveGrid && gw <= kMaxAdaptiveGrid) || !(fabsf(probe) <= 3.0e38f) || gh < 1.0f ||
      gw < 1.0f || roi_w < 0.0f || roi_h < 0.0f) {
    g.start_w = g.start_h = -1.0e30f;
    g.bin_w = g.bin_h = 0.0f;
    gh = gw = 1.0f;
  }
  g.grid_h = __float2int_rz(gh);
  g.grid_w = __float2int_rz(gw);
  g.count = (float)(g.grid_h * g.grid_w);
  return g;
}

// (start + p*bin) + (((i + .5) * bin) / grid), ROIAlign.cpp:59-65
__device__ __forceinline__ float sample_pos(float start, int p, float bin, int i, int grid) {
  const float base = __fadd_rn(start, __fmul_rn((float)p, bin));
  const float off = __fdiv_rn(__fmul_rn((float)i + 0.5f, bin), (float)grid);
  return __fadd_rn(base, off);
}

struct AxisTap {
  int lo, hi;     // pixel indices
  float wl, wh;   // weight of hi (fraction) and of lo (1 - fraction)
  bool valid;     // false: bilinear.h:15-18 rejects the sample
};

__device__ __forceinline__ AxisTap axis_decode(float pos, int extent) {
  AxisTap t;
  t.valid = !(pos < -1.0f || pos > (float)extent);
  if (pos <= 0.0f) pos = 0.0f;
  t.lo = __float2int_rz(pos);
  if (t.lo >= extent - 1) {
    t.hi = t.lo = extent - 1;
    pos = (float)t.lo;
  } else {
    t.hi = t.lo + 1;
  }
  t.wl = __fsub_rn(pos, (float)t.lo);
  t.wh = __fsub_rn(1.0f, t.wl);
  return t;
}

// w1*v1 + w2*v2 + w3*v3 + w4*v4, left to right, no contraction (bilinear.h:56)
__device__ __forceinline__ float tap_sum(float w1, float v1, float w2, float v2, float w3,
                                         float v3, float w4, float v4) {
  float s = __fadd_rn(__fmul_rn(w1, v1), __fmul_rn(w2, v2));
  s = __fadd_rn(s, __fmul_rn(w3, v3));
  return __fadd_rn(s, __fmul_rn(w4, v4));
}

// (dtop * w) / count, ROIAlign.cpp:143-146.  When count is a power of two the
// division is an exact scaling and equals multiplication by the exact reciprocal.
struct CountDiv {
  float count, inv;
  bool pow2;
  __device__ __forceinline__ explicit CountDiv(int n) {
    count = (float)n;
    pow2 = n > 0 && (n & (n - 1)) == 0;
    inv = pow2 ? __fdiv_rn(1.0f, count) : 0.0f;
  }
  __device__ __forceinline__ float operator()(float x) const {
    return pow2 ? __fmul_rn(x, inv) : __fdiv_rn(x, count);
  }
};

// ---- element type of the pooled tensor (forward output / backward dy) ---------------------------
// 0 = fp32 (the reference), 1 = fp16, 2 = bf16.  The arithmetic is fp32 whatever
// the storage type: a result is rounded once (to nearest even) when it is stored, a gradient is
// widened exactly when it is loaded.
constexpr int kTopF32 = 0, kTopF16 = 1, kTopBF16 = 2;

__device__ __forceinline__ void store_top(void* base, size_t idx, float v, int dtype) {
  if (dtype == kTopF32) static_cast<float*>(base)[idx] = v;
  else if (dtype == kTopF16) static_cast<__half*>(base)[idx] = __float2half_rn(v);
  else static_cast<__nv_bfloat16*>(base)[idx] = __float2bfloat16_rn(v);
}

__device__ __forceinline__ float load_top(const void* base, size_t idx, int dtype) {
  if (dtype == kTopF32) return __ldg(static_cast<const float*>(base) + idx);
  if (dtype == kTopF16) return __half2float(__ldg(static_cast<const __half*>(base) + idx));
  return __bfloat162float(__ldg(static_cast<const __nv_bfloat16*>(base) + idx));
}

__host__ __device__ inline size_t top_elem_bytes(int dtype) { return dtype == kTopF32 ? 4 : 2; }

// ---- ROI rows grouped by image ----------------------------------------------------------------
// image of ROI row r, or num_images for a batch index outside [0, N) (such rows read nothing)
__device__ __forceinline__ int segment_of(const float* __restrict__ rois, int r, int num_images) {
  const int b = __float2int_rz(__ldg(rois + (size_t)r * 5));
  return (b >= 0 && b < num_images) ? b : num_images;
}

// block-wide exclusive prefix sum of one value per thread; returns the block total too
template <int THREADS, typename T>
__device__ __forceinline__ T block_exclusive_scan(T v, T* total, T* warp_sums /* [THREADS/32] */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  T incl = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const T n = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += n;
  }
  if (lane == 31) warp_sums[warp] = incl;
  __syncthreads();
  T warp_prefix = 0, all = 0;
  for (int w = 0; w < THREADS / 32; ++w) {
    const T s = warp_sums[w];
    if (w < warp) warp_prefix += s;
    all += s;
  }
  __syncthreads();
  *total = all;
  return warp_prefix + incl - v;
}

// One block per segment (image 0 .. N-1, then the rows with an invalid batch index): the stable
// partition of the ROI rows by image.  order[img_start[n] ..] = rows of image n, ascending (the
// canonical accumulation order of the backward); img_start[N + 1] = num_rois.
template <int THREADS>
__device__ __forceinline__ void roi_group_block(const float* __restrict__ rois, int num_rois, int num_images,
                                                int seg, int* __restrict__ img_start, int* __restrict__ order,
                                                int* warp_sums /* [THREADS / 32] */) {
  // one pass: rows before this segment, rows of this segment, and whether the table is already
  // grouped (batch indices never decrease -- what every detection pipeline emits)
  int before = 0, mine_n = 0, unsorted = 0;
  for (int r = threadIdx.x; r < num_rois; r += THREADS) {
    const int sg = segment_of(rois, r, num_images);
    before += sg < seg ? 1 : 0;
    mine_n += sg == seg ? 1 : 0;
    if (r > 0) unsorted |= segment_of(rois, r - 1, num_images) > sg ? 1 : 0;
  }
  int start, count;
  block_exclusive_scan<THREADS>(before, &start, warp_sums);
  block_exclusive_scan<THREADS>(mine_n, &count, warp_sums);
  if (threadIdx.x == 0) {
    img_start[seg] = start;
    if (seg == num_images) img_start[seg + 1] = num_rois;
  }
  if (!__syncthreads_or(unsorted)) {
    // grouped already: the rows of this segment are [start, start + count), in place
    for (int i = threadIdx.x; i < count; i += THREADS) order[start + i] = start + i;
    return;
  }
  // any order: stable partition, THREADS rows per block-wide scan
  int running = start;
  for (int base = 0; base < num_rois; base += THREADS) {
    const int r = base + threadIdx.x;
    const int mine = (r < num_rois && segment_of(rois, r, num_images) == seg) ? 1 : 0;
    int total;
    const int pos = block_exclusive_scan<THREADS>(mine, &total, warp_sums);
    if (mine) order[running + pos] = r;
    running += total;
  }
}

}  // namespace mobula_b200
