// "gather" kernels: the shape-agnostic ROIAlign path.
//
// One thread owns one (roi, ph, pw) bin for a group of CH channels.  The ROI and
// sample decoding is channel independent (the reference recomputes it for every one
// of the C*PH*PW outputs of a ROI, ROIAlign.cpp:24-54), so it is done once per sample
// and applied to CH channel planes whose 4*CH tap loads are independent and in flight
// together.  Adjacent lanes hold adjacent pw, so the output row of a (roi, channel) is
// written as one contiguous run and the taps of a lane's neighbours sit in the same or
// neighbouring 32-byte sectors.
//
// Per-output arithmetic and order are the reference's (sum over iy then ix of
// w1*v1+w2*v2+w3*v3+w4*v4, then / count), so the forward is bit-identical to it.
#include "roialign_kernels.h"
#include "roialign_common.cuh"

namespace mobula_b200 {

template <int CH>
__global__ void __launch_bounds__(256)
roialign_fwd_gather_kernel(const float* __restrict__ data, const float* __restrict__ rois,
                           float* __restrict__ out, int num_rois, int num_images, int channels,
                           int height, int width, int pooled_h, int pooled_w, float scale,
                           int sampling_ratio, int accumulate) {
  const int bins = pooled_h * pooled_w;
  const long long item = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (item >= (long long)num_rois * bins) return;
  const int r = (int)(item / bins);
  const int bin = (int)(item - (long long)r * bins);
  const int ph = bin / pooled_w, pw = bin - ph * pooled_w;
  const int c0 = blockIdx.y * CH;

  const RoiGeom g = roi_decode(rois + (size_t)r * 5, scale, pooled_h, pooled_w, sampling_ratio);
  const size_t plane_sz = (size_t)height * width;
  const bool image_ok = num_images < 0 || (g.batch >= 0 && g.batch < num_images);

  float acc[CH];
#pragma unroll
  for (int k = 0; k < CH; ++k) acc[k] = 0.0f;

  if (image_ok) {
    const float* base = data + ((size_t)g.batch * channels + c0) * plane_sz;
    for (int iy = 0; iy < g.grid_h; ++iy) {
      const AxisTap ty = axis_decode(sample_pos(g.start_h, ph, g.bin_h, iy, g.grid_h), height);
      if (!ty.valid) continue;   // the reference adds +0.0f, which cannot change the sum
      for (int ix = 0; ix < g.grid_w; ++ix) {
        const AxisTap tx = axis_decode(sample_pos(g.start_w, pw, g.bin_w, ix, g.grid_w), width);
        if (!tx.valid) continue;
        const float w1 = __fmul_rn(ty.wh, tx.wh), w2 = __fmul_rn(ty.wh, tx.wl);
        const float w3 = __fmul_rn(ty.wl, tx.wh), w4 = __fmul_rn(ty.wl, tx.wl);
        const int o1 = ty.lo * width + tx.lo, o2 = ty.lo * width + tx.hi;
        const int o3 = ty.hi * width + tx.lo, o4 = ty.hi * width + tx.hi;
        float v1[CH], v2[CH], v3[CH], v4[CH];
#pragma unroll
        for (int k = 0; k < CH; ++k) {
          if (c0 + k < channels) {
            const float* p = base + (size_t)k * plane_sz;
            v1[k] = __ldg(p + o1);
            v2[k] = __ldg(p + o2);
            v3[k] = __ldg(p + o3);
            v4[k] = __ldg(p + o4);
          } else {
            v1[k] = v2[k] = v3[k] = v4[k] = 0.0f;
          }
        }
#pragma unroll
        for (int k = 0; k < CH; ++k)
          acc[k] = __fadd_rn(acc[k], tap_sum(w1, v1[k], w2, v2[k], w3, v3[k], w4, v4[k]));
      }
    }
  }
  float* o = out + ((size_t)r * channels + c0) * bins + bin;
#pragma unroll
  for (int k = 0; k < CH; ++k) {
    if (c0 + k < channels) {
      const float y = __fdiv_rn(acc[k], g.count);
      if (accumulate) o[(size_t)k * bins] = __fadd_rn(o[(size_t)k * bins], y);
      else o[(size_t)k * bins] = y;
    }
  }
}

// Backward scatter with global fp32 reductions (RED.ADD.F32).  bottom_diff must
// already hold the values to add to (zeros for a plain backward).
template <int CH>
__global__ void __launch_bounds__(256)
roialign_bwd_gather_kernel(const float* __restrict__ dy, const float* __restrict__ rois,
                           float* __restrict__ dx, int num_rois, int num_images, int channels,
                           int height, int width, int pooled_h, int pooled_w, float scale,
                           int sampling_ratio) {
  const int bins = pooled_h * pooled_w;
  const long long item = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (item >= (long long)num_rois * bins) return;
  const int r = (int)(item / bins);
  const int bin = (int)(item - (long long)r * bins);
  const int ph = bin / pooled_w, pw = bin - ph * pooled_w;
  const int c0 = blockIdx.y * CH;

  const RoiGeom g = roi_decode(rois + (size_t)r * 5, scale, pooled_h, pooled_w, sampling_ratio);
  if (num_images >= 0 && (g.batch < 0 || g.batch >= num_images)) return;
  const size_t plane_sz = (size_t)height * width;
  const CountDiv div(g.grid_h * g.grid_w);

  float dtop[CH];
  const float* t = dy + ((size_t)r * channels + c0) * bins + bin;
#pragma unroll
  for (int k = 0; k < CH; ++k) dtop[k] = (c0 + k < channels) ? __ldg(t + (size_t)k * bins) : 0.0f;

  float* base = dx + ((size_t)g.batch * channels + c0) * plane_sz;
  for (int iy = 0; iy < g.grid_h; ++iy) {
    const AxisTap ty = axis_decode(sample_pos(g.start_h, ph, g.bin_h, iy, g.grid_h), height);
    if (!ty.valid) continue;
    for (int ix = 0; ix < g.grid_w; ++ix) {
      const AxisTap tx = axis_decode(sample_pos(g.start_w, pw, g.bin_w, ix, g.grid_w), width);
      if (!tx.valid) continue;
      const float w1 = __fmul_rn(ty.wh, tx.wh), w2 = __fmul_rn(ty.wh, tx.wl);
      const float w3 = __fmul_rn(ty.wl, tx.wh), w4 = __fmul_rn(ty.wl, tx.wl);
      const int o1 = ty.lo * width + tx.lo, o2 = ty.lo * width + tx.hi;
      const int o3 = ty.hi * width + tx.lo, o4 = ty.hi * width + tx.hi;
#pragma unroll
      for (int k = 0; k < CH; ++k) {
        if (c0 + k < channels) {
          float* p = base + (size_t)k * plane_sz;
          atomicAdd(p + o1, div(__fmul_rn(dtop[k], w1)));
          atomicAdd(p + o2, div(__fmul_rn(dtop[k], w2)));
          atomicAdd(p + o3, div(__fmul_rn(dtop[k], w3)));
          atomicAdd(p + o4, div(__fmul_rn(dtop[k], w4)));
        }
      }
    }
  }
}

// Writes the decoded ROI bookkeeping (see mobula_roialign.h); one thread per ROI.
__global__ void roialign_bookkeeping_kernel(const float* __restrict__ rois, int num_rois, int height,
                                            int width, int pooled_h, int pooled_w, float scale,
                                            int sampling_ratio, int32_t* geom_i, float* geom_f,
                                            int max_grid, int32_t* tap_i, float* tap_f) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= num_rois) return;
  const RoiGeom g = roi_decode(rois + (size_t)r * 5, scale, pooled_h, pooled_w, sampling_ratio);
  geom_i[r * 4 + 0] = g.batch;
  geom_i[r * 4 + 1] = g.grid_h;
  geom_i[r * 4 + 2] = g.grid_w;
  geom_i[r * 4 + 3] = 0;
  geom_f[r * 5 + 0] = g.start_w;
  geom_f[r * 5 + 1] = g.start_h;
  geom_f[r * 5 + 2] = g.bin_w;
  geom_f[r * 5 + 3] = g.bin_h;
  geom_f[r * 5 + 4] = g.count;
  if (max_grid <= 0) return;
  const int pooled_max = pooled_h > pooled_w ? pooled_h : pooled_w;
  for (int axis = 0; axis < 2; ++axis) {
    const int pooled = axis == 0 ? pooled_h : pooled_w;
    const int grid = axis == 0 ? g.grid_h : g.grid_w;
    const int extent = axis == 0 ? height : width;
    const float start = axis == 0 ? g.start_h : g.start_w;
    const float bin = axis == 0 ? g.bin_h : g.bin_w;
    for (int p = 0; p < pooled; ++p) {
      for (int i = 0; i < grid && i < max_grid; ++i) {
        const size_t e = ((((size_t)r * 2 + axis) * pooled_max + p) * max_grid + i) * 2;
        const float pos = sample_pos(start, p, bin, i, grid);
        const AxisTap t = axis_decode(pos, extent);
        tap_f[e + 0] = pos;
        tap_i[e + 0] = t.valid ? t.lo : -1;
        tap_i[e + 1] = t.valid ? t.hi : -1;
        tap_f[e + 1] = t.valid ? t.wl : 0.0f;
      }
    }
  }
}

// ---------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------
static inline int pick_channel_group(int channels) { return channels >= 8 ? 8 : (channels >= 4 ? 4 : 1); }

cudaError_t launch_fwd_gather(const RoiAlignArgs& a, const float* data, float* out, int accumulate,
                              cudaStream_t stream) {
  const long long items = (long long)a.num_rois * a.pooled_h * a.pooled_w;
  if (items == 0 || a.channels == 0) return cudaSuccess;
  const int ch = pick_channel_group(a.channels);
  dim3 block(256), grid((unsigned)((items + 255) / 256), (unsigned)((a.channels + ch - 1) / ch));
#define MOBULA_LAUNCH(CH)                                                                        \
  roialign_fwd_gather_kernel<CH><<<grid, block, 0, stream>>>(                                    \
      data, a.rois, out, a.num_rois, a.num_images, a.channels, a.height, a.width, a.pooled_h,   \
      a.pooled_w, a.scale, a.sampling_ratio, accumulate)
  if (ch == 8) MOBULA_LAUNCH(8); else if (ch == 4) MOBULA_LAUNCH(4); else MOBULA_LAUNCH(1);
#undef MOBULA_LAUNCH
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_bwd_gather(const RoiAlignArgs& a, const float* dy, float* dx, cudaStream_t stream) {
  const long long items = (long long)a.num_rois * a.pooled_h * a.pooled_w;
  if (items == 0 || a.channels == 0) return cudaSuccess;
  const int ch = pick_channel_group(a.channels);
  dim3 block(256), grid((unsigned)((items + 255) / 256), (unsigned)((a.channels + ch - 1) / ch));
#define MOBULA_LAUNCH(CH)                                                                        \
  roialign_bwd_gather_kernel<CH><<<grid, block, 0, stream>>>(                                    \
      dy, a.rois, dx, a.num_rois, a.num_images, a.channels, a.height, a.width, a.pooled_h,      \
      a.pooled_w, a.scale, a.sampling_ratio)
  if (ch == 8) MOBULA_LAUNCH(8); else if (ch == 4) MOBULA_LAUNCH(4); else MOBULA_LAUNCH(1);
#undef MOBULA_LAUNCH
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_bookkeeping(const RoiAlignArgs& a, int32_t* geom_i, float* geom_f, int max_grid,
                               int32_t* tap_i, float* tap_f, cudaStream_t stream) {
  if (a.num_rois == 0) return cudaSuccess;
  roialign_bookkeeping_kernel<<<(a.num_rois + 127) / 128, 128, 0, stream>>>(
      a.rois, a.num_rois, a.height, a.width, a.pooled_h, a.pooled_w, a.scale, a.sampling_ratio,
      geom_i, geom_f, max_grid, tap_i, tap_f);
  count_launch();
  return cudaGetLastError();
}

}  // namespace mobula_b200
