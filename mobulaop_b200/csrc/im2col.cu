// im2col / col2im of the opzoo Convolution operator, through the same prebuilt
// library and the same drop-in mechanism as ROIAlign.
//
// Reference: /root/reference/opzoo/Convolution/Convolution.cpp:14-45 (im2col_kernel: one thread
// per element of the column buffer [C*KH*KW, OH, OW]) and :48-92 (col2im_kernel: one thread per
// image pixel gathers the column entries that map to it, in ascending (h_col, w_col) order -- a
// gather, hence deterministic).  Both are HBM-bound index shuffles; here:
//   * im2col: a thread produces FOUR consecutive output columns of one (channel, kernel tap,
//     output row) with one 16-byte store when the row allows it; reads are coalesced along the
//     image row for stride 1;
//   * col2im: same gather and the same summation order as the reference (bit-identical sums),
//     with the divisions of the index decode hoisted out of the tap loops.
#include <cuda_runtime.h>

#include "../../include/mobula_roialign.h"
#include "roialign_kernels.h"

namespace mobula_b200 {

namespace {

struct ConvGeom {
  int height, width, kernel_h, kernel_w, pad_h, pad_w, stride_h, stride_w, dilation_h, dilation_w;
  int height_col, width_col;
};

// column buffer element (c, kh, kw, oh, ow) = image (c, oh*sh - ph + kh*dh, ow*sw - pw + kw*dw) or 0
__global__ void __launch_bounds__(256) im2col_kernel(const long long groups, const float* __restrict__ im,
                                                     const ConvGeom g, float* __restrict__ col) {
  // one thread per group of four consecutive output columns
  const int wq = (g.width_col + 3) >> 2;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < groups;
       t += (long long)gridDim.x * blockDim.x) {
    const int q = (int)(t % wq);
    long long rest = t / wq;
    const int oh = (int)(rest % g.height_col);
    rest /= g.height_col;
    const int kw = (int)(rest % g.kernel_w);
    rest /= g.kernel_w;
    const int kh = (int)(rest % g.kernel_h);
    const int c = (int)(rest / g.kernel_h);
    const int ih = oh * g.stride_h - g.pad_h + kh * g.dilation_h;
    const bool row_ok = (unsigned)ih < (unsigned)g.height;
    const float* src = im + ((size_t)c * g.height + (row_ok ? ih : 0)) * g.width;
    const int ow0 = q * 4;
    float v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int iw = (ow0 + k) * g.stride_w - g.pad_w + kw * g.dilation_w;
      v[k] = (row_ok && (unsigned)iw < (unsigned)g.width) ? __ldg(src + iw) : 0.0f;
    }
    float* dst = col + ((((size_t)c * g.kernel_h + kh) * g.kernel_w + kw) * g.height_col + oh) * g.width_col + ow0;
    if (ow0 + 3 < g.width_col && ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0)) {
      *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (ow0 + k < g.width_col) dst[k] = v[k];
    }
  }
}

// image pixel (c, h, w) = sum over the column entries that were copied from it, ascending
// (h_col, w_col) as in Convolution.cpp:67-83
__global__ void __launch_bounds__(256) col2im_kernel(const long long n, const float* __restrict__ col,
                                                     const ConvGeom g, float* __restrict__ im) {
  const int ext_w = (g.kernel_w - 1) * g.dilation_w + 1, ext_h = (g.kernel_h - 1) * g.dilation_h + 1;
  for (long long index = (long long)blockIdx.x * blockDim.x + threadIdx.x; index < n;
       index += (long long)gridDim.x * blockDim.x) {
    const int w_im = (int)(index % g.width) + g.pad_w;
    const int h_im = (int)((index / g.width) % g.height) + g.pad_h;
    const int c_im = (int)(index / ((long long)g.width * g.height));
    const int w0 = w_im < ext_w ? 0 : (w_im - ext_w) / g.stride_w + 1;
    const int w1 = min(w_im / g.stride_w + 1, g.width_col);
    const int h0 = h_im < ext_h ? 0 : (h_im - ext_h) / g.stride_h + 1;
    const int h1 = min(h_im / g.stride_h + 1, g.height_col);
    float val = 0.0f;
    for (int h_col = h0; h_col < h1; ++h_col) {
      const int hk = h_im - h_col * g.stride_h;
      if (hk % g.dilation_h) continue;
      const int kh = hk / g.dilation_h;
      const float* base = col + ((size_t)(c_im * g.kernel_h + kh) * g.kernel_w) * g.height_col * g.width_col +
                          (size_t)h_col * g.width_col;
      for (int w_col = w0; w_col < w1; ++w_col) {
        const int wk = w_im - w_col * g.stride_w;
        if (wk % g.dilation_w) continue;
        const int kw = wk / g.dilation_w;
        val = __fadd_rn(val, __ldg(base + (size_t)kw * g.height_col * g.width_col + w_col));
      }
    }
    im[index] = val;
  }
}

ConvGeom make_geom(int height, int width, int kernel_h, int kernel_w, int pad_h, int pad_w, int stride_h,
                   int stride_w, int dilation_h, int dilation_w, int height_col, int width_col) {
  ConvGeom g;
  g.height = height;
  g.width = width;
  g.kernel_h = kernel_h;
  g.kernel_w = kernel_w;
  g.pad_h = pad_h;
  g.pad_w = pad_w;
  g.stride_h = stride_h;
  g.stride_w = stride_w;
  g.dilation_h = dilation_h;
  g.dilation_w = dilation_w;
  g.height_col = height_col;
  g.width_col = width_col;
  return g;
}

unsigned grid_for(long long work, int num_sms) {
  const long long blocks = (work + 255) / 256;
  const long long cap = (long long)num_sms * 16;       // grid-stride beyond 16 resident blocks per SM
  return (unsigned)(blocks < 1 ? 1 : (blocks < cap ? blocks : cap));
}

}  // namespace

cudaError_t launch_im2col(const float* im, int channels, int height, int width, int kernel_h, int kernel_w,
                          int pad_h, int pad_w, int stride_h, int stride_w, int dilation_h, int dilation_w,
                          int height_col, int width_col, float* col, int num_sms, cudaStream_t stream) {
  const long long groups = (long long)channels * kernel_h * kernel_w * height_col * ((width_col + 3) / 4);
  if (groups <= 0) return cudaSuccess;
  const ConvGeom g = make_geom(height, width, kernel_h, kernel_w, pad_h, pad_w, stride_h, stride_w, dilation_h,
                               dilation_w, height_col, width_col);
  im2col_kernel<<<grid_for(groups, num_sms), 256, 0, stream>>>(groups, im, g, col);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_col2im(const float* col, int channels, int height, int width, int kernel_h, int kernel_w,
                          int pad_h, int pad_w, int stride_h, int stride_w, int dilation_h, int dilation_w,
                          int height_col, int width_col, float* im, int num_sms, cudaStream_t stream) {
  const long long n = (long long)channels * height * width;
  if (n <= 0) return cudaSuccess;
  const ConvGeom g = make_geom(height, width, kernel_h, kernel_w, pad_h, pad_w, stride_h, stride_w, dilation_h,
                               dilation_w, height_col, width_col);
  col2im_kernel<<<grid_for(n, num_sms), 256, 0, stream>>>(n, col, g, im);
  count_launch();
  return cudaGetLastError();
}

}  // namespace mobula_b200
