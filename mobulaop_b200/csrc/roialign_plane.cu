// "plane" kernels: one CTA owns a whole (image, channel) feature-map plane in shared memory.
//
// B200 has 227 KB of shared memory per CTA -- a 200x272 fp32 FPN plane is 217.6 KB -- so the
// NCHW plane the caller hands us is moved global->shared exactly once by the bulk copy
// engine (cp.async.bulk + mbarrier transaction count), every tap of every ROI of that image
// is then a shared-memory access, and HBM sees the algorithmic traffic only:
//   forward : each plane read once, each output written once
//   backward: each dy element read once, each dX element WRITTEN once (the zero-fill of
//             ROIAlign.py:33-34 is fused: the plane is accumulated in shared memory, no
//             global atomics, no read-modify-write of dX)
// ROI decoding comes from the tables of roialign_tables.cu (shared by all C planes).
//
// Arithmetic and its order are the reference's (ROIAlign.cpp:56-74, 126-158 and
// bilinear.h:50-56), so the forward is bit-identical to it and the backward is
// bit-identical to its single-thread execution: every dX element is owned by one thread,
// which adds the contributions in the canonical order (roi, ph, pw, iy, ix, tap).
#include "roialign_kernels.h"
#include "roialign_common.cuh"
#include "ptx_sm100.cuh"

namespace mobula_b200 {

namespace {

constexpr int kMaxThreads = 1024;
constexpr uint32_t kBulkChunk = 32768;

__device__ __forceinline__ void load_plane(float* plane, const float* __restrict__ src, int elems,
                                           uint64_t* bar, uint32_t& parity, int use_bulk) {
  if (use_bulk) {
    if (threadIdx.x == 0) {
      ptx::fence_proxy_async();
      const uint32_t bytes = (uint32_t)elems * 4u;
      ptx::mbar_arrive_expect_tx(bar, bytes);
      for (uint32_t off = 0; off < bytes; off += kBulkChunk) {
        const uint32_t n = bytes - off < kBulkChunk ? bytes - off : kBulkChunk;
        ptx::bulk_g2s((char*)plane + off, (const char*)src + off, n, bar);
      }
    }
    ptx::mbar_wait(bar, parity);
    parity ^= 1u;
  } else {
    for (int i = threadIdx.x; i < elems; i += blockDim.x) plane[i] = __ldg(src + i);
    __syncthreads();
  }
}

struct RecHead {
  int roi, grid_h, grid_w;
  float count;
  int ny, nx, yoff, xoff;
};

__device__ __forceinline__ RecHead load_rec_head(const RoiRec* __restrict__ rec) {
  const int4 a = __ldg(reinterpret_cast<const int4*>(rec));
  const int4 b = __ldg(reinterpret_cast<const int4*>(rec) + 1);
  RecHead h;
  h.roi = a.x;
  h.grid_h = a.y;
  h.grid_w = a.z;
  h.count = __int_as_float(a.w);
  h.ny = b.x;
  h.nx = b.y;
  h.yoff = b.z;
  h.xoff = b.w;
  return h;
}

__device__ __forceinline__ AxisEntry load_entry(const AxisEntry* __restrict__ e) {
  const int4 v = __ldg(reinterpret_cast<const int4*>(e));
  AxisEntry r;
  r.lo = v.x;
  r.hi = v.y;
  r.w_hi = __int_as_float(v.z);
  r.w_lo = __int_as_float(v.w);
  return r;
}

__device__ __forceinline__ AxisEntry decode_entry(float start, int p, float bin, int i, int grid,
                                                  int extent, int mult) {
  const AxisTap t = axis_decode(sample_pos(start, p, bin, i, grid), extent);
  AxisEntry r;
  r.lo = t.valid ? t.lo * mult : -1;
  r.hi = t.valid ? t.hi * mult : -1;
  r.w_hi = t.valid ? t.wl : 0.0f;
  r.w_lo = t.valid ? t.wh : 0.0f;
  return r;
}

// one bilinear sample from the shared-memory plane, bilinear.h:50-56
__device__ __forceinline__ float sample_plane(const float* plane, const AxisEntry& ey, const AxisEntry& ex) {
  const float w1 = __fmul_rn(ey.w_lo, ex.w_lo), w2 = __fmul_rn(ey.w_lo, ex.w_hi);
  const float w3 = __fmul_rn(ey.w_hi, ex.w_lo), w4 = __fmul_rn(ey.w_hi, ex.w_hi);
  const float v1 = plane[ey.lo + ex.lo], v2 = plane[ey.lo + ex.hi];
  const float v3 = plane[ey.hi + ex.lo], v4 = plane[ey.hi + ex.hi];
  return tap_sum(w1, v1, w2, v2, w3, v3, w4, v4);
}

// ---------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------
// GRID > 0: sampling_ratio known at compile time (loops unrolled, x entries kept in
// registers); GRID == 0: any grid, including the adaptive one.
template <int GRID>
__global__ void __launch_bounds__(kMaxThreads, 1)
roialign_fwd_plane_kernel(const float* __restrict__ data, float* __restrict__ out,
                          const float* __restrict__ rois, const RoiRec* __restrict__ recs,
                          const AxisEntry* __restrict__ axis, const int* __restrict__ img_start,
                          int num_images, int channels, int height, int width, int pooled_h,
                          int pooled_w, float scale, int sampling_ratio, int accumulate, int use_bulk) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ uint64_t bar;
  float* plane = reinterpret_cast<float*>(smem_raw);
  const int plane_elems = height * width;
  const int bins = pooled_h * pooled_w;
  const int tid = threadIdx.x;

  // thread -> (roi slot, bin): consecutive threads hold consecutive bins of one ROI, so the
  // output of a (roi, channel) leaves as one contiguous run
  int slots, slot, bin0, bstride;
  if (bins <= (int)blockDim.x) {
    slots = blockDim.x / bins;
    slot = tid / bins;
    bin0 = tid - slot * bins;
    bstride = bins;
    if (slot >= slots) slot = -1;
  } else {
    slots = 1;
    slot = 0;
    bin0 = tid;
    bstride = blockDim.x;
  }

  if (tid == 0 && use_bulk) {
    ptx::mbar_init(&bar, 1);
    ptx::fence_mbar_init();
  }
  __syncthreads();
  uint32_t parity = 0;

  const int planes = (num_images + 1) * channels;
  for (int p = blockIdx.x; p < planes; p += gridDim.x) {
    const int n = p / channels, c = p - n * channels;
    const int j0 = __ldg(img_start + n), j1 = __ldg(img_start + n + 1);
    if (j0 == j1) continue;
    if (n == num_images) {
      // ROIs whose batch index is outside [0, N) read nothing: their output rows are zero
      if (!accumulate && slot >= 0)
        for (int jj = slot; jj < j1 - j0; jj += slots) {
          const int roi = __ldg(&recs[j0 + jj].roi);
          for (int b = bin0; b < bins; b += bstride) out[((size_t)roi * channels + c) * bins + b] = 0.0f;
        }
      continue;
    }
    __syncthreads();   // every thread is done reading the previous plane
    load_plane(plane, data + ((size_t)n * channels + c) * plane_elems, plane_elems, &bar, parity, use_bulk);
    if (slot < 0) continue;

    for (int jj = slot; jj < j1 - j0; jj += slots) {
      const RecHead r = load_rec_head(recs + j0 + jj);
      const int gh = GRID > 0 ? GRID : r.grid_h, gw = GRID > 0 ? GRID : r.grid_w;
      const CountDiv div(gh * gw);
      float* orow = out + ((size_t)r.roi * channels + c) * bins;
      for (int b = bin0; b < bins; b += bstride) {
        const int ph = b / pooled_w, pw = b - ph * pooled_w;
        float acc = 0.0f;
        if (r.yoff >= 0) {
          const AxisEntry* ye = axis + r.yoff + ph * gh;
          const AxisEntry* xe = axis + r.xoff + pw * gw;
          if (GRID > 0) {
            AxisEntry ex[GRID > 0 ? GRID : 1];
#pragma unroll
            for (int ix = 0; ix < GRID; ++ix) ex[ix] = load_entry(xe + ix);
#pragma unroll
            for (int iy = 0; iy < GRID; ++iy) {
              const AxisEntry ey = load_entry(ye + iy);
              if (ey.lo < 0) continue;
#pragma unroll
              for (int ix = 0; ix < GRID; ++ix)
                if (ex[ix].lo >= 0) acc = __fadd_rn(acc, sample_plane(plane, ey, ex[ix]));
            }
          } else {
            for (int iy = 0; iy < gh; ++iy) {
              const AxisEntry ey = load_entry(ye + iy);
              if (ey.lo < 0) continue;
              for (int ix = 0; ix < gw; ++ix) {
                const AxisEntry ex = load_entry(xe + ix);
                if (ex.lo >= 0) acc = __fadd_rn(acc, sample_plane(plane, ey, ex));
              }
            }
          }
        } else {
          // not tabled (arena exhausted): decode on the fly
          const RoiGeom g = roi_decode(rois + (size_t)r.roi * 5, scale, pooled_h, pooled_w, sampling_ratio);
          for (int iy = 0; iy < g.grid_h; ++iy) {
            const AxisEntry ey = decode_entry(g.start_h, ph, g.bin_h, iy, g.grid_h, height, width);
            if (ey.lo < 0) continue;
            for (int ix = 0; ix < g.grid_w; ++ix) {
              const AxisEntry ex = decode_entry(g.start_w, pw, g.bin_w, ix, g.grid_w, width, 1);
              if (ex.lo >= 0) acc = __fadd_rn(acc, sample_plane(plane, ey, ex));
            }
          }
        }
        const float y = div(acc);
        if (accumulate) orow[b] = __fadd_rn(orow[b], y);
        else orow[b] = y;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// backward (owner computes, canonical order)
// ---------------------------------------------------------------------------------------
struct CellRegs {
  int pix, k_first, k_last, p_first;
};

__device__ __forceinline__ CellRegs load_cell(const AxisCell* __restrict__ c) {
  const int4 v = __ldg(reinterpret_cast<const int4*>(c));
  CellRegs r;
  r.pix = v.x;
  r.k_first = v.y;
  r.k_last = v.z;
  r.p_first = v.w;
  return r;
}

// Adds, in canonical order, every contribution of one ROI to the pixel (rowpix / width, col).
// yc / xc give the sample ranges that can touch the pixel's row / column.
template <typename EntryY, typename EntryX>
__device__ __forceinline__ float accumulate_pixel(float acc, const float* __restrict__ dtop, int pooled_w,
                                                  int gh, int gw, const CountDiv& div, int rowpix, int col,
                                                  const CellRegs& yc, const CellRegs& xc, EntryY entry_y,
                                                  EntryX entry_x) {
  int ky = yc.k_first;
  for (int ph = yc.p_first; ky <= yc.k_last; ++ph) {
    const int ky_end = min(yc.k_last, (ph + 1) * gh - 1);   // last sample of bin row ph in range
    int kx = xc.k_first;
    for (int pw = xc.p_first; kx <= xc.k_last; ++pw) {
      const int kx_end = min(xc.k_last, (pw + 1) * gw - 1);
      const float d = __ldg(dtop + ph * pooled_w + pw);
      for (int a = ky; a <= ky_end; ++a) {
        const AxisEntry ey = entry_y(a);
        const bool y_lo = ey.lo == rowpix, y_hi = ey.hi == rowpix;
        if (!(y_lo | y_hi)) continue;
        for (int b = kx; b <= kx_end; ++b) {
          const AxisEntry ex = entry_x(b);
          const bool x_lo = ex.lo == col, x_hi = ex.hi == col;
          // taps in the reference's order: (lo,lo) (lo,hi) (hi,lo) (hi,hi), ROIAlign.cpp:149-156
          if (y_lo && x_lo) acc = __fadd_rn(acc, div(__fmul_rn(d, __fmul_rn(ey.w_lo, ex.w_lo))));
          if (y_lo && x_hi) acc = __fadd_rn(acc, div(__fmul_rn(d, __fmul_rn(ey.w_lo, ex.w_hi))));
          if (y_hi && x_lo) acc = __fadd_rn(acc, div(__fmul_rn(d, __fmul_rn(ey.w_hi, ex.w_lo))));
          if (y_hi && x_hi) acc = __fadd_rn(acc, div(__fmul_rn(d, __fmul_rn(ey.w_hi, ex.w_hi))));
        }
      }
      kx = kx_end + 1;
    }
    ky = ky_end + 1;
  }
  return acc;
}

__global__ void __launch_bounds__(kMaxThreads, 1)
roialign_bwd_plane_kernel(const float* __restrict__ dy, float* __restrict__ dx,
                          const float* __restrict__ rois, const RoiRec* __restrict__ recs,
                          const AxisEntry* __restrict__ axis, const AxisCell* __restrict__ cells,
                          const int* __restrict__ img_start, int num_images, int channels, int height,
                          int width, int pooled_h, int pooled_w, float scale, int sampling_ratio,
                          int accumulate, int use_bulk) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ uint64_t bar;
  float* plane = reinterpret_cast<float*>(smem_raw);
  const int plane_elems = height * width;
  const int bins = pooled_h * pooled_w;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  // rows [b0, b1) of every plane belong to this warp and to nobody else
  const int b0 = (int)(((long long)warp * height) / nwarps);
  const int b1 = (int)(((long long)(warp + 1) * height) / nwarps);
  const int pix0 = b0 * width, pix1 = b1 * width;

  if (tid == 0 && use_bulk) {
    ptx::mbar_init(&bar, 1);
    ptx::fence_mbar_init();
  }
  __syncthreads();
  uint32_t parity = 0;

  const int planes = num_images * channels;
  for (int p = blockIdx.x; p < planes; p += gridDim.x) {
    const int n = p / channels, c = p - n * channels;
    const int j0 = __ldg(img_start + n), j1 = __ldg(img_start + n + 1);
    float* gplane = dx + ((size_t)n * channels + c) * plane_elems;
    if (j0 == j1) {
      if (!accumulate)
        for (int i = tid; i < plane_elems; i += blockDim.x) gplane[i] = 0.0f;
      continue;
    }
    __syncthreads();   // the previous plane has left shared memory
    if (accumulate) {
      load_plane(plane, gplane, plane_elems, &bar, parity, use_bulk);
    } else {
      for (int i = tid; i < plane_elems; i += blockDim.x) plane[i] = 0.0f;
      __syncthreads();
    }

    if (b1 > b0) {
      for (int jb = j0; jb < j1; jb += 32) {
        // which of the next 32 ROIs touch a row of this warp's band?
        const int j = jb + lane;
        bool touches = false;
        if (j < j1) {
          const int4 t = __ldg(reinterpret_cast<const int4*>(recs + j) + 3);   // y_first,y_last,x_first,x_last
          const int yoff = __ldg(&recs[j].yoff);
          touches = yoff < 0 || (t.x < b1 && t.y >= b0);
        }
        unsigned todo = __ballot_sync(0xffffffffu, touches);
        while (todo) {
          const int s = __ffs(todo) - 1;
          todo &= todo - 1;
          const RoiRec* rec = recs + jb + s;
          const RecHead r = load_rec_head(rec);
          const float* dtop = dy + ((size_t)r.roi * channels + c) * bins;
          const CountDiv div(r.grid_h * r.grid_w);
          if (r.yoff >= 0) {
            const int4 cinfo = __ldg(reinterpret_cast<const int4*>(rec) + 2);  // ycell,xcell,nycell,nxcell
            const AxisCell* ycells = cells + cinfo.x;
            const AxisCell* xcells = cells + cinfo.y;
            const AxisEntry* ytab = axis + r.yoff;
            const AxisEntry* xtab = axis + r.xoff;
            auto entry_y = [&](int k) { return load_entry(ytab + k); };
            auto entry_x = [&](int k) { return load_entry(xtab + k); };
            for (int yb = 0; yb < cinfo.z; yb += 32) {
              CellRegs mine;
              mine.pix = -1;
              if (yb + lane < cinfo.z) mine = load_cell(ycells + yb + lane);
              unsigned rows = __ballot_sync(0xffffffffu, mine.pix >= pix0 && mine.pix < pix1);
              while (rows) {
                const int q = __ffs(rows) - 1;
                rows &= rows - 1;
                CellRegs yc;
                yc.pix = __shfl_sync(0xffffffffu, mine.pix, q);
                yc.k_first = __shfl_sync(0xffffffffu, mine.k_first, q);
                yc.k_last = __shfl_sync(0xffffffffu, mine.k_last, q);
                yc.p_first = __shfl_sync(0xffffffffu, mine.p_first, q);
                for (int xb = lane; xb < cinfo.w; xb += 32) {
                  const CellRegs xc = load_cell(xcells + xb);
                  float* px = plane + yc.pix + xc.pix;
                  *px = accumulate_pixel(*px, dtop, pooled_w, r.grid_h, r.grid_w, div, yc.pix, xc.pix, yc,
                                         xc, entry_y, entry_x);
                }
              }
            }
          } else {
            // not tabled: decode every sample on the fly (correct for any grid, slow)
            const RoiGeom g = roi_decode(rois + (size_t)r.roi * 5, scale, pooled_h, pooled_w, sampling_ratio);
            auto entry_y = [&](int k) {
              return decode_entry(g.start_h, k / g.grid_h, g.bin_h, k % g.grid_h, g.grid_h, height, width);
            };
            auto entry_x = [&](int k) {
              return decode_entry(g.start_w, k / g.grid_w, g.bin_w, k % g.grid_w, g.grid_w, width, 1);
            };
            CellRegs yc, xc;
            yc.k_first = 0;
            yc.k_last = pooled_h * g.grid_h - 1;
            yc.p_first = 0;
            xc.k_first = 0;
            xc.k_last = pooled_w * g.grid_w - 1;
            xc.p_first = 0;
            for (int row = b0; row < b1; ++row)
              for (int col = lane; col < width; col += 32) {
                float* px = plane + row * width + col;
                *px = accumulate_pixel(*px, dtop, pooled_w, g.grid_h, g.grid_w, div, row * width, col, yc,
                                       xc, entry_y, entry_x);
              }
          }
          __syncwarp();   // the next ROI may hand this pixel to another lane
        }
      }
    }
    if (use_bulk) ptx::fence_proxy_async();   // my shared-memory writes -> visible to the copy engine
    __syncthreads();
    // the finished plane leaves for HBM: each dX element is written exactly once
    if (use_bulk) {
      if (tid == 0) {
        const uint32_t bytes = (uint32_t)plane_elems * 4u;
        for (uint32_t off = 0; off < bytes; off += kBulkChunk) {
          const uint32_t nb = bytes - off < kBulkChunk ? bytes - off : kBulkChunk;
          ptx::bulk_s2g((char*)gplane + off, (const char*)plane + off, nb);
        }
        ptx::bulk_commit();
        ptx::bulk_wait_read();
      }
    } else {
      for (int i = tid; i < plane_elems; i += blockDim.x) gplane[i] = plane[i];
    }
  }
  if (use_bulk && tid == 0) ptx::bulk_wait_all();
}

template <typename K>
cudaError_t set_smem(K kernel, size_t bytes) {
  return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

int ctas_per_sm(size_t smem_bytes, int threads, int smem_optin) {
  // 1 KB of shared memory per CTA is reserved by the driver
  int by_smem = (int)((size_t)(smem_optin + 1024) / (smem_bytes + 1024 + 64));
  int by_threads = 2048 / threads;
  int n = by_smem < by_threads ? by_smem : by_threads;
  return n < 1 ? 1 : (n > 8 ? 8 : n);
}

}  // namespace

bool plane_supported(const RoiAlignArgs& a, const void* plane_ptr, int smem_optin, PlanePlan* plan) {
  if (a.num_images < 0 || a.num_images > 4096) return false;
  if ((long long)(a.num_images + 1) * a.num_rois > (1ll << 27)) return false;   // grouping cost
  const size_t plane_bytes = (size_t)a.height * a.width * sizeof(float);
  const size_t smem = (plane_bytes + 127) / 128 * 128;
  if (smem + 64 > (size_t)smem_optin) return false;
  plan->smem_bytes = smem;
  plan->threads = kMaxThreads;
  plan->use_bulk = (plane_bytes % 16 == 0) && ((uintptr_t)plane_ptr % 16 == 0) ? 1 : 0;
  plan->ctas_per_sm = ctas_per_sm(smem, kMaxThreads, smem_optin);
  return true;
}

cudaError_t launch_fwd_plane(const RoiAlignArgs& a, const PlanePlan& plan, const RoiTables& t,
                             const float* data, float* out, int accumulate, int num_sms,
                             cudaStream_t stream) {
  const long long planes = (long long)(a.num_images + 1) * a.channels;
  if (planes == 0 || a.num_rois == 0) return cudaSuccess;
  long long grid = (long long)num_sms * plan.ctas_per_sm;
  if (grid > planes) grid = planes;
#define MOBULA_FWD(G)                                                                            \
  do {                                                                                           \
    cudaError_t e = set_smem(roialign_fwd_plane_kernel<G>, plan.smem_bytes);                     \
    if (e != cudaSuccess) return e;                                                              \
    roialign_fwd_plane_kernel<G><<<(unsigned)grid, plan.threads, plan.smem_bytes, stream>>>(     \
        data, out, a.rois, t.recs, t.axis, t.img_start, a.num_images, a.channels, a.height,      \
        a.width, a.pooled_h, a.pooled_w, a.scale, a.sampling_ratio, accumulate, plan.use_bulk);  \
  } while (0)
  switch (a.sampling_ratio) {
    case 1: MOBULA_FWD(1); break;
    case 2: MOBULA_FWD(2); break;
    case 3: MOBULA_FWD(3); break;
    case 4: MOBULA_FWD(4); break;
    default: MOBULA_FWD(0); break;
  }
#undef MOBULA_FWD
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_bwd_plane(const RoiAlignArgs& a, const PlanePlan& plan, const RoiTables& t,
                             const float* dy, float* dx, int accumulate, int num_sms,
                             cudaStream_t stream) {
  const long long planes = (long long)a.num_images * a.channels;
  if (planes == 0) return cudaSuccess;
  long long grid = (long long)num_sms * plan.ctas_per_sm;
  if (grid > planes) grid = planes;
  cudaError_t e = set_smem(roialign_bwd_plane_kernel, plan.smem_bytes);
  if (e != cudaSuccess) return e;
  roialign_bwd_plane_kernel<<<(unsigned)grid, plan.threads, plan.smem_bytes, stream>>>(
      dy, dx, a.rois, t.recs, t.axis, t.cells, t.img_start, a.num_images, a.channels, a.height, a.width,
      a.pooled_h, a.pooled_w, a.scale, a.sampling_ratio, accumulate, plan.use_bulk);
  count_launch();
  return cudaGetLastError();
}

}  // namespace mobula_b200
