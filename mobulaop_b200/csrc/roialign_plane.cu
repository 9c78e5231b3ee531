// "plane" kernels: one CTA owns a whole (image, channel) feature-map plane in shared memory.
//
// B200 has 227 KB of shared memory per CTA -- a 200x272 fp32 FPN plane is 217.6 KB -- so the
// NCHW plane the caller hands us is moved global->shared exactly once by the bulk copy
// engine (cp.async.bulk + mbarrier transaction count), every tap of every ROI of that image
// is then a shared-memory access, and HBM sees the algorithmic traffic only:
//   forward : each plane read once, each output written once
//   backward: each dX element WRITTEN once (the zero-fill of ROIAlign.py:33-34 is fused: the
//             plane is accumulated in shared memory, no global atomics, no read-modify-write of
//             dX), contributions added in the reference's single-thread order (deterministic
//             mode; the fast non-deterministic backward lives in roialign_resident.cu)
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

// Moves a [height x width] fp32 plane between global memory (dense rows) and shared memory
// (rows `pitch` floats apart).  With use_bulk every row is one cp.async.bulk issued by a lane
// of warp 0; completion is tracked by `bar` (loads) or the bulk async-group (stores).
__device__ __forceinline__ void load_plane(float* plane, const float* __restrict__ src, int height,
                                           int width, int pitch, uint64_t* bar, uint32_t& parity,
                                           int use_bulk) {
  if (use_bulk) {
    if (threadIdx.x < 32) {
      const uint32_t row_bytes = (uint32_t)width * 4u;
      if (threadIdx.x == 0) {
        ptx::fence_proxy_async();
        ptx::mbar_arrive_expect_tx(bar, row_bytes * (uint32_t)height);
      }
      __syncwarp();
      if (pitch == width) {
        // dense: a few large copies
        const uint32_t bytes = row_bytes * (uint32_t)height, chunk = 32768u;
        for (uint32_t off = threadIdx.x * chunk; off < bytes; off += 32u * chunk)
          ptx::bulk_g2s((char*)plane + off, (const char*)src + off, min(chunk, bytes - off), bar);
      } else {
        for (int r = threadIdx.x; r < height; r += 32)
          ptx::bulk_g2s(plane + (size_t)r * pitch, src + (size_t)r * width, row_bytes, bar);
      }
    }
    ptx::mbar_wait(bar, parity);
    parity ^= 1u;
  } else {
    for (int i = threadIdx.x; i < height * width; i += blockDim.x) {
      const int r = i / width;
      plane[r * pitch + (i - r * width)] = __ldg(src + i);
    }
    __syncthreads();
  }
}

// Call with every thread, after a __syncthreads() that followed the last shared-memory write
// (each writer having executed fence_proxy_async() when use_bulk is set).
__device__ __forceinline__ void store_plane(float* __restrict__ dst, const float* plane, int height,
                                            int width, int pitch, int use_bulk) {
  if (use_bulk) {
    if (threadIdx.x < 32) {
      const uint32_t row_bytes = (uint32_t)width * 4u;
      if (pitch == width) {
        const uint32_t bytes = row_bytes * (uint32_t)height, chunk = 32768u;
        for (uint32_t off = threadIdx.x * chunk; off < bytes; off += 32u * chunk)
          ptx::bulk_s2g((char*)dst + off, (const char*)plane + off, min(chunk, bytes - off));
      } else {
        for (int r = threadIdx.x; r < height; r += 32)
          ptx::bulk_s2g(dst + (size_t)r * width, plane + (size_t)r * pitch, row_bytes);
      }
      ptx::bulk_commit();
      ptx::bulk_wait_read();   // the shared-memory source may be overwritten afterwards
    }
  } else {
    for (int i = threadIdx.x; i < height * width; i += blockDim.x) {
      const int r = i / width;
      dst[i] = plane[r * pitch + (i - r * width)];
    }
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
// One pooled output.  GRID > 0: sampling_ratio known at compile time (unrolled, x entries in
// registers, no branch when every sample is inside the plane); GRID == 0: any grid.
template <int GRID>
__device__ __forceinline__ float pool_bin(const float* plane, const AxisEntry* __restrict__ ye,
                                          const AxisEntry* __restrict__ xe, int gh, int gw) {
  float acc = 0.0f;
  if (GRID > 0) {
    AxisEntry ey[GRID > 0 ? GRID : 1], ex[GRID > 0 ? GRID : 1];
    int any_invalid = 0;
#pragma unroll
    for (int i = 0; i < GRID; ++i) {
      ey[i] = load_entry(ye + i);
      ex[i] = load_entry(xe + i);
      any_invalid |= ey[i].lo | ex[i].lo;      // invalid entries carry lo = -1
    }
    if (any_invalid >= 0) {
#pragma unroll
      for (int iy = 0; iy < GRID; ++iy)
#pragma unroll
        for (int ix = 0; ix < GRID; ++ix) acc = __fadd_rn(acc, sample_plane(plane, ey[iy], ex[ix]));
    } else {
#pragma unroll
      for (int iy = 0; iy < GRID; ++iy)
#pragma unroll
        for (int ix = 0; ix < GRID; ++ix)
          if ((ey[iy].lo | ex[ix].lo) >= 0) acc = __fadd_rn(acc, sample_plane(plane, ey[iy], ex[ix]));
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
  return acc;
}

template <int GRID>
__global__ void __launch_bounds__(kMaxThreads, 1)
roialign_fwd_plane_kernel(const float* __restrict__ data, float* __restrict__ out,
                          const float* __restrict__ rois, const RoiRec* __restrict__ recs,
                          const AxisEntry* __restrict__ axis, const int* __restrict__ img_start,
                          int num_images, int channels, int height, int width, int pitch, int pooled_h,
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
    bstride = bins;              // a single bin per thread
    if (slot >= slots) slot = -1;
  } else {
    slots = 1;
    slot = 0;
    bin0 = tid;
    bstride = blockDim.x;
  }
  const int ph0 = bin0 / pooled_w, pw0 = bin0 - ph0 * pooled_w;

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
    load_plane(plane, data + ((size_t)n * channels + c) * plane_elems, height, width, pitch, &bar, parity,
               use_bulk);
    if (slot < 0) continue;

    for (int jj = slot; jj < j1 - j0; jj += slots) {
      const RecHead r = load_rec_head(recs + j0 + jj);
      const int gh = GRID > 0 ? GRID : r.grid_h, gw = GRID > 0 ? GRID : r.grid_w;
      const CountDiv div(gh * gw);
      float* orow = out + ((size_t)r.roi * channels + c) * bins;
      int ph = ph0, pw = pw0;
      for (int b = bin0; b < bins; b += bstride) {
        if (b != bin0) {
          ph = b / pooled_w;
          pw = b - ph * pooled_w;
        }
        float acc;
        if (r.yoff >= 0) {
          acc = pool_bin<GRID>(plane, axis + r.yoff + ph * gh, axis + r.xoff + pw * gw, gh, gw);
        } else {
          // not tabled (arena exhausted): decode on the fly
          acc = 0.0f;
          const RoiGeom g = roi_decode(rois + (size_t)r.roi * 5, scale, pooled_h, pooled_w, sampling_ratio);
          for (int iy = 0; iy < g.grid_h; ++iy) {
            const AxisEntry ey = decode_entry(g.start_h, ph, g.bin_h, iy, g.grid_h, height, pitch);
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

// Forward for a compile-time sampling ratio with every ROI tabled: the tables are then
// regular (roi j owns entries [j*E, (j+1)*E), E = (PH+PW)*GRID), so no per-ROI record has to
// be read before the sample entries, and the loads of U consecutive items of a thread are all
// issued before the first one is used (the table latency, not the arithmetic, bounds this
// kernel otherwise).  512 threads leave each thread 128 registers for that.
constexpr int kRegularThreads = 512;

template <int GRID, int U>
__global__ void __launch_bounds__(kRegularThreads, 1)
roialign_fwd_plane_regular_kernel(const float* __restrict__ data, float* __restrict__ out,
                                  const int* __restrict__ order, const AxisEntry* __restrict__ axis,
                                  const int* __restrict__ img_start, int num_images, int channels,
                                  int height, int width, int pitch, int pooled_h, int pooled_w,
                                  int accumulate, int use_bulk) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ uint64_t bar;
  float* plane = reinterpret_cast<float*>(smem_raw);
  const int plane_elems = height * width;
  const int bins = pooled_h * pooled_w;
  const int tid = threadIdx.x;
  const int slots = blockDim.x / bins;          // launcher guarantees bins <= blockDim.x
  int slot = tid / bins;
  const int bin = tid - slot * bins;
  if (slot >= slots) slot = -1;
  const int ph = bin / pooled_w, pw = bin - ph * pooled_w;
  const int per_roi = (pooled_h + pooled_w) * GRID;
  const AxisEntry* ybase = axis + ph * GRID;
  const AxisEntry* xbase = axis + pooled_h * GRID + pw * GRID;
  const CountDiv div(GRID * GRID);

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
    const int nroi = j1 - j0;
    if (nroi == 0) continue;
    if (n == num_images) {
      if (!accumulate && slot >= 0)
        for (int jj = slot; jj < nroi; jj += slots)
          out[((size_t)__ldg(order + j0 + jj) * channels + c) * bins + bin] = 0.0f;
      continue;
    }
    __syncthreads();   // every thread is done reading the previous plane
    load_plane(plane, data + ((size_t)n * channels + c) * plane_elems, height, width, pitch, &bar, parity,
               use_bulk);
    if (slot < 0) continue;

    for (int jj = slot; jj < nroi; jj += U * slots) {
      AxisEntry ey[U][GRID], ex[U][GRID];
      int roi[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int j = jj + u * slots;
        roi[u] = -1;
        if (j < nroi) {
          const size_t base = (size_t)(j0 + j) * per_roi;
          roi[u] = __ldg(order + j0 + j);
#pragma unroll
          for (int i = 0; i < GRID; ++i) {
            ey[u][i] = load_entry(ybase + base + i);
            ex[u][i] = load_entry(xbase + base + i);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (roi[u] < 0) continue;
        int any_invalid = 0;
#pragma unroll
        for (int i = 0; i < GRID; ++i) any_invalid |= ey[u][i].lo | ex[u][i].lo;
        float acc = 0.0f;
        if (any_invalid >= 0) {
#pragma unroll
          for (int iy = 0; iy < GRID; ++iy)
#pragma unroll
            for (int ix = 0; ix < GRID; ++ix) acc = __fadd_rn(acc, sample_plane(plane, ey[u][iy], ex[u][ix]));
        } else {
#pragma unroll
          for (int iy = 0; iy < GRID; ++iy)
#pragma unroll
            for (int ix = 0; ix < GRID; ++ix)
              if ((ey[u][iy].lo | ex[u][ix].lo) >= 0)
                acc = __fadd_rn(acc, sample_plane(plane, ey[u][iy], ex[u][ix]));
        }
        const float y = div(acc);
        float* o = out + ((size_t)roi[u] * channels + c) * bins + bin;
        if (accumulate) *o = __fadd_rn(*o, y);
        else *o = y;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// backward (owner computes): kBands warps per CTA, warp w owns rows [w*H/kBands, (w+1)*H/kBands)
// of the plane and nobody else touches them -> plain shared-memory read-modify-write.
// ---------------------------------------------------------------------------------------
constexpr int kBwdThreads = kBands * 32;

struct CellRegs {
  int pix, k_first, k_last, p_first;
};

__device__ __forceinline__ CellRegs load_cell(const AxisCell* __restrict__ c) {
  const int4 v = __ldg(reinterpret_cast<const int4*>(c));   // pix, p_first, k_first, k_last
  CellRegs r;
  r.pix = v.x;
  r.p_first = v.y;
  r.k_first = v.z;
  r.k_last = v.w;
  return r;
}

// Adds, in canonical order, every contribution of one ROI to the pixel (rowpix / pitch, col).
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
      const float d = dtop[ph * pooled_w + pw];
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

// Copies dy[roi, c, :, :] (bins floats) into the warp's slot; returns the pointer to read from.
__device__ __forceinline__ const float* stage_dtop(const float* __restrict__ dtop, float* gslot, int bins,
                                                   int lane) {
  if (gslot == nullptr) return dtop;
  __syncwarp();                       // the previous tile is no longer being read
  for (int i = lane; i < bins; i += 32) gslot[i] = __ldg(dtop + i);
  __syncwarp();
  return gslot;
}

// ---- plane prologue / epilogue shared by both backward kernels ----
struct BwdPlane {
  float* plane;
  float* gslot;
  uint64_t* bar;
  uint32_t parity;
};

// Canonical order: every contribution added separately, in the order of the single-thread
// reference -> bit-exact dX ("deterministic" mode).
__global__ void __launch_bounds__(kBwdThreads, 1)
roialign_bwd_plane_canonical_kernel(const float* __restrict__ dy, float* __restrict__ dx,
                                    const float* __restrict__ rois, const RoiRec* __restrict__ recs,
                                    const AxisEntry* __restrict__ axis, const AxisCell* __restrict__ cells,
                                    const int* __restrict__ img_start, int num_images, int channels,
                                    int height, int width, int pitch, int pooled_h, int pooled_w,
                                    float scale, int sampling_ratio, int accumulate, int use_bulk,
                                    int gslot_floats) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ uint64_t bar;
  float* plane = reinterpret_cast<float*>(smem_raw);
  const int plane_elems = height * width;
  const int bins = pooled_h * pooled_w;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float* gslot = gslot_floats > 0 ? plane + (size_t)height * pitch + (size_t)warp * gslot_floats : nullptr;
  const int b0 = (int)(((long long)warp * height) / kBands);
  const int b1 = (int)(((long long)(warp + 1) * height) / kBands);
  const int pix0 = b0 * pitch, pix1 = b1 * pitch;

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
      load_plane(plane, gplane, height, width, pitch, &bar, parity, use_bulk);
    } else {
      for (int i = tid; i < height * pitch; i += blockDim.x) plane[i] = 0.0f;
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
          const CountDiv div(r.grid_h * r.grid_w);
          if (r.yoff >= 0) {
            const int4 cinfo = __ldg(reinterpret_cast<const int4*>(rec) + 2);  // ycell,xcell,nycell,nxcell
            const AxisCell* ycells = cells + cinfo.x;
            const AxisCell* xcells = cells + cinfo.y;
            const AxisEntry* ytab = axis + r.yoff;
            const AxisEntry* xtab = axis + r.xoff;
            auto entry_y = [&](int k) { return load_entry(ytab + k); };
            auto entry_x = [&](int k) { return load_entry(xtab + k); };
            const float* dtop = nullptr;
            for (int yb = 0; yb < cinfo.z; yb += 32) {
              CellRegs mine;
              mine.pix = -1;
              mine.k_first = mine.k_last = mine.p_first = 0;
              if (yb + lane < cinfo.z) mine = load_cell(ycells + yb + lane);
              unsigned rows = __ballot_sync(0xffffffffu, mine.pix >= pix0 && mine.pix < pix1);
              if (rows && dtop == nullptr)
                dtop = stage_dtop(dy + ((size_t)r.roi * channels + c) * bins, gslot, bins, lane);
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
            const float* dtop = stage_dtop(dy + ((size_t)r.roi * channels + c) * bins, gslot, bins, lane);
            auto entry_y = [&](int k) {
              return decode_entry(g.start_h, k / g.grid_h, g.bin_h, k % g.grid_h, g.grid_h, height, pitch);
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
                float* px = plane + row * pitch + col;
                *px = accumulate_pixel(*px, dtop, pooled_w, g.grid_h, g.grid_w, div, row * pitch, col, yc,
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
    store_plane(gplane, plane, height, width, pitch, use_bulk);
  }
  if (use_bulk && tid < 32) ptx::bulk_wait_all();
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

// Row pitch against shared-memory bank conflicts: lanes of a warp read the same columns of
// several rows, so consecutive rows must start in different banks.  A pitch that is 4 mod 8
// floats gives 8 distinct row phases and keeps rows 16-byte aligned for the bulk copies.
int pick_pitch(int width) {
  int p = (width + 3) / 4 * 4;
  if ((p / 4) % 2 == 0) p += 4;
  return p;
}

}  // namespace

bool plane_supported(const RoiAlignArgs& a, const void* plane_ptr, int smem_optin, PlanePlan* plan) {
  if (a.num_images < 0 || a.num_images > 4096) return false;
  if ((long long)(a.num_images + 1) * a.num_rois > (1ll << 27)) return false;   // grouping cost
  const size_t budget = (size_t)smem_optin - 64;      // static shared memory of the kernels
  const size_t dense = (size_t)a.height * a.width * sizeof(float);
  if (dense > budget) return false;
  int pitch = pick_pitch(a.width);
  if ((size_t)a.height * pitch * sizeof(float) > budget) pitch = a.width;   // dense rows as a last resort
  plan->pitch = pitch;
  plan->plane_bytes = (size_t)a.height * pitch * sizeof(float);
  plan->bwd_pitch = a.width;
  plan->bwd_plane_bytes = dense;
  const int bins = a.pooled_h * a.pooled_w;
  const size_t gslot = (size_t)((bins + 31) / 32 * 32);
  plan->gslot_floats = dense + gslot * sizeof(float) * kBands <= budget ? (int)gslot : 0;
  plan->use_bulk = (a.width % 4 == 0) && ((uintptr_t)plane_ptr % 16 == 0) ? 1 : 0;
  plan->ctas_per_sm = ctas_per_sm(plan->plane_bytes, kRegularThreads, smem_optin);
  plan->bwd_ctas_per_sm =
      ctas_per_sm(dense + (size_t)plan->gslot_floats * sizeof(float) * kBands, kBwdThreads, smem_optin);
  return true;
}

cudaError_t launch_fwd_plane(const RoiAlignArgs& a, const PlanePlan& plan, const RoiTables& t,
                             const float* data, float* out, int accumulate, int num_sms,
                             cudaStream_t stream) {
  const long long planes = (long long)(a.num_images + 1) * a.channels;
  if (planes == 0 || a.num_rois == 0) return cudaSuccess;
  long long grid = (long long)num_sms * plan.ctas_per_sm;
  if (grid > planes) grid = planes;
  const int bins = a.pooled_h * a.pooled_w;
  const bool regular = t.regular && bins <= kRegularThreads && a.sampling_ratio >= 1 && a.sampling_ratio <= 4;
#define MOBULA_FWD_REG(G, U)                                                                        \
  do {                                                                                              \
    cudaError_t e = set_smem(roialign_fwd_plane_regular_kernel<G, U>, plan.plane_bytes);            \
    if (e != cudaSuccess) return e;                                                                 \
    roialign_fwd_plane_regular_kernel<G, U><<<(unsigned)grid, kRegularThreads, plan.plane_bytes,    \
                                              stream>>>(                                            \
        data, out, t.order, t.axis, t.img_start, a.num_images, a.channels, a.height, a.width,       \
        plan.pitch, a.pooled_h, a.pooled_w, accumulate, plan.use_bulk);                             \
  } while (0)
#define MOBULA_FWD(G)                                                                            \
  do {                                                                                           \
    cudaError_t e = set_smem(roialign_fwd_plane_kernel<G>, plan.plane_bytes);                    \
    if (e != cudaSuccess) return e;                                                              \
    roialign_fwd_plane_kernel<G><<<(unsigned)grid, kMaxThreads, plan.plane_bytes, stream>>>(    \
        data, out, a.rois, t.recs, t.axis, t.img_start, a.num_images, a.channels, a.height,      \
        a.width, plan.pitch, a.pooled_h, a.pooled_w, a.scale, a.sampling_ratio, accumulate,      \
        plan.use_bulk);                                                                          \
  } while (0)
  if (regular) {
    switch (a.sampling_ratio) {
      case 1: MOBULA_FWD_REG(1, 4); break;
      case 2: MOBULA_FWD_REG(2, 4); break;
      case 3: MOBULA_FWD_REG(3, 2); break;
      default: MOBULA_FWD_REG(4, 2); break;
    }
  } else {
    MOBULA_FWD(0);
  }
#undef MOBULA_FWD
#undef MOBULA_FWD_REG
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_bwd_plane(const RoiAlignArgs& a, const PlanePlan& plan, const RoiTables& t,
                             const float* dy, float* dx, int accumulate, int num_sms,
                             cudaStream_t stream) {
  const long long planes = (long long)a.num_images * a.channels;
  if (planes == 0) return cudaSuccess;
  long long grid = (long long)num_sms * plan.bwd_ctas_per_sm;
  if (grid > planes) grid = planes;
  const size_t smem = plan.bwd_plane_bytes + (size_t)plan.gslot_floats * sizeof(float) * kBands;
  cudaError_t e = set_smem(roialign_bwd_plane_canonical_kernel, smem);
  if (e != cudaSuccess) return e;
  roialign_bwd_plane_canonical_kernel<<<(unsigned)grid, kBwdThreads, smem, stream>>>(
      dy, dx, a.rois, t.recs, t.axis, t.cells, t.img_start, a.num_images, a.channels, a.height, a.width,
      plan.bwd_pitch, a.pooled_h, a.pooled_w, a.scale, a.sampling_ratio, accumulate, plan.use_bulk,
      plan.gslot_floats);
  count_launch();
  return cudaGetLastError();
}

}  // namespace mobula_b200
