// ROI tables: the channel-independent part of ROIAlign, computed once per call.
//
// The reference decodes the ROI and every sample coordinate again for each of the
// C*PH*PW outputs of a ROI (ROIAlign.cpp:24-65).  Here three small kernels do it once:
//   1. roi_group_kernel    stable partition of the ROI rows by image (ascending inside an
//                          image = the canonical accumulation order of the backward)
//   2. roi_records_kernel  per-ROI geometry + arena offsets (exclusive scan)
//   3. roi_fill_kernel     per-axis sample tables and, for the backward, the inverse map
//                          row/column -> contiguous range of samples touching it
// All arithmetic goes through roialign_common.cuh, i.e. rounds like the reference.
#include "roialign_kernels.h"
#include "roialign_common.cuh"
#include "ptx_sm100.cuh"

namespace mobula_b200 {

namespace {

constexpr int kGroupThreads = 1024;
constexpr int kRecThreads = 1024;

__global__ void __launch_bounds__(kGroupThreads)
roi_group_kernel(const float* __restrict__ rois, int num_rois, int num_images,
                 int* __restrict__ img_start, int* __restrict__ order) {
  ptx::chain_enter();
  __shared__ int warp_sums[kGroupThreads / 32];
  roi_group_block<kGroupThreads>(rois, num_rois, num_images, blockIdx.x, img_start, order, warp_sums);
}

__global__ void __launch_bounds__(kRecThreads)
roi_records_kernel(const float* __restrict__ rois, int num_rois, int num_images, int pooled_h,
                   int pooled_w, float scale, int sampling_ratio, const int* __restrict__ img_start,
                   const int* __restrict__ order, RoiRec* __restrict__ recs, long long arena) {
  __shared__ long long warp_sums[kRecThreads / 32];
  const int first_invalid = img_start[num_images];
  long long running = 0;
  for (int base = 0; base < num_rois; base += kRecThreads) {
    const int j = base + threadIdx.x;
    RoiRec rec;
    long long need = 0;
    if (j < num_rois) {
      rec.roi = order[j];
      const RoiGeom g = roi_decode(rois + (size_t)rec.roi * 5, scale, pooled_h, pooled_w, sampling_ratio);
      rec.grid_h = g.grid_h;
      rec.grid_w = g.grid_w;
      rec.count = g.count;
      const long long ny = (long long)pooled_h * max(g.grid_h, 0);
      const long long nx = (long long)pooled_w * max(g.grid_w, 0);
      const bool tabled_size = ny + nx < (1ll << 28);
      rec.ny = tabled_size ? (int)ny : 0;
      rec.nx = tabled_size ? (int)nx : 0;
      need = (j < first_invalid) ? (tabled_size ? ny + nx : (1ll << 40)) : 0;
    }
    long long total;
    const long long off = running + block_exclusive_scan<kRecThreads>(need, &total, warp_sums);
    if (j < num_rois) {
      const bool fits = j < first_invalid && off + need <= arena;
      rec.yoff = fits ? (int)off : -1;
      rec.xoff = fits ? (int)off + rec.ny : -1;
      rec.ycell = fits ? (int)(2 * off) : -1;
      rec.xcell = fits ? (int)(2 * off) + 2 * rec.ny : -1;
      rec.nycell = rec.nxcell = 0;
      rec.y_first = rec.x_first = 0;
      rec.y_last = rec.x_last = -1;
      recs[j] = rec;
    }
    running += total;
    if (running > (1ll << 50)) running = 1ll << 50;   // saturate: everything later is untabled
  }
}

// One thread per (roi, axis): walks the samples of the axis in order.
__global__ void __launch_bounds__(128)
roi_fill_kernel(const float* __restrict__ rois, int num_rois, int height, int width, int pitch,
                int pooled_h, int pooled_w, float scale, int sampling_ratio, RoiRec* __restrict__ recs,
                AxisEntry* __restrict__ axis, AxisCell* __restrict__ cells, unsigned need) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = t >> 1, ax = t & 1;        // ax 0: y (rows), 1: x (columns)
  if (j >= num_rois) return;
  RoiRec* rec = recs + j;
  const int off = ax ? rec->xoff : rec->yoff;
  if (off < 0) return;
  const RoiGeom g = roi_decode(rois + (size_t)rec->roi * 5, scale, pooled_h, pooled_w, sampling_ratio);
  const int n = ax ? rec->nx : rec->ny;
  const int grid = ax ? g.grid_w : g.grid_h;
  const float start = ax ? g.start_w : g.start_h;
  const float bin = ax ? g.bin_w : g.bin_h;
  const int extent = ax ? width : height;
  const int mult = ax ? 1 : pitch;
  AxisEntry* e = axis + off;
  const int cell_base = ax ? rec->xcell : rec->ycell;
  AxisCell* c = cells + cell_base;
  int ncell = 0;
  int p = 0, i = 0;
  int pix_min = 0x7fffffff, pix_max = -1;
  for (int k = 0; k < n; ++k) {
    const AxisTap tap = axis_decode(sample_pos(start, p, bin, i, grid), extent);
    AxisEntry en;
    en.lo = tap.valid ? tap.lo * mult : -1;
    en.hi = tap.valid ? tap.hi * mult : -1;
    en.w_hi = tap.valid ? tap.wl : 0.0f;
    en.w_lo = tap.valid ? tap.wh : 0.0f;
    e[k] = en;
    if (need && tap.valid) {
      // lo and hi are non-decreasing in k and hi <= lo + 1, so a touched pixel is normally
      // the newest cell, the one before it, or a new one; the search only goes further back
      // if rounding ever made the positions non-monotonic (cells stay unique per pixel)
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        const int pix = (s ? tap.hi : tap.lo) * mult;
        int q = ncell - 1;
        while (q >= 0 && c[q].pix != pix && (q > ncell - 3 || c[q].pix > pix)) --q;
        if (q >= 0 && c[q].pix == pix) {
          c[q].k_last = k;
        } else {
          AxisCell cell;
          cell.pix = pix;
          cell.k_first = cell.k_last = k;
          cell.p_first = p;
          c[ncell++] = cell;
          pix_min = min(pix_min, pix);
          pix_max = max(pix_max, pix);
        }
      }
    }
    if (++i == grid) {
      i = 0;
      ++p;
    }
  }
  if (need) {
    const int first = ncell ? pix_min / mult : 0;
    const int last = ncell ? pix_max / mult : -1;
    if (ax) {
      rec->nxcell = ncell;
      rec->x_first = first;
      rec->x_last = last;
    } else {
      rec->nycell = ncell;
      rec->y_first = first;
      rec->y_last = last;
    }
  }
}

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

thread_local long long t_arena_override = -1;

long long arena_entries(const RoiAlignArgs& a) {
  if (t_arena_override >= 0) return t_arena_override;
  const long long per_grid = (long long)a.pooled_h + a.pooled_w;
  if (a.sampling_ratio > 0) return per_grid * a.sampling_ratio * a.num_rois;
  // adaptive grids: room for an average grid of 12 samples per bin side; ROIs that do not
  // fit stay untabled and are decoded on the fly by the kernels
  long long e = per_grid * 12 * a.num_rois;
  if (e < (1ll << 20)) e = 1ll << 20;
  if (e > (1ll << 27)) e = 1ll << 27;
  return e;
}

struct Layout {
  size_t img_start, order, recs, axis, cells, total;
};

Layout layout_of(const RoiAlignArgs& a, unsigned need) {
  const size_t R = (size_t)a.num_rois, N = (size_t)(a.num_images > 0 ? a.num_images : 0);
  const size_t arena = (size_t)arena_entries(a);
  Layout l;
  size_t p = 0;
  l.img_start = p;  p += align_up((N + 2) * sizeof(int), 256);
  l.order = p;      p += align_up(R * sizeof(int), 256);
  l.recs = p;       p += align_up(R * sizeof(RoiRec), 256);
  l.axis = p;       p += align_up(arena * sizeof(AxisEntry), 256);
  l.cells = p;      p += need ? align_up(2 * arena * sizeof(AxisCell), 256) : 0;
  l.total = p;
  return l;
}

}  // namespace

void roi_tables_set_arena_override(long long entries) { t_arena_override = entries; }

size_t roi_tables_bytes(const RoiAlignArgs& a, unsigned need) { return layout_of(a, need).total; }

RoiTables roi_tables_carve(const RoiAlignArgs& a, unsigned need, int pitch, void* base) {
  const Layout l = layout_of(a, need);
  char* p = (char*)base;
  RoiTables t;
  t.img_start = (int*)(p + l.img_start);
  t.order = (int*)(p + l.order);
  t.recs = (RoiRec*)(p + l.recs);
  t.axis = (AxisEntry*)(p + l.axis);
  t.cells = (AxisCell*)(p + l.cells);
  t.arena = arena_entries(a);
  t.pitch = pitch;
  t.regular = (a.sampling_ratio > 0 && t_arena_override < 0) ? 1 : 0;
  return t;
}

cudaError_t launch_roi_group(const RoiAlignArgs& a, int* img_start, int* order, cudaStream_t stream) {
  if (a.num_rois == 0) return cudaMemsetAsync(img_start, 0, (size_t)(a.num_images + 2) * sizeof(int), stream);
  if (a.pre_img_start && a.pre_order) {
    // prepared once for this ROI table (forward and backward share it): two small copies
    cudaError_t e = cudaMemcpyAsync(img_start, a.pre_img_start, (size_t)(a.num_images + 2) * sizeof(int),
                                    cudaMemcpyDeviceToDevice, stream);
    if (e != cudaSuccess) return e;
    return cudaMemcpyAsync(order, a.pre_order, (size_t)a.num_rois * sizeof(int), cudaMemcpyDeviceToDevice, stream);
  }
  chained_launch(roi_group_kernel, a.num_images + 1, kGroupThreads, 0, stream, a.rois, a.num_rois, a.num_images,
                                                                   img_start, order);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_roi_tables(const RoiAlignArgs& a, const RoiTables& t, unsigned need,
                              cudaStream_t stream) {
  if (a.num_rois == 0) {
    // every segment is empty
    return cudaMemsetAsync(t.img_start, 0, (size_t)(a.num_images + 2) * sizeof(int), stream);
  }
  roi_group_kernel<<<a.num_images + 1, kGroupThreads, 0, stream>>>(a.rois, a.num_rois, a.num_images,
                                                                   t.img_start, t.order);
  count_launch();
  roi_records_kernel<<<1, kRecThreads, 0, stream>>>(a.rois, a.num_rois, a.num_images, a.pooled_h,
                                                    a.pooled_w, a.scale, a.sampling_ratio,
                                                    t.img_start, t.order, t.recs, t.arena);
  count_launch();
  const int threads = 2 * a.num_rois;
  roi_fill_kernel<<<(threads + 127) / 128, 128, 0, stream>>>(
      a.rois, a.num_rois, a.height, a.width, t.pitch, a.pooled_h, a.pooled_w, a.scale,
      a.sampling_ratio, t.recs, t.axis, t.cells, need);
  count_launch();
  return cudaGetLastError();
}

}  // namespace mobula_b200
