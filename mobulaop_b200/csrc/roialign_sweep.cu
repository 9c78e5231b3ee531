// "sweep" kernels: lane = channel.
//
// The resident kernels keep ONE channel plane in shared memory and map lanes to samples, so
// every bilinear tap is a 4-byte gather with whatever bank conflicts the ROI geometry produces
// (measured 1.8 wavefronts per load, 47 % of all shared-memory wavefronts are replays).  Here a
// CTA holds a few full-width ROWS of 32 channels of one image and the 32 lanes of a warp are
// those 32 channels: a warp works on one sample at a time, the tap address is warp-uniform up
// to the lane's channel offset, and the channel offsets are laid out so that the 32 lanes hit
// the 32 banks -- one wavefront per tap load whatever the ROI looks like.  Geometry being
// warp-uniform also means no divergence: adaptive grids (sampling_ratio <= 0), any pooled
// shape, clamped and rejected samples are uniform branches.
//
// Data movement: a ring of full-width rows in shared memory, [row slot][32 channels][pitch] with an
// ODD pitch, so channel c of pixel x sits at word c * pitch + x and the 32 channels of one pixel
// occupy the 32 banks.  That skew needs 4-byte placement: every copy engine that moves 16-byte
// granules (TMA tensor / bulk copies, 16-byte cp.async) keeps pixel x at a word == x (mod 4) and
// would leave the 32 lanes on 8 banks (tools/tma_probe.cu, profiles/r2_tma_probe.txt: a tiled TMA
// load with a start coordinate that is not a multiple of 16 bytes raises an illegal instruction).
// The ring is therefore filled by four producer warps with coalesced LDG.128 (8 lanes = 32 pixels
// of one channel, 4 channels per instruction) and 4 x STS.32, conflict-free on both sides; 16
// consumer warps compute.  The CTA sweeps its rows top to bottom; every feature-map row is read
// once per sweep.
//
// Work: (image, 32-channel group, row segment) units from a queue; inside a unit the "items"
// -- runs of sample rows of one bin row of one ROI whose taps span at most kSpan rows -- sorted
// by first row and handed to the 16 consumer warps in that order.  A bin row whose sample rows
// are too far apart for one item is a chain of items: the running sums travel through a
// scratch buffer (L2) and a "done" bit per item, so the reference's summation order (sample
// rows outer, sample columns inner, ROIAlign.cpp:56-74) is kept and the result is bit-identical.
// A bin row belongs to the segment that holds its LAST sample row; a segment's sweep starts as
// far above its first row as its bin rows reach, so no partial sum crosses CTAs.
//
// Reference arithmetic: /root/reference/opzoo/ROIAlign/ROIAlign.cpp:9-76, bilinear.h:10-59.
#include <cstdio>
#include <cstdlib>
#include <mutex>

#include "roialign_kernels.h"
#include "roialign_common.cuh"
#include "ptx_sm100.cuh"

namespace mobula_b200 {

namespace {

constexpr int kConsumers = 16;                        // warps that compute
constexpr int kProducers = 4;                         // warps that fill the row ring
constexpr int kFillBatch = 8;                         // 16-byte loads in flight per producer lane
constexpr int kSweepThreads = (kConsumers + kProducers) * 32;
constexpr int kMaxRing = 16;
constexpr int kMinRing = 4;
constexpr int kMaxPooledW = 16;
constexpr int kMaxPlaneH = 4095;                      // item records keep rows in 12 bits

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ---- table records ----------------------------------------------------------------------
// One ROI, channel independent (64 bytes).
struct __align__(16) SweepHdr {
  int roi;               // row of the caller's roi table = output row
  int batch;             // image, -1: batch index outside [0, N)
  int grid;              // grid_h | grid_w << 16
  float count;           // (float)(grid_h * grid_w)
  unsigned yr, xr;       // plane rows / columns the ROI's samples touch: min | max << 16 (0xffff | 0: none)
  unsigned col_base;     // first column record of this ROI
  int ncols;             // compacted tables: number of (valid) column records
  unsigned short cstart[16];   // compacted tables: first record of bin column pw (end = next / ncols)
};

// A sample column: pixel columns of the low / high tap (lo | hi << 16; 0xffffffff: the sample is
// outside the plane, bilinear.h:15-18) and the fraction lx = weight of the high tap.
typedef uint2 ColRec;
constexpr unsigned kColInvalid = 0xffffffffu;
// The sampling_ratio == 2 kernel takes its sample columns ready to use (16 bytes): byte offsets
// of the two taps inside a channel row, lx, hx = 1 - lx.  A sample outside the plane points both
// taps at the zero pad column behind the row with lx = hx = 0: it contributes exactly +0, no branch.
typedef uint4 ColRec4;
// A sample row inside the plane: lo | hi << 16 (plane rows of the two taps), ly.  Indexed by the
// ordinal of the row among the ROI's sample rows inside the plane.
typedef uint2 RowRec;

// An item: n sample rows (row records [ord, ord + n)) of bin row ph of ROI `roi`; taps in plane
// rows [key, key + dy].
//   w0 roi   w1 ord | n << 16   w2 key | dy << 12 | ph << 16 | flags << 30   w3 predecessor item
// flags: 1 = first item of its bin row (sums start at 0), 2 = last (divide and write the output)
typedef uint4 ItemRec;
constexpr unsigned kNoDep = 0xffffffffu;

struct SweepParams {
  const float* rois;
  float* out;
  int num_rois, num_images;
  int channels;          // channel stride of data / out
  int c_work;            // channels this call works on
  int ncg;               // ceil(c_work / 32)
  int height, width, pooled_h, pooled_w;
  float scale;
  int sampling_ratio;
  int aligned;
  int accumulate;
  int top_dtype, top_layout;   // element type / layout of `out` (RoiAlignArgs::top_dtype, top_layout)
  const float* data;
  // shared-memory geometry
  int pitch;             // floats between channels of a ring row (odd, >= width)
  int fill_mode;         // 0: LDG.128 + STS.32 (needs 16-byte aligned rows), 1: 4-byte cp.async
  int rowf;              // floats per ring row = 32 * pitch
  int ring, span;        // ring rows, max rows an item's taps may span
  unsigned ring_magic;   // floor(2^32 / ring) + 1: q / ring == umulhi(q, magic)
  int pwp;               // staging pitch per channel (pooled_w | 1)
  int segs;              // row segments per image
  int colcap;            // column records per ROI
  int rowcap;            // row records per ROI
  int compact;           // column records hold only the samples inside the plane, per-bin ranges
                         // in the header (every sampling ratio but 2)
  // tables
  int* cnt;              // [N * segs * H] items per (image, segment, first row); cursor in pass 2
  int* start;            // [N * segs * H + 1] exclusive scan of cnt
  int* unit_rows;        // [N * segs] 1 + last row any item of the unit touches
  int* unit_counter;     // persistent-CTA queue
  unsigned* done;        // one bit per (item, channel group)
  SweepHdr* hdr;         // [R]
  ColRec* cols;          // [R * colcap] (ColRec4 when !compact)
  RowRec* rows;          // [R * rowcap]
  uint2* phr;            // [R * PH] bin row: first | end row record, first tap row | last tap row << 16 (backward)
  uint2* pwr;            // [R * 16] bin column: first | end column record, first | last tap column (backward)
  int tables_only;       // pass 1 builds the per-ROI tables and no items (backward)
  ItemRec* items;
  float* scratch;        // running sums of chained items: [(roi * PH + ph) * ncg + cg][PW][32]
};

__host__ __device__ inline int seg_of_row(int y, int segs, int height) { return (int)(((long long)y * segs) / height); }

// ---- pass 1 / pass 2 over the ROIs: one warp per ROI ---------------------------------------
// PASS 1: header, column records, item counts.  PASS 2: item records at their sorted position.
template <int PASS>
__global__ void __launch_bounds__(256) sweep_roi_kernel(SweepParams p) {
  ptx::chain_enter();
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= p.num_rois) return;
  const RoiGeom g = roi_decode(p.rois + (size_t)r * 5, p.scale, p.pooled_h, p.pooled_w, p.sampling_ratio, p.aligned != 0);
  const int n = (g.batch >= 0 && g.batch < p.num_images) ? g.batch : -1;
  const int PH = p.pooled_h, PW = p.pooled_w, H = p.height, W = p.width;
  if (PASS == 1) {
    SweepHdr* h = p.hdr + r;
    const unsigned col_base = (unsigned)r * (unsigned)p.colcap;
    if (lane == 0) {
      h->roi = r;
      h->batch = n;
      h->grid = g.grid_h | (g.grid_w << 16);
      h->count = g.count;
      h->col_base = col_base;
    }
    if (n < 0) {
      // a ROI that points at no image reads nothing: its output rows are zero
      if (!p.accumulate && !p.tables_only) {
        const size_t row0 = (size_t)r * p.channels * PH * PW;
        if (p.top_layout == 0) {
          const size_t per = (size_t)p.c_work * PH * PW;
          for (size_t i = lane; i < per; i += 32) store_top(p.out, row0 + i, 0.0f, p.top_dtype);
        } else {
          for (int b = 0; b < PH * PW; ++b)
            for (int c = lane; c < p.c_work; c += 32) store_top(p.out, row0 + (size_t)b * p.channels + c, 0.0f, p.top_dtype);
        }
      }
      if (lane == 0) {
        h->ncols = 0;
        h->yr = h->xr = 0xffffu;
      }
      return;
    }
    ColRec* cols = p.cols + col_base;
    int xmin = 0xffff, xmax = 0;
    if (!p.compact) {
      ColRec4* cols4 = reinterpret_cast<ColRec4*>(p.cols) + col_base;
      const int gw = g.grid_w, total = PW * gw;
      for (int k = lane; k < total; k += 32) {
        const int pw = k / gw, ix = k - pw * gw;
        const AxisTap t = axis_decode(sample_pos(g.start_w, pw, g.bin_w, ix, gw), W);
        ColRec4 c = make_uint4((unsigned)W * 4u, (unsigned)W * 4u, 0u, 0u);      // the zero pad column
        if (t.valid) {
          c = make_uint4((unsigned)t.lo * 4u, (unsigned)t.hi * 4u, __float_as_uint(t.wl), __float_as_uint(t.wh));
          xmin = min(xmin, t.lo);
          xmax = max(xmax, t.hi);
        }
        cols4[k] = c;
      }
      if (lane == 0) h->ncols = total;
    } else {
      // only the samples inside the plane are kept, bin column by bin column (adaptive grids:
      // at most 2W + 3 of them, the sample spacing bin / ceil(bin) is at least half a pixel)
      int base = 0;
      for (int pw = 0; pw < PW; ++pw) {
        if (lane == 0 && pw < 16) h->cstart[pw] = (unsigned short)base;
        const int bin_first = base;
        int bmin = 0xffff, bmax = 0;
        for (int i0 = 0; i0 < g.grid_w; i0 += 32) {
          const int ix = i0 + lane;
          bool valid = false;
          ColRec c = make_uint2(kColInvalid, 0u);
          if (ix < g.grid_w) {
            const AxisTap t = axis_decode(sample_pos(g.start_w, pw, g.bin_w, ix, g.grid_w), W);
            valid = t.valid;
            if (valid) {
              c = make_uint2((unsigned)t.lo | ((unsigned)t.hi << 16), __float_as_uint(t.wl));
              xmin = min(xmin, t.lo);
              xmax = max(xmax, t.hi);
              bmin = min(bmin, t.lo);
              bmax = max(bmax, t.hi);
            }
          }
          const unsigned m = __ballot_sync(0xffffffffu, valid);
          if (valid) {
            const int pos = base + __popc(m & ((1u << lane) - 1u));
            if (pos < p.colcap) cols[pos] = c;
          }
          base += __popc(m);
        }
        if (p.pwr) {
#pragma unroll
          for (int d = 16; d > 0; d >>= 1) {
            bmin = min(bmin, __shfl_xor_sync(0xffffffffu, bmin, d));
            bmax = max(bmax, __shfl_xor_sync(0xffffffffu, bmax, d));
          }
          if (lane == 0 && pw < 16)
            p.pwr[(size_t)r * 16 + pw] = make_uint2((unsigned)bin_first | ((unsigned)min(base, p.colcap) << 16),
                                                    (unsigned)(bmin & 0xffff) | ((unsigned)bmax << 16));
        }
      }
      if (base > p.colcap) base = p.colcap;     // cannot happen (see the bound above)
      if (lane == 0) {
        for (int pw = PW; pw < 16; ++pw) h->cstart[pw] = (unsigned short)base;
        h->ncols = base;
      }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      xmin = min(xmin, __shfl_xor_sync(0xffffffffu, xmin, d));
      xmax = max(xmax, __shfl_xor_sync(0xffffffffu, xmax, d));
    }
    if (lane == 0) h->xr = (unsigned)xmin | ((unsigned)xmax << 16);
  }
  if (n < 0) return;
  // ---- items: lane = bin row --------------------------------------------------------------
  RowRec* rows = p.rows + (size_t)r * p.rowcap;
  int ymin = 0xffff, ymax = 0;
  int ord_base = 0;       // sample rows inside the plane in the bin rows before this chunk of 32
  for (int ph0 = 0; ph0 < PH; ph0 += 32) {
    const int ph = ph0 + lane;
    // pass A: the last sample row inside the plane decides the segment of the bin row; the
    // number of rows inside the plane gives the ordinal of its first row record
    int last_lo = -1, inside = 0, first_lo = -1, last_hi = 0;
    if (ph < PH)
      for (int iy = 0; iy < g.grid_h; ++iy) {
        const AxisTap t = axis_decode(sample_pos(g.start_h, ph, g.bin_h, iy, g.grid_h), H);
        if (t.valid) {
          if (first_lo < 0) first_lo = t.lo;
          last_lo = t.lo;
          last_hi = t.hi;
          ++inside;
        }
      }
    int incl = inside;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += v;
    }
    int ord = ord_base + incl - inside;
    ord_base += __shfl_sync(0xffffffffu, incl, 31);
    if (PASS == 1 && p.phr && ph < PH)
      p.phr[(size_t)r * PH + ph] = make_uint2((unsigned)ord | ((unsigned)(ord + inside) << 16),
                                              (unsigned)(first_lo < 0 ? 0 : first_lo) | ((unsigned)last_hi << 16));
    if (PASS == 1 && p.tables_only) {
      if (ph < PH)
        for (int iy = 0; iy < g.grid_h; ++iy) {
          const AxisTap t = axis_decode(sample_pos(g.start_h, ph, g.bin_h, iy, g.grid_h), H);
          if (!t.valid) continue;
          if (ord < p.rowcap) rows[ord] = make_uint2((unsigned)t.lo | ((unsigned)t.hi << 16), __float_as_uint(t.wl));
          ymin = min(ymin, t.lo);
          ymax = max(ymax, t.hi);
          ++ord;
        }
      continue;
    }
    if (ph >= PH) continue;
    const int seg = last_lo < 0 ? 0 : seg_of_row(last_lo, p.segs, H);
    const int unit = n * p.segs + seg;
    int* cnt = p.cnt + (size_t)unit * H;
    const int* start = p.start + (size_t)unit * H;
    unsigned dep = kNoDep;
    int open_key = -1, open_hi = 0, open_ord = 0, open_n = 0;
    bool first = true;
    auto emit = [&](bool last) {
      const int key = open_key < 0 ? 0 : open_key;
      if (PASS == 1) {
        atomicAdd(cnt + key, 1);
        atomicMax(p.unit_rows + unit, open_hi + 1);
      } else {
        const unsigned pos = (unsigned)(__ldg(start + key) + atomicAdd(cnt + key, 1));
        const unsigned flags = (first ? 1u : 0u) | (last ? 2u : 0u);
        p.items[pos] = make_uint4((unsigned)r, (unsigned)open_ord | ((unsigned)open_n << 16),
                                  (unsigned)key | ((unsigned)(open_hi - key) << 12) | ((unsigned)ph << 16) |
                                      (flags << 30),
                                  dep);
        dep = pos;
      }
      first = false;
    };
    for (int iy = 0; iy < g.grid_h; ++iy) {
      const AxisTap t = axis_decode(sample_pos(g.start_h, ph, g.bin_h, iy, g.grid_h), H);
      if (!t.valid) continue;
      if (PASS == 1 && ord < p.rowcap)
        rows[ord] = make_uint2((unsigned)t.lo | ((unsigned)t.hi << 16), __float_as_uint(t.wl));
      ymin = min(ymin, t.lo);
      ymax = max(ymax, t.hi);
      if (open_key >= 0 && t.hi - open_key + 1 > p.span) {
        emit(false);
        open_key = -1;
      }
      if (open_key < 0) {
        open_key = t.lo;
        open_ord = ord;
        open_n = 0;
      }
      open_hi = t.hi;
      ++open_n;
      ++ord;
    }
    if (open_key < 0) {       // no sample row inside the plane: the bin row is all zeros
      open_ord = 0;
      open_n = 0;
      open_hi = 0;
    }
    emit(true);
  }
  if (PASS == 1) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      ymin = min(ymin, __shfl_xor_sync(0xffffffffu, ymin, d));
      ymax = max(ymax, __shfl_xor_sync(0xffffffffu, ymax, d));
    }
    if (lane == 0) p.hdr[r].yr = (unsigned)ymin | ((unsigned)ymax << 16);
  }
}

// exclusive scan of cnt[0, m) into start[0, m]; cnt is cleared (pass 2 uses it as cursors)
__global__ void __launch_bounds__(1024) sweep_scan_kernel(int* __restrict__ cnt, int* __restrict__ start, int m) {
  ptx::chain_enter();
  __shared__ int warp_sums[32];
  const int per = (m + 1023) / 1024;
  const int lo = threadIdx.x * per, hi = min(m, lo + per);
  int sum = 0;
  for (int i = lo; i < hi; ++i) sum += cnt[i];
  int total;
  int run = block_exclusive_scan<1024>(sum, &total, warp_sums);
  for (int i = lo; i < hi; ++i) {
    const int c = cnt[i];
    start[i] = run;
    cnt[i] = 0;
    run += c;
  }
  if (threadIdx.x == 0) start[m] = total;
}

// ---- ring barriers ------------------------------------------------------------------------
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(ptx::smem_addr(bar)) : "memory");
}

// A consumer warp is done with ring rows [from, to) of the current unit.  Barrier phases are told
// apart by ONE parity bit, so a thread may only wait for fill r of a slot after it has seen fill
// r - 1 complete (otherwise "fill r - 1 still pending" reads as "fill r done"), and nobody may
// get two phases ahead of anybody.  Hence every lane waits for every row of the unit in order,
// needed or not, and a row is released only after it has been loaded -- which also implies that
// every warp had released the previous user of its slot.
__device__ __forceinline__ void release_rows(uint64_t* full, uint64_t* empty, unsigned seq0, int from, int to,
                                             int ring, unsigned magic, int lane) {
  for (int q = from; q < to; ++q) {
    const unsigned s = seq0 + (unsigned)q;
    const unsigned round = __umulhi(s, magic), slot = s - round * (unsigned)ring;
    ptx::mbar_wait(full + slot, round & 1u);
    if (lane == 0) mbar_arrive(empty + slot);
  }
}

// 4-byte asynchronous copy global -> shared (LDGSTS.32)
__device__ __forceinline__ void cp_async4(unsigned dst_smem, const float* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst_smem), "l"(src) : "memory");
}
// arrival on `bar` once every cp.async this thread issued so far has completed (the barrier's
// expected count already includes it)
__device__ __forceinline__ void cp_async_arrive(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(ptx::smem_addr(bar)) : "memory");
}

__device__ __forceinline__ float lds_f32(unsigned addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}

__device__ __forceinline__ float ld_cg(const float* p) {
  float v;
  asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
// orders this thread's earlier and later global accesses at GPU scope (no L1 invalidation: the
// running sums go through st.cg / ld.cg)
__device__ __forceinline__ void fence_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ void st_cg(float* p, float v) { asm volatile("st.global.cg.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory"); }

// A finished bin row in a half-precision and / or channel-last output (RoiAlignArgs::top_dtype,
// top_layout); `vals` = the lane's channel, one value per bin column.
__device__ __noinline__ void sweep_store_typed(const SweepParams& p, const float* vals, int pooled_w, float* stage,
                                               int roi, int ph, int c0, int cvalid, int bins, int lane) {
  if (p.top_layout != 0) {
    // channel-last [R, PH, PW, C]: the lanes ARE the channels, one coalesced store per bin
    const size_t o = (((size_t)roi * p.pooled_h + ph) * pooled_w) * p.channels + c0 + lane;
    if (lane < cvalid)
      for (int pw = 0; pw < pooled_w; ++pw) store_top(p.out, o + (size_t)pw * p.channels, vals[pw], p.top_dtype);
    return;
  }
  __syncwarp();
  for (int pw = 0; pw < pooled_w; ++pw) stage[lane * p.pwp + pw] = vals[pw];
  __syncwarp();
  const size_t o = ((size_t)roi * p.channels + c0) * bins + ph * pooled_w;
  for (int f = lane; f < 32 * pooled_w; f += 32) {
    const int c = f / pooled_w, pw = f - c * pooled_w;
    if (c < cvalid) store_top(p.out, o + (size_t)c * bins + pw, stage[c * p.pwp + pw], p.top_dtype);
  }
}

// ---- the forward sweep ---------------------------------------------------------------------
// PWT: pooled width.  GWT > 0: fixed sampling ratio, column records at k = pw * GWT + ix;
// GWT == 0: compacted column records with per-bin ranges (adaptive grids).
// TYPED: the output is half precision and / or channel-last (a variant of its own: the reference's
// format keeps exactly the code -- and the register allocation -- it had without the feature; a
// run-time test in the epilogue alone cost the adaptive-grid workload 25 %).
template <int PWT, int GWT, bool TYPED>
__global__ void __launch_bounds__(kSweepThreads, 1)
sweep_fwd_kernel(const SweepParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* ring = reinterpret_cast<float*>(smem_raw);
  float* stage_all = ring + (size_t)p.ring * p.rowf;
  uint64_t* full = reinterpret_cast<uint64_t*>(stage_all + kConsumers * 32 * p.pwp);
  uint64_t* empty = full + kMaxRing;
  int* ctl = reinterpret_cast<int*>(empty + kMaxRing);     // [0] unit, [1] item cursor

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < p.ring; ++i) {
      ptx::mbar_init(full + i, kProducers * 32);
      ptx::mbar_init(empty + i, kConsumers);
    }
    ptx::fence_mbar_init();
  }
  ptx::chain_enter();
  const int H = p.height, PH = p.pooled_h;
  const int bins = PH * PWT;
  // the zero pad column behind every channel row (samples outside the plane read it)
  for (int i = threadIdx.x; i < p.ring * 32; i += kSweepThreads) ring[(size_t)i * p.pitch + p.width] = 0.0f;
  const unsigned ring_base = ptx::smem_addr(ring);
  const unsigned lane_off = (unsigned)(lane * p.pitch * 4);
  const unsigned row_bytes = (unsigned)p.rowf * 4u;
  float* stage = stage_all + (warp < kConsumers ? warp : 0) * 32 * p.pwp;
  const int units = p.num_images * p.segs * p.ncg;
  unsigned seq0 = 0;      // ring rows loaded before this unit (same in every thread)

  for (;;) {
    __syncthreads();      // every warp is done with the previous unit (and its ctl words)
    if (threadIdx.x == 0) {
      ctl[0] = atomicAdd(p.unit_counter, 1);
      ctl[1] = 0;
    }
    __syncthreads();
    const int u = ctl[0];
    if (u >= units) break;
    const int cg = u % p.ncg, us = u / p.ncg;              // us = image * segs + segment
    const int n = us / p.segs;
    const int item_lo = __ldg(p.start + (size_t)us * H), item_hi = __ldg(p.start + (size_t)(us + 1) * H);
    if (item_lo >= item_hi) continue;
    const int ya = (int)(__ldg(&p.items[item_lo].z) & 0xfffu);    // items are sorted by first row
    const int nrows = __ldg(p.unit_rows + us) - ya;         // rows [ya, ya + nrows) are swept
    const int c0 = cg * 32;

    if (warp >= kConsumers) {
      // ---- producers: warp pw fills channels [8 pw, 8 pw + 8) of every ring row with 4-byte
      // cp.async (LDGSTS): lane = pixel, 128 coalesced bytes per instruction land on 32 consecutive
      // banks; nothing is staged in registers, so a producer runs as many rows ahead as the ring
      // has free slots.  Each lane's arrival on full[slot] fires when its own copies have landed.
      const int pw8 = (warp - kConsumers) * 8;
      const int W = p.width;
      const int cend = min(pw8 + 8, min(32, p.c_work - c0));
      const size_t plane = (size_t)H * W;
      const float* src0 = p.data + ((size_t)n * p.channels + c0) * plane + (size_t)ya * W + lane;
      for (int q = 0; q < nrows; ++q) {
        const unsigned s = seq0 + (unsigned)q;
        const unsigned round = __umulhi(s, p.ring_magic), slot = s - round * (unsigned)p.ring;
        if (round > 0) ptx::mbar_wait(empty + slot, (round & 1u) ^ 1u);   // every consumer has left the old row
        const float* src_row = src0 + (size_t)q * W;
        if (p.fill_mode == 1) {
          const unsigned dst_row = ring_base + slot * row_bytes + (unsigned)lane * 4u;
          for (int c = pw8; c < cend; ++c) {
            const float* src = src_row + (size_t)c * plane;
            const unsigned dst = dst_row + (unsigned)(c * p.pitch) * 4u;
#pragma unroll 3
            for (int x = 0; x < W - lane; x += 32) cp_async4(dst + (unsigned)x * 4u, src + x);
          }
          cp_async_arrive(full + slot);
        } else {
          // LDG.128 + 4 x STS.32: lane = (channel ci of 4, quad j of 8): 4 channels x 32 pixels per
          // load instruction; the four stores of a quad hit banks (ci * pitch + 4 j + k) mod 32, all
          // distinct.  Eight loads are in flight per lane.
          const int j = lane & 7, ci = lane >> 3;
          const float* src = src_row - lane;                    // src0 carries + lane
          float* row = ring + (size_t)slot * p.rowf;
          const int nxc = (W + 31) >> 5, tasks = 2 * nxc;       // (channel half, 32-pixel chunk)
          for (int t0 = 0; t0 < tasks; t0 += kFillBatch) {
            float4 v[kFillBatch];
#pragma unroll
            for (int u = 0; u < kFillBatch; ++u) {
              const int t = t0 + u, half = t >= nxc ? 1 : 0, xc = t - half * nxc;
              const int c = pw8 + half * 4 + ci, x = xc * 32 + 4 * j;
              v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
              if (t < tasks && x < W && c < cend)
                v[u] = __ldg(reinterpret_cast<const float4*>(src + (size_t)c * plane + x));
            }
#pragma unroll
            for (int u = 0; u < kFillBatch; ++u) {
              const int t = t0 + u, half = t >= nxc ? 1 : 0, xc = t - half * nxc;
              const int c = pw8 + half * 4 + ci, x = xc * 32 + 4 * j;
              if (t < tasks && x < W) {
                float* d = row + c * p.pitch + x;
                d[0] = v[u].x;
                d[1] = v[u].y;
                d[2] = v[u].z;
                d[3] = v[u].w;
              }
            }
          }
          mbar_arrive(full + slot);
        }
      }
    } else {
      // ---- consumers ------------------------------------------------------------------
      int my_q = 0;       // rows [0, my_q) have been released by this warp
      for (;;) {
        int idx = 0;
        if (lane == 0) idx = atomicAdd(&ctl[1], 1);
        idx = __shfl_sync(0xffffffffu, idx, 0) + item_lo;
        if (idx >= item_hi) break;
        const ItemRec it = __ldg(p.items + idx);
        const int roi = (int)it.x;
        const SweepHdr* hdr = p.hdr + roi;
        const uint2 gc = __ldg(reinterpret_cast<const uint2*>(hdr) + 1);   // grid sizes, count: used at the end
        const size_t col_base = (size_t)roi * p.colcap;
        const ColRec* cols = p.cols + col_base;
        const ColRec4* cols4 = reinterpret_cast<const ColRec4*>(p.cols) + col_base;
        const int nsy = (int)(it.y >> 16);
        const RowRec* rows = p.rows + (size_t)roi * p.rowcap + (it.y & 0xffffu);
        const int key = (int)(it.z & 0xfffu), dy = (int)((it.z >> 12) & 0xfu);
        const int ph = (int)((it.z >> 16) & 0x3fffu);
        const unsigned flags = it.z >> 30;
        // ring: release the rows above this item, wait for the rows it reads
        const int q0 = key - ya;
        release_rows(full, empty, seq0, my_q, q0, p.ring, p.ring_magic, lane);
        if (q0 > my_q) my_q = q0;
        if (nsy > 0)
          for (int q = q0; q <= q0 + dy; ++q) {
            const unsigned s = seq0 + (unsigned)q;
            const unsigned round = __umulhi(s, p.ring_magic);
            ptx::mbar_wait(full + (s - round * (unsigned)p.ring), round & 1u);
          }
        // running sums: zero, or what the predecessor item left
        float acc[PWT];
        float* scr = p.scratch + (((size_t)roi * PH + ph) * p.ncg + cg) * (PWT * 32) + lane;
        if (flags & 1u) {
#pragma unroll
          for (int pw = 0; pw < PWT; ++pw) acc[pw] = 0.0f;
        } else {
          // one "done" bit per (item, channel group): the same item list serves every group
          const unsigned long long dep = (unsigned long long)it.w * (unsigned)p.ncg + (unsigned)cg;
          if (lane == 0) {
            const volatile unsigned* word = p.done + (dep >> 5);
            while (!((*word >> (unsigned)(dep & 31u)) & 1u)) {
            }
          }
          __syncwarp();
          fence_gpu();
#pragma unroll
          for (int pw = 0; pw < PWT; ++pw) acc[pw] = ld_cg(scr + pw * 32);
        }
        uint4 cs0 = make_uint4(0, 0, 0, 0), cs1 = cs0;
        int ncols = 0;
        if (GWT == 0) {
          cs0 = __ldg(reinterpret_cast<const uint4*>(hdr) + 2);
          cs1 = __ldg(reinterpret_cast<const uint4*>(hdr) + 3);
          ncols = __ldg(&hdr->ncols);
        }
        for (int k = 0; k < nsy; ++k) {
          const RowRec rr = __ldg(rows + k);
          const unsigned slo = seq0 + ((rr.x & 0xffffu) - (unsigned)ya), shi = seq0 + ((rr.x >> 16) - (unsigned)ya);
          const unsigned row_lo = ring_base + lane_off + (slo - __umulhi(slo, p.ring_magic) * (unsigned)p.ring) * row_bytes;
          const unsigned row_hi = ring_base + lane_off + (shi - __umulhi(shi, p.ring_magic) * (unsigned)p.ring) * row_bytes;
          const float ly = __uint_as_float(rr.y), hy = __fsub_rn(1.0f, ly);
          if (GWT > 0) {
#pragma unroll
            for (int pw = 0; pw < PWT; ++pw) {
#pragma unroll
              for (int ix = 0; ix < GWT; ++ix) {
                const ColRec4 c = __ldg(cols4 + pw * GWT + ix);
                const float lx = __uint_as_float(c.z), hx = __uint_as_float(c.w);
                const float v1 = lds_f32(row_lo + c.x), v2 = lds_f32(row_lo + c.y);
                const float v3 = lds_f32(row_hi + c.x), v4 = lds_f32(row_hi + c.y);
                const float s = tap_sum(__fmul_rn(hy, hx), v1, __fmul_rn(hy, lx), v2, __fmul_rn(ly, hx), v3,
                                        __fmul_rn(ly, lx), v4);
                acc[pw] = __fadd_rn(acc[pw], s);
              }
            }
          } else {
            unsigned prev = kColInvalid;
            float v1 = 0.f, v2 = 0.f, v3 = 0.f, v4 = 0.f;
#pragma unroll
            for (int pw = 0; pw < PWT; ++pw) {
              const unsigned w = pw < 8 ? (pw < 4 ? (pw < 2 ? cs0.x : cs0.y) : (pw < 6 ? cs0.z : cs0.w))
                                        : (pw < 12 ? (pw < 10 ? cs1.x : cs1.y) : (pw < 14 ? cs1.z : cs1.w));
              const int k0 = (int)((pw & 1) ? (w >> 16) : (w & 0xffffu));
              int k1 = ncols;
              if (pw + 1 < 16) {
                const int q = pw + 1;
                const unsigned w1 = q < 8 ? (q < 4 ? (q < 2 ? cs0.x : cs0.y) : (q < 6 ? cs0.z : cs0.w))
                                          : (q < 12 ? (q < 10 ? cs1.x : cs1.y) : (q < 14 ? cs1.z : cs1.w));
                k1 = (int)((q & 1) ? (w1 >> 16) : (w1 & 0xffffu));
              }
              for (int k = k0; k < k1; ++k) {
                const ColRec c = __ldg(cols + k);
                if (c.x != prev) {      // neighbouring samples of a dense grid share their taps
                  const unsigned olo = (c.x & 0xffffu) << 2, ohi = (c.x >> 16) << 2;
                  v1 = lds_f32(row_lo + olo);
                  v2 = lds_f32(row_lo + ohi);
                  v3 = lds_f32(row_hi + olo);
                  v4 = lds_f32(row_hi + ohi);
                  prev = c.x;
                }
                const float lx = __uint_as_float(c.y), hx = __fsub_rn(1.0f, lx);
                const float s = tap_sum(__fmul_rn(hy, hx), v1, __fmul_rn(hy, lx), v2, __fmul_rn(ly, hx), v3,
                                        __fmul_rn(ly, lx), v4);
                acc[pw] = __fadd_rn(acc[pw], s);
              }
            }
          }
        }
        if (flags & 2u) {
          // ---- the bin row is complete: divide, transpose through shared memory, store ----
          const int cnt_i = (int)(gc.x & 0xffffu) * (int)(gc.x >> 16);
          const float count = __uint_as_float(gc.y);
          const bool pow2 = (cnt_i & (cnt_i - 1)) == 0;
          const float inv = __fdiv_rn(1.0f, count);
          const int cvalid = min(32, p.c_work - c0);
          if (TYPED) {
            // half-precision and / or channel-last output: out of line, the reference's format below
            // keeps the code it always had
            float vals[PWT];
#pragma unroll
            for (int pw = 0; pw < PWT; ++pw) vals[pw] = pow2 ? __fmul_rn(acc[pw], inv) : __fdiv_rn(acc[pw], count);
            sweep_store_typed(p, vals, PWT, stage, roi, ph, c0, cvalid, bins, lane);
          } else {
            __syncwarp();
#pragma unroll
            for (int pw = 0; pw < PWT; ++pw)
              stage[lane * p.pwp + pw] = pow2 ? __fmul_rn(acc[pw], inv) : __fdiv_rn(acc[pw], count);
            __syncwarp();
            float* o = p.out + ((size_t)roi * p.channels + c0) * bins + ph * PWT;
#pragma unroll
            for (int k = 0; k < PWT; ++k) {
              const int f = k * 32 + lane;
              const int c = f / PWT, pw = f - c * PWT;
              if (c < cvalid) {
                const float v = stage[c * p.pwp + pw];
                float* dst = o + c * bins + pw;
                *dst = p.accumulate ? __fadd_rn(*dst, v) : v;
              }
            }
          }
        } else {
          // ---- hand the running sums to the successor item ------------------------------
#pragma unroll
          for (int pw = 0; pw < PWT; ++pw) st_cg(scr + pw * 32, acc[pw]);
          fence_gpu();
          __syncwarp();
          if (lane == 0) {
            const unsigned long long bit = (unsigned long long)idx * (unsigned)p.ncg + (unsigned)cg;
            atomicOr(p.done + (bit >> 5), 1u << (unsigned)(bit & 31u));
          }
        }
      }
      // release the remaining rows of the unit
      release_rows(full, empty, seq0, my_q, nrows, p.ring, p.ring_magic, lane);
    }
    seq0 += (unsigned)nrows;
  }
}

// one tap of the backward: shared-memory read-modify-write by the pixel's only writer
template <bool DET>
__device__ __forceinline__ void tap_add(unsigned addr, float g, float w, float count, float inv, bool pow2) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  if (DET) {
    const float t = __fmul_rn(g, w);                              // dtop * w, then / count (ROIAlign.cpp:143-146)
    v = __fadd_rn(v, pow2 ? __fmul_rn(t, inv) : __fdiv_rn(t, count));
  } else {
    v = fmaf(g, w, v);
  }
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}

// the four taps of one sample when they are four different pixels of this warp's strip
template <bool DET>
__device__ __forceinline__ void taps4(unsigned a1, unsigned a2, unsigned a3, unsigned a4, float g, float w1, float w2,
                                      float w3, float w4, float count, float inv, bool pow2) {
  float v1, v2, v3, v4;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v1) : "r"(a1) : "memory");
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v2) : "r"(a2));
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v3) : "r"(a3));
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v4) : "r"(a4));
  if (DET) {
    const float t1 = __fmul_rn(g, w1), t2 = __fmul_rn(g, w2), t3 = __fmul_rn(g, w3), t4 = __fmul_rn(g, w4);
    v1 = __fadd_rn(v1, pow2 ? __fmul_rn(t1, inv) : __fdiv_rn(t1, count));
    v2 = __fadd_rn(v2, pow2 ? __fmul_rn(t2, inv) : __fdiv_rn(t2, count));
    v3 = __fadd_rn(v3, pow2 ? __fmul_rn(t3, inv) : __fdiv_rn(t3, count));
    v4 = __fadd_rn(v4, pow2 ? __fmul_rn(t4, inv) : __fdiv_rn(t4, count));
  } else {
    v1 = fmaf(g, w1, v1);
    v2 = fmaf(g, w2, v2);
    v3 = fmaf(g, w3, v3);
    v4 = fmaf(g, w4, v4);
  }
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(a1), "f"(v1));
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(a2), "f"(v2));
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(a3), "f"(v3));
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(a4), "f"(v4) : "memory");
}

// ---- the backward: row tiles, strip-owner warps ------------------------------------------------
// dX is cut into tiles of `tr` full-width rows of 32 channels that live in shared memory as
// [row][channel][pitch] (the forward's layout: lane = channel, conflict-free); a CTA takes
// (image, channel group, tile) units from a queue.  Each of its 16 warps OWNS a strip of columns
// of the tile -- every pixel has exactly one writer, so there are no atomics -- and walks the
// image's ROIs in ascending order; for every bin row of a ROI that reaches into the tile it
// fetches the 32 x PW gradients (transposed through shared memory), then visits the samples in
// the reference's order (bin column, sample row, sample column; ROIAlign.cpp:126-157) and adds
// the taps that fall into its strip and the tile's rows.  A finished tile is written once,
// coalesced; the zero-fill of ROIAlign.py:33-34 is fused.
//   DET: every tap is added separately as v += (dtop * (wy * wx)) / count -- each pixel receives
//        exactly the single-thread reference's sequence of additions: bit-identical dX;
//   else: v = fma(dtop / count, wy * wx, v): same order, contracted arithmetic.
// Any sampling grid; the tables are the forward's (pass 1 with compacted column records).
constexpr int kTileWarps = 16;

struct TileParams {
  const float* dy;
  float* dx;
  int num_images, channels, c_work, ncg;
  int height, width, pooled_h, pooled_w;
  int accumulate;
  int pitch, tr, strip;     // floats between channels of a tile row (odd), rows per tile, columns per warp
  int pwp;
  int colcap, rowcap;
  const int* img_start;     // [N + 2] grouped position of the first ROI of every image
  const int* order;         // [R] ROI rows grouped by image, ascending inside an image
  const SweepHdr* hdr;
  const ColRec* cols;
  const RowRec* rows;
  const uint2* phr;
  const uint2* pwr;
  int* unit_counter;
};

template <int PWT, bool DET>
__global__ void __launch_bounds__(kTileWarps * 32, 1) tile_bwd_kernel(const TileParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* tile = reinterpret_cast<float*>(smem_raw);
  const int rowf = 32 * p.pitch;
  float* stage_all = tile + (size_t)p.tr * rowf;
  __shared__ int s_unit;
  ptx::chain_enter();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* stage = stage_all + warp * 32 * p.pwp;
  const int H = p.height, W = p.width, PH = p.pooled_h;
  const int bins = PH * PWT;
  const int tiles = (H + p.tr - 1) / p.tr;
  const int units = p.num_images * p.ncg * tiles;
  const unsigned tile_base = ptx::smem_addr(tile) + (unsigned)(lane * p.pitch * 4);
  const int sx0 = warp * p.strip, sx1 = min(W, sx0 + p.strip);
  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) s_unit = atomicAdd(p.unit_counter, 1);
    __syncthreads();
    const int u = s_unit;
    if (u >= units) break;
    const int cg = u % p.ncg, t = (u / p.ncg) % tiles, n = u / (p.ncg * tiles);
    const int r0 = t * p.tr, r1 = min(H, r0 + p.tr), c0 = cg * 32;
    for (int i = threadIdx.x; i < (r1 - r0) * rowf; i += kTileWarps * 32) tile[i] = 0.0f;
    __syncthreads();
    if (sx0 < sx1) {
      const int j0 = __ldg(p.img_start + n), j1 = __ldg(p.img_start + n + 1);
      // 32 ROIs at a time: lane = ROI for the range tests, then the hits one by one in ascending order
      for (int jb = j0; jb < j1; jb += 32) {
        int roi_l = -1;
        bool hit = false;
        if (jb + lane < j1) {
          roi_l = __ldg(p.order + jb + lane);
          const uint2 rng = __ldg(reinterpret_cast<const uint2*>(p.hdr + roi_l) + 2);    // yr, xr
          hit = (int)(rng.x >> 16) >= r0 && (int)(rng.x & 0xffffu) < r1 && (int)(rng.y >> 16) >= sx0 &&
                (int)(rng.y & 0xffffu) < sx1;
        }
        unsigned hits = __ballot_sync(0xffffffffu, hit);
        while (hits) {
          const int b = __ffs(hits) - 1;
          hits &= hits - 1;
          const int roi = __shfl_sync(0xffffffffu, roi_l, b);
          // bin columns (lane = pw): record range and tap columns; which of them reach into the strip
          uint2 cr = make_uint2(0u, 0u);
          if (lane < PWT) cr = __ldg(p.pwr + (size_t)roi * 16 + lane);
          const bool chit = (cr.x & 0xffffu) < (cr.x >> 16) && (int)(cr.y >> 16) >= sx0 && (int)(cr.y & 0xffffu) < sx1;
          const unsigned cmask = __ballot_sync(0xffffffffu, chit);
          if (!cmask) continue;
          const uint4 h0 = __ldg(reinterpret_cast<const uint4*>(p.hdr + roi));           // roi, batch, grid, count
          const float count = __uint_as_float(h0.w);
          const int cnt_i = (int)(h0.z & 0xffffu) * (int)(h0.z >> 16);
          const bool pow2 = (cnt_i & (cnt_i - 1)) == 0;
          const float inv = __fdiv_rn(1.0f, count);
          const ColRec* cols = p.cols + (size_t)roi * p.colcap;
          const RowRec* rows = p.rows + (size_t)roi * p.rowcap;
          for (int ph0 = 0; ph0 < PH; ph0 += 32) {
            // bin rows (lane = ph): which of them reach into the tile
            uint2 rr0 = make_uint2(0u, 0u);
            if (ph0 + lane < PH) rr0 = __ldg(p.phr + (size_t)roi * PH + ph0 + lane);
            const bool rhit = (rr0.x & 0xffffu) < (rr0.x >> 16) && (int)(rr0.y >> 16) >= r0 && (int)(rr0.y & 0xffffu) < r1;
            unsigned rmask = __ballot_sync(0xffffffffu, rhit);
            while (rmask) {
              const int bl = __ffs(rmask) - 1;
              rmask &= rmask - 1;
              const int ph = ph0 + bl;
              const unsigned ox = __shfl_sync(0xffffffffu, rr0.x, bl);
              const int o0 = (int)(ox & 0xffffu), o1 = (int)(ox >> 16);
              // the 32 x PW gradients of this bin row, [channel][bin column] -> lane = channel
              float g[PWT];
              {
                const float* src = p.dy + ((size_t)roi * p.channels + c0) * bins + ph * PWT;
                const int cvalid = min(32, p.c_work - c0);
                __syncwarp();
#pragma unroll
                for (int k = 0; k < PWT; ++k) {
                  const int f = k * 32 + lane;
                  const int c = f / PWT, pw = f - c * PWT;
                  if (c < cvalid) stage[c * p.pwp + pw] = __ldg(src + c * bins + pw);
                }
                __syncwarp();
#pragma unroll
                for (int pw = 0; pw < PWT; ++pw) {
                  const float v = stage[lane * p.pwp + pw];
                  g[pw] = DET ? v : (pow2 ? __fmul_rn(v, inv) : __fdiv_rn(v, count));
                }
              }
#pragma unroll
              for (int pw = 0; pw < PWT; ++pw) {
                const unsigned kx = __shfl_sync(0xffffffffu, cr.x, pw);
                if (!((cmask >> pw) & 1u)) continue;
                const int k0 = (int)(kx & 0xffffu), k1 = (int)(kx >> 16);
                const float gp = g[pw];
                for (int o = o0; o < o1; ++o) {
                  const RowRec rr = __ldg(rows + o);
                  const int ylo = (int)(rr.x & 0xffffu), yhi = (int)(rr.x >> 16);
                  if (yhi < r0 || ylo >= r1) continue;
                  const float ly = __uint_as_float(rr.y), hy = __fsub_rn(1.0f, ly);
                  const bool lo_in = ylo >= r0, hi_in = yhi < r1;
                  const unsigned row_lo = tile_base + (unsigned)((ylo - r0) * rowf * 4);
                  const unsigned row_hi = tile_base + (unsigned)((yhi - r0) * rowf * 4);
                  for (int k = k0; k < k1; ++k) {
                    const ColRec c = __ldg(cols + k);
                    const int xlo = (int)(c.x & 0xffffu), xhi = (int)(c.x >> 16);
                    if (xhi < sx0 || xlo >= sx1) continue;
                    const float lx = __uint_as_float(c.y), hx = __fsub_rn(1.0f, lx);
                    const bool xl_in = xlo >= sx0, xh_in = xhi < sx1;
                    const float w1 = __fmul_rn(hy, hx), w2 = __fmul_rn(hy, lx), w3 = __fmul_rn(ly, hx), w4 = __fmul_rn(ly, lx);
                    if (lo_in && hi_in && xl_in && xh_in && xlo != xhi && ylo != yhi) {
                      // four different pixels, all ours: the read-modify-writes overlap
                      taps4<DET>(row_lo + (unsigned)xlo * 4u, row_lo + (unsigned)xhi * 4u, row_hi + (unsigned)xlo * 4u,
                                 row_hi + (unsigned)xhi * 4u, gp, w1, w2, w3, w4, count, inv, pow2);
                    } else {
                      // taps in the reference's order: (lo,lo) (lo,hi) (hi,lo) (hi,hi), ROIAlign.cpp:149-156
                      if (lo_in && xl_in) tap_add<DET>(row_lo + (unsigned)xlo * 4u, gp, w1, count, inv, pow2);
                      if (lo_in && xh_in) tap_add<DET>(row_lo + (unsigned)xhi * 4u, gp, w2, count, inv, pow2);
                      if (hi_in && xl_in) tap_add<DET>(row_hi + (unsigned)xlo * 4u, gp, w3, count, inv, pow2);
                      if (hi_in && xh_in) tap_add<DET>(row_hi + (unsigned)xhi * 4u, gp, w4, count, inv, pow2);
                    }
                  }
                }
              }
            }
          }
        }
      }
    }
    __syncthreads();
    // ---- write the tile: every (row, channel) is one coalesced run of W floats ----------------
    const int cvalid = min(32, p.c_work - c0);
    for (int rc = warp; rc < (r1 - r0) * cvalid; rc += kTileWarps) {
      const int rr = rc / cvalid, c = rc - rr * cvalid;
      const float* srow = tile + (size_t)rr * rowf + c * p.pitch;
      float* drow = p.dx + (((size_t)n * p.channels + c0 + c) * H + r0 + rr) * W;
      for (int x = lane; x < W; x += 32) drow[x] = p.accumulate ? __fadd_rn(drow[x], srow[x]) : srow[x];
    }
  }
}

// ---- host side ------------------------------------------------------------------------------
struct SweepLayout {
  size_t zero_bytes;     // leading part of the workspace that is cleared per call
  size_t cnt, start, unit_rows, unit_counter, done, hdr, cols, rows, items, scratch, total;
  long long item_cap;
  int m;                 // N * segs * H
};

SweepLayout sweep_layout(const RoiAlignArgs& a, const SweepPlan& plan) {
  SweepLayout L;
  const int c_work = a.c_count > 0 ? a.c_count : a.channels;
  const int ncg = (c_work + 31) / 32;
  L.m = a.num_images * plan.segs * a.height;
  const long long per_roi = a.sampling_ratio > 0 ? (long long)a.pooled_h * a.sampling_ratio
                                                 : (long long)a.pooled_h + a.height + 2;
  L.item_cap = (long long)a.num_rois * per_roi;
  size_t off = 0;
  L.cnt = off;          off = align_up(off + (size_t)L.m * 4, 16);
  L.unit_rows = off;    off = align_up(off + (size_t)a.num_images * plan.segs * 4, 16);
  L.unit_counter = off; off = align_up(off + 16, 16);
  L.done = off;         off = align_up(off + (size_t)((L.item_cap * ncg + 31) / 32) * 4, 256);
  L.zero_bytes = off;
  L.start = off;        off = align_up(off + (size_t)(L.m + 1) * 4, 256);
  L.hdr = off;          off = align_up(off + (size_t)a.num_rois * sizeof(SweepHdr), 256);
  L.cols = off;
  off = align_up(off + (size_t)a.num_rois * plan.colcap * (a.sampling_ratio == 2 ? sizeof(ColRec4) : sizeof(ColRec)), 256);
  L.rows = off;         off = align_up(off + (size_t)a.num_rois * plan.rowcap * sizeof(RowRec), 256);
  L.items = off;        off = align_up(off + (size_t)L.item_cap * sizeof(ItemRec), 256);
  L.scratch = off;
  off = align_up(off + (size_t)a.num_rois * a.pooled_h * ncg * a.pooled_w * 32 * sizeof(float), 256);
  L.total = off;
  return L;
}

template <int PWT>
cudaError_t launch_main(int gwt, dim3 grid, size_t smem, cudaStream_t stream, const SweepParams& p) {
  auto go = [&](auto kernel) -> cudaError_t {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    count_launch();
    return chained_launch(kernel, grid, dim3(kSweepThreads), smem, stream, p);
  };
  if ((p.top_dtype | p.top_layout) != 0)
    return gwt == 2 ? go(sweep_fwd_kernel<PWT, 2, true>) : go(sweep_fwd_kernel<PWT, 0, true>);
  return gwt == 2 ? go(sweep_fwd_kernel<PWT, 2, false>) : go(sweep_fwd_kernel<PWT, 0, false>);
}

}  // namespace

bool sweep_supported(const RoiAlignArgs& a, const void* plane_ptr, int smem_optin, int num_sms, SweepPlan* plan) {
  if (a.num_images <= 0 || a.num_images > 4096 || a.num_rois <= 0) return false;
  if (a.width > 65535 || a.height > kMaxPlaneH || a.pooled_w > kMaxPooledW || a.pooled_h > 16383) return false;
  if ((long long)a.pooled_h * (a.sampling_ratio > 0 ? a.sampling_ratio : 1024) > 65535) return false;
  (void)plane_ptr;
  plan->pitch = (a.width + 1) | 1;       // odd: the 32 channels of a pixel hit the 32 banks; >= 1 pad word
  plan->rowf = 32 * plan->pitch;
  plan->pwp = a.pooled_w | 1;
  const size_t fixed = (size_t)kConsumers * 32 * plan->pwp * 4 + 2 * kMaxRing * 8 + 64;
  const size_t row_bytes = (size_t)plan->rowf * 4;
  if ((size_t)smem_optin < fixed + kMinRing * row_bytes) return false;
  int ring = (int)(((size_t)smem_optin - fixed) / row_bytes);
  if (ring > kMaxRing) ring = kMaxRing;
  if (const char* e = getenv("MOBULA_SWEEP_RING")) ring = max(kMinRing, min(atoi(e), ring));   // experiments
  plan->ring = ring;
  plan->span = ring >= 8 ? 4 : (ring >= 6 ? 4 : 3);
  if (plan->span > ring - 2) plan->span = ring - 2;
  if (const char* e = getenv("MOBULA_SWEEP_SPAN")) plan->span = max(2, min(atoi(e), ring - 2));   // experiments
  plan->smem = fixed + (size_t)ring * row_bytes;
  // row segments per image: fill the SMs without paying too much for the lead-in rows
  const int c_work = a.c_count > 0 ? a.c_count : a.channels;
  const int ncg = (c_work + 31) / 32;
  const long long base_units = (long long)a.num_images * ncg;
  int best = 1;
  double best_score = -1.0;
  for (int s = 1; s <= 32 && s * 8 <= a.height; ++s) {
    const long long units = base_units * s;
    const long long rounds = (units + num_sms - 1) / num_sms;
    const double fill = (double)units / (double)(rounds * num_sms);
    const double lead = 1.0 / (1.0 + 6.0 * (s - 1) / (double)a.height);
    const double score = fill * lead;
    if (score > best_score + 1e-9) {
      best_score = score;
      best = s;
    }
  }
  plan->segs = best;
  if (const char* e = getenv("MOBULA_SWEEP_SEGS")) plan->segs = max(1, min(atoi(e), a.height));   // experiments
  plan->colcap = a.sampling_ratio > 0 ? ((a.pooled_w * a.sampling_ratio + 1) & ~1)
                                      : ((max(a.pooled_w, 2 * a.width + 3) + 2) & ~1);
  plan->rowcap = a.sampling_ratio > 0 ? a.pooled_h * a.sampling_ratio : 2 * a.height + 3 + a.pooled_h;
  return true;
}

size_t sweep_fwd_workspace(const RoiAlignArgs& a, const SweepPlan& plan) { return sweep_layout(a, plan).total; }

cudaError_t launch_fwd_sweep(const RoiAlignArgs& a, const SweepPlan& plan, void* workspace, const float* data,
                             float* out, int accumulate, int num_sms, cudaStream_t stream) {
  const SweepLayout L = sweep_layout(a, plan);
  char* ws = static_cast<char*>(workspace);
  const int c_work = a.c_count > 0 ? a.c_count : a.channels;
  SweepParams p = {};
  p.rois = a.rois;
  p.out = out;
  p.num_rois = a.num_rois;
  p.num_images = a.num_images;
  p.channels = a.channels;
  p.c_work = c_work;
  p.ncg = (c_work + 31) / 32;
  p.height = a.height;
  p.width = a.width;
  p.pooled_h = a.pooled_h;
  p.pooled_w = a.pooled_w;
  p.scale = a.scale;
  p.sampling_ratio = a.sampling_ratio;
  p.aligned = a.aligned;
  p.accumulate = accumulate;
  p.top_dtype = a.top_dtype;
  p.top_layout = a.top_layout;
  p.data = data;
  p.pitch = plan.pitch;
  p.fill_mode = (a.width % 4 == 0 && (reinterpret_cast<uintptr_t>(data) & 15u) == 0) ? 0 : 1;
  if (const char* e = getenv("MOBULA_SWEEP_FILL")) p.fill_mode = atoi(e) ? 1 : p.fill_mode;   // experiments
  p.rowf = plan.rowf;
  p.ring = plan.ring;
  p.span = plan.span;
  p.ring_magic = (unsigned)(0x100000000ull / (unsigned)plan.ring) + 1u;
  p.pwp = plan.pwp;
  p.segs = plan.segs;
  p.colcap = plan.colcap;
  p.rowcap = plan.rowcap;
  p.compact = a.sampling_ratio == 2 ? 0 : 1;
  p.cnt = reinterpret_cast<int*>(ws + L.cnt);
  p.start = reinterpret_cast<int*>(ws + L.start);
  p.unit_rows = reinterpret_cast<int*>(ws + L.unit_rows);
  p.unit_counter = reinterpret_cast<int*>(ws + L.unit_counter);
  p.done = reinterpret_cast<unsigned*>(ws + L.done);
  p.hdr = reinterpret_cast<SweepHdr*>(ws + L.hdr);
  p.cols = reinterpret_cast<ColRec*>(ws + L.cols);
  p.rows = reinterpret_cast<RowRec*>(ws + L.rows);
  p.items = reinterpret_cast<ItemRec*>(ws + L.items);
  p.scratch = reinterpret_cast<float*>(ws + L.scratch);

  static const bool debug = getenv("MOBULA_SWEEP_DEBUG") != nullptr;
  auto checkpoint = [&](const char* what) -> cudaError_t {
    if (!debug) return cudaSuccess;
    const cudaError_t s = cudaStreamSynchronize(stream);
    fprintf(stderr, "[sweep] %s: %s\n", what, cudaGetErrorString(s));
    return s;
  };
  cudaError_t e = cudaMemsetAsync(ws, 0, L.zero_bytes, stream);
  if (e != cudaSuccess) return e;
  if ((e = checkpoint("memset")) != cudaSuccess) return e;
  const dim3 rgrid((a.num_rois + 7) / 8), rblock(256);
  count_launch();
  e = chained_launch(sweep_roi_kernel<1>, rgrid, rblock, 0, stream, p);
  if (e != cudaSuccess) return e;
  if ((e = checkpoint("pass 1")) != cudaSuccess) return e;
  count_launch();
  e = chained_launch(sweep_scan_kernel, dim3(1), dim3(1024), 0, stream, p.cnt, p.start, L.m);
  if (e != cudaSuccess) return e;
  if ((e = checkpoint("scan")) != cudaSuccess) return e;
  count_launch();
  e = chained_launch(sweep_roi_kernel<2>, rgrid, rblock, 0, stream, p);
  if (e != cudaSuccess) return e;
  if ((e = checkpoint("pass 2")) != cudaSuccess) return e;
  if (debug)
    fprintf(stderr, "[sweep] pitch %d rowf %d ring %d span %d segs %d colcap %d smem %zu units %lld\n",
            plan.pitch, plan.rowf, plan.ring, plan.span, plan.segs, plan.colcap, plan.smem,
            (long long)a.num_images * plan.segs * p.ncg);
  const long long units = (long long)a.num_images * plan.segs * p.ncg;
  const dim3 grid((unsigned)(units < num_sms ? units : num_sms));
  const int gwt = a.sampling_ratio == 2 ? 2 : 0;
  switch (a.pooled_w) {
    case 1: return launch_main<1>(gwt, grid, plan.smem, stream, p);
    case 2: return launch_main<2>(gwt, grid, plan.smem, stream, p);
    case 3: return launch_main<3>(gwt, grid, plan.smem, stream, p);
    case 4: return launch_main<4>(gwt, grid, plan.smem, stream, p);
    case 5: return launch_main<5>(gwt, grid, plan.smem, stream, p);
    case 6: return launch_main<6>(gwt, grid, plan.smem, stream, p);
    case 7: return launch_main<7>(gwt, grid, plan.smem, stream, p);
    case 8: return launch_main<8>(gwt, grid, plan.smem, stream, p);
    case 9: return launch_main<9>(gwt, grid, plan.smem, stream, p);
    case 10: return launch_main<10>(gwt, grid, plan.smem, stream, p);
    case 11: return launch_main<11>(gwt, grid, plan.smem, stream, p);
    case 12: return launch_main<12>(gwt, grid, plan.smem, stream, p);
    case 13: return launch_main<13>(gwt, grid, plan.smem, stream, p);
    case 14: return launch_main<14>(gwt, grid, plan.smem, stream, p);
    case 15: return launch_main<15>(gwt, grid, plan.smem, stream, p);
    default: return launch_main<16>(gwt, grid, plan.smem, stream, p);
  }
}

// ---- backward (row tiles) ----------------------------------------------------------------------
namespace {

struct TileLayout {
  size_t counter, img_start, order, hdr, cols, rows, phr, pwr, total;
};

TileLayout tile_layout(const RoiAlignArgs& a, const TilePlan& plan) {
  TileLayout L;
  size_t off = 0;
  L.counter = off;   off = align_up(off + 16, 16);
  L.img_start = off; off = align_up(off + (size_t)(a.num_images + 2) * 4, 16);
  L.order = off;     off = align_up(off + (size_t)a.num_rois * 4, 256);
  L.hdr = off;       off = align_up(off + (size_t)a.num_rois * sizeof(SweepHdr), 256);
  L.cols = off;      off = align_up(off + (size_t)a.num_rois * plan.colcap * sizeof(ColRec), 256);
  L.rows = off;      off = align_up(off + (size_t)a.num_rois * plan.rowcap * sizeof(RowRec), 256);
  L.phr = off;       off = align_up(off + (size_t)a.num_rois * a.pooled_h * sizeof(uint2), 256);
  L.pwr = off;       off = align_up(off + (size_t)a.num_rois * 16 * sizeof(uint2), 256);
  L.total = off;
  return L;
}

template <int PWT>
cudaError_t launch_tile(bool det, dim3 grid, size_t smem, cudaStream_t stream, const TileParams& p) {
  auto go = [&](auto kernel) -> cudaError_t {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    count_launch();
    return chained_launch(kernel, grid, dim3(kTileWarps * 32), smem, stream, p);
  };
  return det ? go(tile_bwd_kernel<PWT, true>) : go(tile_bwd_kernel<PWT, false>);
}

}  // namespace

bool tile_supported(const RoiAlignArgs& a, int smem_optin, int num_sms, TilePlan* plan) {
  if (a.num_images <= 0 || a.num_images > 4096 || a.num_rois < 0) return false;
  if (a.width > 65535 || a.height > kMaxPlaneH || a.pooled_w > kMaxPooledW || a.pooled_h > 16383) return false;
  if ((long long)a.pooled_h * (a.sampling_ratio > 0 ? a.sampling_ratio : 1024) > 65535) return false;
  plan->pitch = (a.width + 1) | 1;
  plan->pwp = a.pooled_w | 1;
  const size_t fixed = (size_t)kTileWarps * 32 * plan->pwp * 4 + 256;
  const size_t row_bytes = (size_t)32 * plan->pitch * 4;
  if ((size_t)smem_optin < fixed + 2 * row_bytes) return false;
  int tr = (int)(((size_t)smem_optin - fixed) / row_bytes);
  // enough tiles to keep every SM busy, as tall as that allows (fewer ROI scans, fewer bin rows cut)
  const int c_work = a.c_count > 0 ? a.c_count : a.channels;
  const long long per_row_units = (long long)a.num_images * ((c_work + 31) / 32);
  const long long want_tiles = (2LL * num_sms + per_row_units - 1) / per_row_units;
  const int balanced = (int)max(2LL, (a.height + want_tiles - 1) / want_tiles);
  if (tr > balanced) tr = balanced;
  if (tr > a.height) tr = a.height;
  if (tr > 32) tr = 32;
  plan->tr = tr;
  plan->strip = (a.width + kTileWarps - 1) / kTileWarps;
  plan->smem = fixed + (size_t)tr * row_bytes;
  plan->colcap = a.sampling_ratio > 0 ? ((a.pooled_w * a.sampling_ratio + 1) & ~1)
                                      : ((max(a.pooled_w, 2 * a.width + 3) + 2) & ~1);
  plan->rowcap = a.sampling_ratio > 0 ? a.pooled_h * a.sampling_ratio : 2 * a.height + 3 + a.pooled_h;
  return true;
}

size_t tile_bwd_workspace(const RoiAlignArgs& a, const TilePlan& plan) { return tile_layout(a, plan).total; }

cudaError_t launch_bwd_tile(const RoiAlignArgs& a, const TilePlan& plan, void* workspace, const float* dy,
                            float* dx, int accumulate, int deterministic, int num_sms, cudaStream_t stream) {
  const TileLayout L = tile_layout(a, plan);
  char* ws = static_cast<char*>(workspace);
  const int c_work = a.c_count > 0 ? a.c_count : a.channels;
  cudaError_t e = cudaMemsetAsync(ws + L.counter, 0, 16, stream);
  if (e != cudaSuccess) return e;
  int* img_start = reinterpret_cast<int*>(ws + L.img_start);
  int* order = reinterpret_cast<int*>(ws + L.order);
  e = launch_roi_group(a, img_start, order, stream);
  if (e != cudaSuccess) return e;
  if (a.num_rois > 0) {
    SweepParams sp = {};
    sp.rois = a.rois;
    sp.num_rois = a.num_rois;
    sp.num_images = a.num_images;
    sp.channels = a.channels;
    sp.c_work = c_work;
    sp.ncg = (c_work + 31) / 32;
    sp.height = a.height;
    sp.width = a.width;
    sp.pooled_h = a.pooled_h;
    sp.pooled_w = a.pooled_w;
    sp.scale = a.scale;
    sp.sampling_ratio = a.sampling_ratio;
    sp.aligned = a.aligned;
    sp.accumulate = 1;
    sp.segs = 1;
    sp.span = 2;
    sp.colcap = plan.colcap;
    sp.rowcap = plan.rowcap;
    sp.compact = 1;
    sp.tables_only = 1;
    sp.hdr = reinterpret_cast<SweepHdr*>(ws + L.hdr);
    sp.cols = reinterpret_cast<ColRec*>(ws + L.cols);
    sp.rows = reinterpret_cast<RowRec*>(ws + L.rows);
    sp.phr = reinterpret_cast<uint2*>(ws + L.phr);
    sp.pwr = reinterpret_cast<uint2*>(ws + L.pwr);
    count_launch();
    e = chained_launch(sweep_roi_kernel<1>, dim3((a.num_rois + 7) / 8), dim3(256), 0, stream, sp);
    if (e != cudaSuccess) return e;
  }
  TileParams p = {};
  p.dy = dy;
  p.dx = dx;
  p.num_images = a.num_images;
  p.channels = a.channels;
  p.c_work = c_work;
  p.ncg = (c_work + 31) / 32;
  p.height = a.height;
  p.width = a.width;
  p.pooled_h = a.pooled_h;
  p.pooled_w = a.pooled_w;
  p.accumulate = accumulate;
  p.pitch = plan.pitch;
  p.tr = plan.tr;
  p.strip = plan.strip;
  p.pwp = plan.pwp;
  p.colcap = plan.colcap;
  p.rowcap = plan.rowcap;
  p.img_start = img_start;
  p.order = order;
  p.hdr = reinterpret_cast<const SweepHdr*>(ws + L.hdr);
  p.cols = reinterpret_cast<const ColRec*>(ws + L.cols);
  p.rows = reinterpret_cast<const RowRec*>(ws + L.rows);
  p.phr = reinterpret_cast<const uint2*>(ws + L.phr);
  p.pwr = reinterpret_cast<const uint2*>(ws + L.pwr);
  p.unit_counter = reinterpret_cast<int*>(ws + L.counter);
  const long long units = (long long)a.num_images * p.ncg * ((a.height + plan.tr - 1) / plan.tr);
  const dim3 grid((unsigned)(units < num_sms ? units : num_sms));
  const bool det = deterministic != 0;
  switch (a.pooled_w) {
    case 1: return launch_tile<1>(det, grid, plan.smem, stream, p);
    case 2: return launch_tile<2>(det, grid, plan.smem, stream, p);
    case 3: return launch_tile<3>(det, grid, plan.smem, stream, p);
    case 4: return launch_tile<4>(det, grid, plan.smem, stream, p);
    case 5: return launch_tile<5>(det, grid, plan.smem, stream, p);
    case 6: return launch_tile<6>(det, grid, plan.smem, stream, p);
    case 7: return launch_tile<7>(det, grid, plan.smem, stream, p);
    case 8: return launch_tile<8>(det, grid, plan.smem, stream, p);
    case 9: return launch_tile<9>(det, grid, plan.smem, stream, p);
    case 10: return launch_tile<10>(det, grid, plan.smem, stream, p);
    case 11: return launch_tile<11>(det, grid, plan.smem, stream, p);
    case 12: return launch_tile<12>(det, grid, plan.smem, stream, p);
    case 13: return launch_tile<13>(det, grid, plan.smem, stream, p);
    case 14: return launch_tile<14>(det, grid, plan.smem, stream, p);
    case 15: return launch_tile<15>(det, grid, plan.smem, stream, p);
    default: return launch_tile<16>(det, grid, plan.smem, stream, p);
  }
}

}  // namespace mobula_b200
