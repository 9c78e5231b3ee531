// "pull" backward: every dX pixel is produced by one warp that walks the pixel's own list of
// contributions, lane = channel.
//
// The scatter of roi_align_backward_kernel (/root/reference/opzoo/ROIAlign/ROIAlign.cpp:78-161)
// is turned around.  Which (ROI, bin) pairs reach a pixel, and with which bilinear weight, does
// not depend on the channel, so it is worked out ONCE per call as a list per pixel:
//     entry = { (roi * PH + ph) * PW + pw ,  weight }
// and the main kernel is then nothing but
//     dX[n, c, y, x] = sum over the entries of (y, x):  weight * dy[roi, c, ph, pw]
// with the 32 lanes of a warp on 32 * CPL different channels: entry loads and addresses are
// warp-uniform, every gradient load is one coalesced line of the channel-last copy of dy, the
// sums live in registers, there are no atomics, no shared-memory read-modify-writes and no
// zero-fill pass (a pixel without entries is written as 0; ROIAlign.py:33-34 fused).  Every
// pixel is written exactly once, coalesced, after a transpose through shared memory.
//
//   DET = false: the taps of one (pixel, bin) are merged into one entry (weights summed, 1/count
//                folded in): within 1e-5 of the reference, reproducible run to run;
//   DET = true : one entry per bilinear tap, sorted by (roi, ph, pw, iy, ix, tap) = the order in
//                which the single-thread reference adds into that pixel (ROIAlign.cpp:126-157),
//                each added as v += (dtop * w) / count: bit-identical dX.
//
// Tables (all channel independent), built per call by small kernels:
//   geometry  one warp per ROI: the taps of each axis, sorted by pixel, and a byte index
//             "pixel -> first tap" per axis (pull_geom_kernel);
//   row mask  per feature-map row, one bit per ROI of that image whose rows reach it;
//   count     one thread per pixel counts its entries; scan; fill writes them (no atomics:
//             a pixel's list is written by one thread in the final order).
// Fixed sampling ratios only (an adaptive grid has up to H * W taps per ROI; those run on the
// dense kernels, roialign_dense.cu).
#include <cstdio>
#include <cstdlib>

#include "roialign_kernels.h"
#include "roialign_common.cuh"
#include "ptx_sm100.cuh"

namespace mobula_b200 {

namespace {

constexpr int kMaxFixedSamples = 64;          // fixed sampling ratio: samples per axis, PH * G and PW * G
constexpr int kMaxAdaptiveExtent = 2048;      // adaptive grids: plane height / width the geometry kernel's scratch covers
constexpr int kMaxGeomWarps = 4;

struct PullParams {
  const float* rois;
  int num_rois, num_images;
  int channels;          // channel stride of dy / dx
  int c_work;            // channels this call works on
  int cw;                // channels per row of the channel-last copy of dy (c_work rounded up to 4)
  int height, width, pooled_h, pooled_w;
  int grid;              // sampling ratio; <= 0: adaptive, ceil(bin) samples per bin and axis (ROIAlign.cpp:47-51)
  float scale;
  int aligned;
  int accumulate;
  int caps;              // valid samples per ROI and axis the tables have room for
  int capy, capx;        // taps per ROI and axis
  int geom_warps;        // warps per block of the geometry kernel
  int tw;                // row-mask words per feature-map row
  long long pixels;      // N * H * W
  const int* img_start;  // [N + 2]
  const int* order;      // [R] ROI rows grouped by image, ascending inside an image
  uint4* gbox;           // [R] by grouped position: rows min | max << 16, columns min | max << 16,
                         // grid_h | grid_w << 16, count (float bits)
  unsigned short* yidx;  // [R][H + 2] by grouped position: first tap of row y at [y], for ymin <= y <= ymax + 1
  unsigned short* xidx;  // [R][W + 2]
  uint2* ytap;           // [R][capy] taps sorted by row: x = row | high << 12 | i << 13 | p << 23 (sample i of
                         // bin p; merged taps: row | p << 23), y = weight
  uint2* xtap;           // [R][capx]
  unsigned* rowmask;     // [H][tw]
  uint2* ptr;            // [pixels] first entry and number of entries of every pixel
  unsigned* cursor;      // entries handed out so far (zeroed by the geometry kernel)
  unsigned long long* need;   // entries this call needs: sum over the ROIs of row taps x column taps
  void* entries;         // uint2: x = byte offset of the bin's channel row in dyt, ((roi * PH + ph) * PW + pw) * cw * 4
                         // (a multiple of 16) | pixel & 7 | 8 on the last entry of the pixel; y = weight.
                         // Deterministic mode on adaptive grids: uint4, z = count, w = 1 / count if that is exact, else 0
  unsigned entry_cap;    // entries the array has room for
  float* dyt;            // [R * PH * PW][cw] channel-last fp32 copy of dy
  const void* dy;
  int dy_dtype, dy_layout;   // RoiAlignArgs::top_dtype / top_layout of dy
  int widen_flat;            // copy blocks run pull_widen_block (channel-last half-precision dy, whole rows)
  int tables_q;              // list / copy grid: cap of the interleaving period (0 = none), see pull_tables_kernel
  float* dx;
};

// first row-mask word of image n: its ROIs sit at grouped positions [img_start[n], img_start[n + 1])
// and use the words (position >> 5) + n, so that two images never share a word
__device__ __forceinline__ int mask_word0(const int* __restrict__ img_start, int n) {
  return (__ldg(img_start + n) >> 5) + n;
}

// ---- geometry: one warp per ROI (grouped position) ------------------------------------------
// Scratch of one warp in shared memory, for `caps` samples.
struct AxisScratch {
  unsigned short* lo;    // [caps] pixel of the low tap of valid sample j (samples in ascending order)
  unsigned short* hi;    // [caps]
  unsigned* info;        // [caps] i << 13 | p << 23
  float* frac;           // [caps] weight of the high tap
  unsigned* skey;        // [2 caps] sorted taps: x word of the tap record
  float* sw;             // [2 caps]
};

__host__ __device__ inline size_t axis_scratch_bytes(int caps) { return (size_t)caps * 28 + 64; }

__device__ __forceinline__ AxisScratch carve_scratch(unsigned char* base, int caps) {
  AxisScratch a;
  a.skey = reinterpret_cast<unsigned*>(base);
  a.sw = reinterpret_cast<float*>(a.skey + 2 * caps);
  a.info = reinterpret_cast<unsigned*>(a.sw + 2 * caps);
  a.frac = reinterpret_cast<float*>(a.info + caps);
  a.lo = reinterpret_cast<unsigned short*>(a.frac + caps);
  a.hi = a.lo + caps;
  return a;
}

// One axis: the taps of the ROI's samples inside the plane, sorted by (pixel, sample, low/high)
// [DET] or merged per (pixel, bin) [!DET], and the index "pixel -> first tap".  Returns the number
// of taps written; *range = min | max << 16 of the touched pixels.
//
// Sample positions ascend with the sample index, hence so do the low and the high pixels, and the
// position of a tap in the sorted order is a count of two prefixes:
//   low tap of sample j : j + #{j' < j : hi[j'] <= lo[j]}          (low taps before it + high taps before it)
//   high tap of sample j: j + #{j' : lo[j'] < hi[j]} (>= j + 1)    (high taps before it + low taps before it)
// both found by binary search.  Should rounding ever break the ascent (bins narrower than an ulp
// of the coordinate), the warp falls back to counting smaller keys one by one.
template <bool DET>
__device__ int pull_axis(float start, float bin, int pooled, int grid, int extent, float wscale, int caps, int lane,
                         const AxisScratch& sc, uint2* __restrict__ taps, unsigned short* __restrict__ idx,
                         unsigned* range, bool absolute) {
  const int ns = pooled * grid;
  // ---- the valid samples, in order ----
  int nv = 0;
  for (int s0 = 0; s0 < ns; s0 += 32) {
    const int s = s0 + lane;
    AxisTap t;
    t.valid = false;
    int p = 0, i = 0;
    if (s < ns) {
      p = s / grid;
      i = s - p * grid;
      t = axis_decode(sample_pos(start, p, bin, i, grid), extent);
    }
    const unsigned m = __ballot_sync(0xffffffffu, t.valid);
    if (t.valid) {
      const int j = nv + __popc(m & ((1u << lane) - 1u));
      if (j < caps) {
        sc.lo[j] = (unsigned short)t.lo;
        sc.hi[j] = (unsigned short)t.hi;
        sc.info[j] = ((unsigned)i << 13) | ((unsigned)p << 23);
        sc.frac[j] = t.wl;
      }
    }
    nv += __popc(m);
  }
  nv = min(nv, caps);
  __syncwarp();
  *range = 0xffffu;
  if (nv == 0) return 0;
  // ---- ascending? ----
  bool bad = false;
  for (int j = lane; j + 1 < nv; j += 32) bad |= sc.lo[j] > sc.lo[j + 1] || sc.hi[j] > sc.hi[j + 1];
  const bool ascending = !__any_sync(0xffffffffu, bad);
  // ---- sorted position of every tap ----
  for (int j = lane; j < nv; j += 32) {
    const unsigned lo = sc.lo[j], hi = sc.hi[j];
    const float f = sc.frac[j];
    const unsigned base = sc.info[j];
    int ra, rb;
    if (ascending) {
      int a = 0, b = j;                        // first j' in [0, j) with hi[j'] > lo
      while (a < b) {
        const int mid = (a + b) >> 1;
        if (sc.hi[mid] > lo) b = mid;
        else a = mid + 1;
      }
      ra = j + a;
      a = j + 1;
      b = nv;                                  // first j' in (j, nv) with lo[j'] >= hi
      while (a < b) {
        const int mid = (a + b) >> 1;
        if (sc.lo[mid] >= hi) b = mid;
        else a = mid + 1;
      }
      rb = j + a;
    } else {
      ra = rb = 0;
      for (int k = 0; k < nv; ++k) {
        const unsigned klo = sc.lo[k], khi = sc.hi[k];
        // keys: (pixel, sample, high)
        ra += (klo < lo || (klo == lo && k < j)) ? 1 : 0;            // low taps before my low tap
        ra += (khi < lo || (khi == lo && k < j)) ? 1 : 0;            // high taps before my low tap
        rb += (klo < hi || (klo == hi && k <= j)) ? 1 : 0;           // low taps before my high tap
        rb += (khi < hi || (khi == hi && k < j)) ? 1 : 0;            // high taps before my high tap
      }
    }
    sc.skey[ra] = lo | base;
    sc.sw[ra] = __fsub_rn(1.0f, f);            // weight of the low tap
    sc.skey[rb] = hi | base | (1u << 12);
    sc.sw[rb] = f;
  }
  __syncwarp();
  const int nt = 2 * nv;
  int nfinal = nt;
  if (DET) {
    for (int e = lane; e < nt; e += 32) taps[e] = make_uint2(sc.skey[e], __float_as_uint(sc.sw[e]));
  } else {
    // runs of one (pixel, bin) become one tap with the summed weight; the merged x words go back
    // to the front of skey (never ahead of the reader: a run is at least one tap long)
    int out_base = 0;
    for (int e0 = 0; e0 < nt; e0 += 32) {
      const int e = e0 + lane;
      bool head = false;
      unsigned id = 0;
      float w = 0.0f;
      if (e < nt) {
        id = sc.skey[e] & 0xff800fffu;         // pixel and bin
        head = e == 0 || (sc.skey[e - 1] & 0xff800fffu) != id;
        if (head) {
          w = sc.sw[e];
          for (int k = e + 1; k < nt && (sc.skey[k] & 0xff800fffu) == id; ++k) w += sc.sw[k];
        }
      }
      const unsigned heads = __ballot_sync(0xffffffffu, head);
      __syncwarp();
      if (head) {
        const int pos = out_base + __popc(heads & ((1u << lane) - 1u));
        taps[pos] = make_uint2(id, __float_as_uint(w * wscale));
        sc.skey[pos] = id;
      }
      out_base += __popc(heads);
      __syncwarp();
    }
    nfinal = out_base;
  }
  const int pmin = (int)(sc.skey[0] & 0xfffu), pmax = (int)(sc.skey[nfinal - 1] & 0xfffu);
  // idx[q] = first tap whose pixel is >= pmin + q, q = 0 .. pmax - pmin + 1; `absolute`: the index is
  // addressed by the pixel itself, idx[pmin + q] (the rows: a reader then needs nothing of the ROI to
  // find its entry, so the load does not wait for the box)
  idx += absolute ? pmin : 0;
  for (int q = lane; q <= pmax - pmin + 1; q += 32) {
    const unsigned want = (unsigned)(pmin + q);
    int a = 0, b = nfinal;
    while (a < b) {
      const int mid = (a + b) >> 1;
      if ((sc.skey[mid] & 0xfffu) < want) a = mid + 1;
      else b = mid;
    }
    idx[q] = (unsigned short)a;
  }
  __syncwarp();
  *range = (unsigned)pmin | ((unsigned)pmax << 16);
  return nfinal;
}

template <bool DET>
__global__ void __launch_bounds__(kMaxGeomWarps * 32) pull_geom_kernel(const PullParams p) {
  extern __shared__ __align__(16) unsigned char s_geom[];
  ptx::chain_enter();
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int pos = blockIdx.x * p.geom_warps + wib;
  if (blockIdx.x == 0 && threadIdx.x == 0) *p.cursor = 0u;
  if (pos >= p.num_rois) return;
  const AxisScratch sc = carve_scratch(s_geom + (size_t)wib * axis_scratch_bytes(p.caps), p.caps);
  uint4 box = make_uint4(0xffffu, 0xffffu, 0u, 0u);
  if (pos < __ldg(p.img_start + p.num_images)) {       // rows with a batch index outside [0, N) scatter nothing
    const int r = __ldg(p.order + pos);
    const RoiGeom g = roi_decode(p.rois + (size_t)r * 5, p.scale, p.pooled_h, p.pooled_w, p.grid, p.aligned != 0);
    const float inv_count = __fdiv_rn(1.0f, g.count);
    unsigned yr, xr;
    const int nyt = pull_axis<DET>(g.start_h, g.bin_h, p.pooled_h, g.grid_h, p.height, 1.0f, p.caps, lane, sc,
                                   p.ytap + (size_t)pos * p.capy, p.yidx + (size_t)pos * (p.height + 2), &yr, true);
    __syncwarp();
    const int nxt = pull_axis<DET>(g.start_w, g.bin_w, p.pooled_w, g.grid_w, p.width, inv_count, p.caps, lane, sc,
                                   p.xtap + (size_t)pos * p.capx, p.xidx + (size_t)pos * (p.width + 2), &xr, false);
    if (nyt > 0 && nxt > 0) {
      box = make_uint4(yr, xr, (unsigned)g.grid_h | ((unsigned)g.grid_w << 16), __float_as_uint(g.count));
      // every (row tap, column tap) pair is one entry of exactly one pixel
      if (lane == 0) atomicAdd(p.need, (unsigned long long)nyt * (unsigned long long)nxt);
    }
  }
  if (lane == 0) p.gbox[pos] = box;
}

// ---- row mask: one warp per (mask word = 32 ROIs, chunk of 32 rows), lane = ROI -----------------
// (a warp per (row, word) spent its time on the binary search and one dependent load: C4 40 us)
__global__ void __launch_bounds__(256) pull_rowmask_kernel(const PullParams p) {
  ptx::chain_enter();
  const int lane = threadIdx.x & 31;
  const int chunks = (p.height + 31) >> 5;
  const long long wid = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (wid >= (long long)p.tw * chunks) return;
  const int wg = (int)(wid / chunks), y0 = (int)(wid - (long long)wg * chunks) * 32;
  // image of this word: the last n with mask_word0(n) <= wg
  int lo = 0, hi = p.num_images;        // answer in [lo, hi]; hi == num_images: behind the last image
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (mask_word0(p.img_start, mid) <= wg) lo = mid;
    else hi = mid - 1;
  }
  const int n = lo;
  int ymin = 0x7fffffff, ymax = -1;     // no row
  if (n < p.num_images) {
    const int pos = ((wg - n) << 5) + lane;
    if (pos >= __ldg(p.img_start + n) && pos < __ldg(p.img_start + n + 1)) {
      const unsigned yr = __ldg(&p.gbox[pos].x);
      ymin = (int)(yr & 0xffffu);       // a ROI without taps has ymin = 0xffff > every row
      ymax = (int)(yr >> 16);
    }
  }
  // lane l gathers the word of row y0 + l: the 32 stores of the chunk are one instruction
  unsigned mine = 0;
#pragma unroll
  for (int k = 0; k < 32; ++k) {
    const int y = y0 + k;
    const unsigned m = __ballot_sync(0xffffffffu, ymin <= y && y <= ymax);
    if (lane == k) mine = m;
  }
  if (y0 + lane < p.height) p.rowmask[(size_t)(y0 + lane) * p.tw + wg] = mine;
}

// ---- per-pixel lists: one thread per pixel, a warp = 32 consecutive pixels of one row ----------
template <bool WIDE>
struct EntryOf {
  typedef uint2 type;
};
template <>
struct EntryOf<true> {
  typedef uint4 type;
};

// walk<FILL = false> counts the entries of the lane's pixel, walk<true> writes them: both visit
// the ROIs of the row in ascending order, so a list is written in its final order by one thread.
// WIDE (deterministic mode on adaptive grids): the entry carries the ROI's sample count.
template <bool FILL, bool DET, bool WIDE>
__device__ __forceinline__ unsigned pull_walk(const PullParams& p, int n, int y, int x0, int lane,
                                              typename EntryOf<WIDE>::type* out) {
  const int x = x0 + lane;
  const bool in = x < p.width;
  const int w_lo = mask_word0(p.img_start, n), w_hi = mask_word0(p.img_start, n + 1);
  const int B = p.pooled_h * p.pooled_w;
  unsigned total = 0;
  for (int wbase = w_lo; wbase < w_hi; wbase += 32) {
    // 32 mask words at a time, one per lane; then the words with a bit set, one after the other
    const unsigned my_word = wbase + lane < w_hi ? __ldg(p.rowmask + (size_t)y * p.tw + wbase + lane) : 0u;
    unsigned words = __ballot_sync(0xffffffffu, my_word != 0u);
    while (words) {
      const int wj = __ffs(words) - 1;
      words &= words - 1;
      const int wg = wbase + wj;
      const unsigned m = __shfl_sync(0xffffffffu, my_word, wj);
      // the <= 32 ROIs of this word, one per lane: does it reach this row segment, and which of its
      // row taps sit on row y?  (independent loads across the lanes instead of a serial chain)
      unsigned my_r = 0, my_x = 0;               // ra | rb << 16, xmin | xmax << 16
      bool cand = (m >> lane) & 1u;
      if (cand) {
        const int pos = ((wg - n) << 5) + lane;
        // (the row index is addressed by y itself: its loads do not wait for the box)
        const unsigned xr = __ldg(&p.gbox[pos].y);
        const unsigned short* yi = p.yidx + (size_t)pos * (p.height + 2) + y;
        const unsigned ra = __ldg(yi), rb = __ldg(yi + 1);
        const int xmin = (int)(xr & 0xffffu), xmax = (int)(xr >> 16);
        cand = !(xmax < x0 || xmin > x0 + 31) && ra != rb;
        my_r = ra | (rb << 16);
        my_x = xr;
      }
      unsigned live = __ballot_sync(0xffffffffu, cand);
      // The live ROIs one after the other, software-pipelined: the index loads of the next one (its column
      // taps for this pixel, its row in the ROI table) are in flight while the current one is worked on.
      int n_pos = 0, n_ra = 0, n_rb = 0, n_ca = 0, n_cb = 0;
      unsigned n_row = 0;
      auto fetch = [&]() {                       // pops the lowest live ROI into n_*
        const int b = __ffs(live) - 1;
        live &= live - 1;
        n_pos = ((wg - n) << 5) + b;
        const unsigned rr = __shfl_sync(0xffffffffu, my_r, b), xx = __shfl_sync(0xffffffffu, my_x, b);
        n_ra = (int)(rr & 0xffffu);
        n_rb = (int)(rr >> 16);
        const int xmin = (int)(xx & 0xffffu), xmax = (int)(xx >> 16);
        n_ca = n_cb = 0;
        if (in && x >= xmin && x <= xmax) {
          const unsigned short* xi = p.xidx + (size_t)n_pos * (p.width + 2) + (x - xmin);
          n_ca = __ldg(xi);
          n_cb = __ldg(xi + 1);
        }
        if (FILL) n_row = (unsigned)__ldg(p.order + n_pos);
      };
      bool have = live != 0u;
      if (have) fetch();
      while (have) {
        const int pos = n_pos, ra = n_ra, rb = n_rb, ca = n_ca, cb = n_cb;
        const unsigned base = n_row * (unsigned)B;
        have = live != 0u;
        if (have) fetch();
        if (!FILL) {
          total += (unsigned)((rb - ra) * (cb - ca));
          continue;
        }
        if (ca == cb) continue;
        const uint2* yt = p.ytap + (size_t)pos * p.capy;
        const uint2* xt = p.xtap + (size_t)pos * p.capx;
        const unsigned cwb = (unsigned)(p.cw * 4), pxbits = (unsigned)(x & 7);     // cwb is a multiple of 16
        unsigned cnt_bits = 0, inv_bits = 0;
        if (WIDE) {
          const uint2 gc = __ldg(reinterpret_cast<const uint2*>(p.gbox + pos) + 1);   // grid sizes, count
          const unsigned cnt_i = (gc.x & 0xffffu) * (gc.x >> 16);
          cnt_bits = gc.y;
          if ((cnt_i & (cnt_i - 1u)) == 0u) inv_bits = __float_as_uint(__fdiv_rn(1.0f, __uint_as_float(gc.y)));
        }
        auto emit = [&](unsigned id, float w) {
          if (WIDE) *reinterpret_cast<uint4*>(out) = make_uint4(id, __float_as_uint(w), cnt_bits, inv_bits);
          else *reinterpret_cast<uint2*>(out) = make_uint2(id, __float_as_uint(w));
          ++out;
        };
        if (!DET) {
          for (int a = ra; a < rb; ++a) {
            const uint2 ya = __ldg(yt + a);
            const unsigned ph = ya.x >> 23;
            const float wy = __uint_as_float(ya.y);
            for (int c = ca; c < cb; ++c) {
              const uint2 xc = __ldg(xt + c);
              emit((base + ph * (unsigned)p.pooled_w + (xc.x >> 23)) * cwb | pxbits, wy * __uint_as_float(xc.y));
            }
          }
        } else {
          // the reference's order inside one ROI: bin row, bin column, sample row, sample column,
          // then the four taps (low/low, low/high, high/low, high/high); ROIAlign.cpp:126-157
          int a0 = ra;
          while (a0 < rb) {
            const unsigned ph = __ldg(&yt[a0].x) >> 23;
            int a1 = a0 + 1;
            while (a1 < rb && (__ldg(&yt[a1].x) >> 23) == ph) ++a1;
            int c0 = ca;
            while (c0 < cb) {
              const unsigned pw = __ldg(&xt[c0].x) >> 23;
              int c1 = c0 + 1;
              while (c1 < cb && (__ldg(&xt[c1].x) >> 23) == pw) ++c1;
              const unsigned id = (base + ph * (unsigned)p.pooled_w + pw) * cwb | pxbits;
              int a = a0;
              while (a < a1) {                                   // taps of one sample row: 1 or 2
                const unsigned sa = __ldg(&yt[a].x) >> 13;
                const int a2 = (a + 1 < a1 && (__ldg(&yt[a + 1].x) >> 13) == sa) ? a + 2 : a + 1;
                int c = c0;
                while (c < c1) {                                 // taps of one sample column
                  const unsigned sc = __ldg(&xt[c].x) >> 13;
                  const int c2 = (c + 1 < c1 && (__ldg(&xt[c + 1].x) >> 13) == sc) ? c + 2 : c + 1;
                  for (int aa = a; aa < a2; ++aa) {
                    const float wy = __uint_as_float(__ldg(&yt[aa].y));
                    for (int cc = c; cc < c2; ++cc) emit(id, __fmul_rn(wy, __uint_as_float(__ldg(&xt[cc].y))));
                  }
                  c = c2;
                }
                a = a2;
              }
              c0 = c1;
            }
            a0 = a1;
          }
        }
      }
    }
  }
  return total;
}

// Count, reserve, fill in one kernel: the warp counts the entries of its 32 pixels, takes one run
// of the entry array for all of them (one atomic: WHERE a list lives varies from run to run, what
// it holds does not), and writes them on a second walk over tables that are in the cache by then.
template <bool DET, bool WIDE>
__device__ __forceinline__ void pull_list_block(const PullParams& p, unsigned block) {
  typedef typename EntryOf<WIDE>::type Entry;
  const int lane = threadIdx.x & 31;
  const int segs = (p.width + 31) >> 5;
  const long long wid = (long long)block * 8 + (threadIdx.x >> 5);
  if (wid >= (long long)p.num_images * p.height * segs) return;
  const int seg = (int)(wid % segs);
  const int row = (int)(wid / segs);               // n * H + y
  const int n = row / p.height, y = row - n * p.height;
  const int x0 = seg * 32;
  const unsigned total = pull_walk<false, DET, WIDE>(p, n, y, x0, lane, nullptr);
  unsigned incl = total;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const unsigned v = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += v;
  }
  const unsigned warp_total = __shfl_sync(0xffffffffu, incl, 31);
  unsigned base = 0;
  if (lane == 0 && warp_total) base = atomicAdd(p.cursor, warp_total);
  base = __shfl_sync(0xffffffffu, base, 0);
  // (the host sized the array from the geometry kernel's count, or from the bound R * capy * capx)
  const bool fits = (unsigned long long)base + warp_total <= p.entry_cap;
  const unsigned begin = base + incl - total;
  if (x0 + lane < p.width) p.ptr[(long long)row * p.width + x0 + lane] = make_uint2(begin, fits ? total : 0u);
  if (!warp_total || !fits) return;
  Entry* mine = static_cast<Entry*>(p.entries) + begin;
  pull_walk<true, DET, WIDE>(p, n, y, x0, lane, mine);
  if (total) mine[total - 1].x |= 8u;       // the last entry of the pixel (written by this thread)
}

// ---- grouped lists (merged taps only): one entry per (ROI, bin, row) and GROUP of 8 pixels -------
// A bin of a box-head ROI covers ~3.4 consecutive pixels of a row, and the per-pixel kernel loads
// the bin's gradient line once for each of them (its LSU data pipe is 88 % busy with those loads).
// Here the 8 pixels a warp of the main kernel owns share ONE entry per bin: the line's offset and
// the 8 weights (0 where the bin does not reach the pixel).  The line is loaded once and multiplied
// into 8 accumulator sets held in registers.  A pixel meets its bins in the same order as in the
// per-pixel lists (ROI ascending, row tap, bin column), so the sums are the same fused
// multiply-adds in the same order -- bit-identical to the per-pixel merged kernel.
struct GroupEntry {
  unsigned hdr;        // byte offset of the bin's channel row in dyt (a multiple of 16) | pixel-pair mask:
                       // bit q set = pixel 2q or 2q + 1 has a non-zero weight
  unsigned pad[3];
  float w[8];
};
static_assert(sizeof(GroupEntry) == 48, "GroupEntry is three 16-byte words");

// Same walk as pull_walk, lane = pixel; the 8 lanes of a group execute in lock step (what they
// branch on is the same for the whole group) and write one weight each.
template <bool FILL>
__device__ __forceinline__ unsigned pull_walk8(const PullParams& p, int n, int y, int x0, int lane, GroupEntry* out) {
  const int x = x0 + lane;
  const bool in = x < p.width;
  const int j = lane & 7, gshift = lane & ~7;
  const int xg0 = x0 + gshift, xg1 = xg0 + 7;
  const unsigned gmask = 0xffu << gshift;
  const int w_lo = mask_word0(p.img_start, n), w_hi = mask_word0(p.img_start, n + 1);
  const int B = p.pooled_h * p.pooled_w;
  unsigned total = 0;
  for (int wbase = w_lo; wbase < w_hi; wbase += 32) {
    const unsigned my_word = wbase + lane < w_hi ? __ldg(p.rowmask + (size_t)y * p.tw + wbase + lane) : 0u;
    unsigned words = __ballot_sync(0xffffffffu, my_word != 0u);
    while (words) {
      const int wj = __ffs(words) - 1;
      words &= words - 1;
      const int wg = wbase + wj;
      const unsigned m = __shfl_sync(0xffffffffu, my_word, wj);
      unsigned my_r = 0, my_x = 0;
      bool cand = (m >> lane) & 1u;
      if (cand) {
        const int pos = ((wg - n) << 5) + lane;
        // (the row index is addressed by y itself: its loads do not wait for the box)
        const unsigned xr = __ldg(&p.gbox[pos].y);
        const unsigned short* yi = p.yidx + (size_t)pos * (p.height + 2) + y;
        const unsigned ra = __ldg(yi), rb = __ldg(yi + 1);
        const int xmin = (int)(xr & 0xffffu), xmax = (int)(xr >> 16);
        cand = !(xmax < x0 || xmin > x0 + 31) && ra != rb;
        my_r = ra | (rb << 16);
        my_x = xr;
      }
      unsigned live = __ballot_sync(0xffffffffu, cand);
      while (live) {
        const int b = __ffs(live) - 1;
        live &= live - 1;
        const int pos = ((wg - n) << 5) + b;
        const unsigned rr = __shfl_sync(0xffffffffu, my_r, b), xx = __shfl_sync(0xffffffffu, my_x, b);
        const int ra = (int)(rr & 0xffffu), rb = (int)(rr >> 16);
        const int xmin = (int)(xx & 0xffffu), xmax = (int)(xx >> 16);
        const unsigned short* xi = p.xidx + (size_t)pos * (p.width + 2);
        const uint2* xt = p.xtap + (size_t)pos * p.capx;
        // the bin columns that reach the group: the taps of its pixels are one run, sorted by
        // (pixel, bin column), and every bin column between the first and the last one is present
        int npw = 0, pwf = 0;
        const int glo = max(xg0, xmin), ghi = min(xg1, xmax);
        if (glo <= ghi) {
          const int ga = __ldg(xi + (glo - xmin)), gb = __ldg(xi + (ghi + 1 - xmin));
          if (gb > ga) {
            pwf = (int)(__ldg(&xt[ga].x) >> 23);
            npw = (int)(__ldg(&xt[gb - 1].x) >> 23) - pwf + 1;
          }
        }
        if (!FILL) {
          total += (unsigned)((rb - ra) * npw);
          continue;
        }
        if (npw == 0) continue;
        int ca = 0, cb = 0;
        if (in && x >= xmin && x <= xmax) {
          ca = __ldg(xi + (x - xmin));
          cb = __ldg(xi + (x - xmin) + 1);
        }
        const unsigned base = (unsigned)__ldg(p.order + pos) * (unsigned)B;
        const uint2* yt = p.ytap + (size_t)pos * p.capy;
        const unsigned cwb = (unsigned)(p.cw * 4);
        for (int a = ra; a < rb; ++a) {
          const uint2 ya = __ldg(yt + a);
          const unsigned ph = ya.x >> 23;
          const float wy = __uint_as_float(ya.y);
          int c = ca;
          for (int q = 0; q < npw; ++q) {
            float w = 0.0f;
            if (c < cb) {
              const uint2 xc = __ldg(xt + c);
              if ((int)(xc.x >> 23) == pwf + q) {
                w = wy * __uint_as_float(xc.y);
                ++c;
              }
            }
            const unsigned nz = (__ballot_sync(gmask, w != 0.0f) >> gshift) & 0xffu;
            out->w[j] = w;
            if (j == 0) {
              const unsigned pairs = ((nz & 0x03u) ? 1u : 0u) | ((nz & 0x0cu) ? 2u : 0u) | ((nz & 0x30u) ? 4u : 0u) |
                                     ((nz & 0xc0u) ? 8u : 0u);
              out->hdr = (base + ph * (unsigned)p.pooled_w + (unsigned)(pwf + q)) * cwb | pairs;
            }
            ++out;
          }
        }
      }
    }
  }
  return total;
}

// p.ptr is indexed by group here: ((n * H + y) * ceil(W / 8) + x / 8) -> first entry, entries
__device__ __forceinline__ void pull_list_block8(const PullParams& p, unsigned block) {
  const int lane = threadIdx.x & 31;
  const int segs = (p.width + 31) >> 5, segs8 = (p.width + 7) >> 3;
  const long long wid = (long long)block * 8 + (threadIdx.x >> 5);
  if (wid >= (long long)p.num_images * p.height * segs) return;
  const int seg = (int)(wid % segs);
  const int row = (int)(wid / segs);               // n * H + y
  const int n = row / p.height, y = row - n * p.height;
  const int x0 = seg * 32;
  const unsigned total = pull_walk8<false>(p, n, y, x0, lane, nullptr);     // the same on the 8 lanes of a group
  const unsigned t0 = __shfl_sync(0xffffffffu, total, 0), t1 = __shfl_sync(0xffffffffu, total, 8),
                 t2 = __shfl_sync(0xffffffffu, total, 16), t3 = __shfl_sync(0xffffffffu, total, 24);
  const unsigned warp_total = t0 + t1 + t2 + t3;
  const int g = lane >> 3;
  const unsigned before = (g > 0 ? t0 : 0u) + (g > 1 ? t1 : 0u) + (g > 2 ? t2 : 0u);
  unsigned base = 0;
  if (lane == 0 && warp_total) base = atomicAdd(p.cursor, warp_total);
  base = __shfl_sync(0xffffffffu, base, 0);
  const bool fits = (unsigned long long)base + warp_total <= p.entry_cap;
  const unsigned begin = base + before;
  if ((lane & 7) == 0 && x0 + lane < p.width)
    p.ptr[(long long)row * segs8 + ((x0 + lane) >> 3)] = make_uint2(begin, fits ? total : 0u);
  if (!warp_total || !fits) return;
  pull_walk8<true>(p, n, y, x0, lane, static_cast<GroupEntry*>(p.entries) + begin);
}

// ---- channel-last copy of dy: [R][C][B] -> [R * B][cw] -------------------------------------
// One block per (ROI, 128 channels): the block's source is one contiguous run of 128 * B floats.
// It goes to shared memory as [channel][pitch] (pitch odd: the transposed read, lane = channel,
// is conflict-free) and leaves as B rows of 128 consecutive channels.
// kTrChannels = 128 for box-head bins (49 per channel), 32 for the mask head's 196: more blocks per SM
template <int kTrChannels>
__device__ __forceinline__ void pull_transpose_block(const PullParams& p, unsigned block, float* s_t) {
  const int B = p.pooled_h * p.pooled_w, pitch = B | 1;
  const int groups = (p.cw + kTrChannels - 1) / kTrChannels;
  const int r = block / groups, c0 = (block - r * groups) * kTrChannels;
  const int nc = min(kTrChannels, p.c_work - c0);          // real channels of this block
  const int ncw = min(kTrChannels, p.cw - c0);             // incl. the zero pad channels
  if (p.dy_layout != 0) {
    // dy is channel-last already ([R, PH, PW, C], any element type): widen and copy, coalesced both ways
    // (a half-precision dy whose rows are whole, cw == C, goes through pull_widen_block instead)
    const int c = threadIdx.x & (kTrChannels - 1);
    if (c < ncw)
      for (int bb = threadIdx.x / kTrChannels; bb < B; bb += 256 / kTrChannels)
        p.dyt[((size_t)r * B + bb) * p.cw + c0 + c] =
            c < nc ? load_top(p.dy, ((size_t)r * B + bb) * p.channels + c0 + c, p.dy_dtype) : 0.0f;
    return;
  }
  const size_t src0 = ((size_t)r * p.channels + c0) * B;    // element index of the block's source run
  const float* src = static_cast<const float*>(p.dy) + src0;
  const int total = nc * B;
  if (p.dy_dtype != kTopF32) {
    // 2-byte elements, widened on the way into shared memory (the type test stays outside the loop)
    const unsigned short* src16 = static_cast<const unsigned short*>(p.dy) + src0;
    const int dc = 256 / B, db = 256 - dc * B;
    int c = threadIdx.x / B, bb = threadIdx.x - c * B;
    if (p.dy_dtype == kTopF16) {
      for (int i = threadIdx.x; i < total; i += 256) {
        s_t[c * pitch + bb] = __half2float(__ushort_as_half(__ldg(src16 + i)));
        c += dc;
        bb += db;
        if (bb >= B) {
          bb -= B;
          ++c;
        }
      }
    } else {
      for (int i = threadIdx.x; i < total; i += 256) {
        s_t[c * pitch + bb] = __uint_as_float((unsigned)__ldg(src16 + i) << 16);      // bf16 -> fp32: exact
        c += dc;
        bb += db;
        if (bb >= B) {
          bb -= B;
          ++c;
        }
      }
    }
  } else if (pitch == B && (reinterpret_cast<uintptr_t>(src) & 15u) == 0) {
    // odd B: the shared-memory image is the source run itself, copied 16 bytes at a time
    const int quads = total >> 2;
    float4* dst4 = reinterpret_cast<float4*>(s_t);
    for (int q = threadIdx.x; q < quads; q += 256) dst4[q] = __ldg(reinterpret_cast<const float4*>(src) + q);
    for (int i = (quads << 2) + threadIdx.x; i < total; i += 256) s_t[i] = __ldg(src + i);
  } else if ((B & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 15u) == 0) {
    // B a multiple of 4 (the mask head's 14 x 14): 16-byte loads that never straddle two channels;
    // (channel, bin) of a thread's quad advance by (1024 / B, 1024 % B) per step
    const int quads = total >> 2;
    const int dc = 1024 / B, db = 1024 - dc * B;
    int c = (4 * threadIdx.x) / B, bb = 4 * threadIdx.x - c * B;
    for (int q = threadIdx.x; q < quads; q += 256) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(src) + q);
      float* d = s_t + c * pitch + bb;
      d[0] = v.x;
      d[1] = v.y;
      d[2] = v.z;
      d[3] = v.w;
      c += dc;
      bb += db;
      if (bb >= B) {
        bb -= B;
        ++c;
      }
    }
  } else {
    // (channel, bin) of element i advance by (256 / B, 256 % B) per step: no division in the loop
    const int dc = 256 / B, db = 256 - dc * B;
    int c = threadIdx.x / B, bb = threadIdx.x - c * B;
    for (int i = threadIdx.x; i < total; i += 256) {
      s_t[c * pitch + bb] = __ldg(src + i);
      c += dc;
      bb += db;
      if (bb >= B) {
        bb -= B;
        ++c;
      }
    }
  }
  __syncthreads();
  const int c = threadIdx.x & (kTrChannels - 1);
  if (c < ncw) {
    // pointers stepped per row of the copy: no 64-bit multiply in the loop
    constexpr int kStep = 256 / kTrChannels;
    int bb = threadIdx.x / kTrChannels;
    float* d = p.dyt + ((size_t)r * B + bb) * p.cw + c0 + c;
    const size_t dstep = (size_t)kStep * p.cw;
    if (c < nc) {
      const float* m = s_t + c * pitch + bb;
#pragma unroll 4
      for (; bb < B; bb += kStep, d += dstep, m += kStep) *d = *m;
    } else {
      for (; bb < B; bb += kStep, d += dstep) *d = 0.0f;
    }
  }
}

// A channel-last half-precision dy whose rows are whole (cw == C == c_work): the "copy" is a flat
// widening of R * B * C elements, 8 per thread (one 16-byte load, two 16-byte stores).
__device__ __forceinline__ void pull_widen_block(const PullParams& p, unsigned block) {
  const size_t total = (size_t)p.num_rois * p.pooled_h * p.pooled_w * p.cw;
  const size_t i = ((size_t)block * 256 + threadIdx.x) * 8;
  if (i >= total) return;
  const unsigned short* src = static_cast<const unsigned short*>(p.dy) + i;
  float v[8];
  if (i + 8 <= total && (reinterpret_cast<uintptr_t>(src) & 15u) == 0) {
    const uint4 raw = __ldg(reinterpret_cast<const uint4*>(src));
    const unsigned w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (p.dy_dtype == kTopF16) {
        v[2 * k] = __half2float(__ushort_as_half((unsigned short)(w[k] & 0xffffu)));
        v[2 * k + 1] = __half2float(__ushort_as_half((unsigned short)(w[k] >> 16)));
      } else {
        v[2 * k] = __uint_as_float(w[k] << 16);
        v[2 * k + 1] = __uint_as_float(w[k] & 0xffff0000u);
      }
    }
    float4* dst = reinterpret_cast<float4*>(p.dyt + i);      // cw % 4 == 0 and i % 8 == 0: 16-byte aligned
    dst[0] = make_float4(v[0], v[1], v[2], v[3]);
    dst[1] = make_float4(v[4], v[5], v[6], v[7]);
  } else {
    for (size_t k = i; k < total && k < i + 8; ++k) p.dyt[k] = load_top(p.dy, k, p.dy_dtype);
  }
}

// The lists and the channel-last copy in ONE grid: building lists is latency-bound index work,
// the copy is pure bandwidth, and neither depends on the other -- side by side the copy is free.
// Block roles are interleaved (`lblocks` list blocks, `tblocks` copy blocks) so both progress.
// resident CTAs per SM the kernel is compiled for: the list walk waits on dependent table loads,
// a fifth CTA (<= 48 registers, four spilled words) hides more of them (measured 4 / 5 / 6 CTAs on the
// whole backward: C3 2.07 / 1.96 / 1.96 ms, C4 3.30 / 3.18 / 3.19 ms, C2 143 / 143 / 147 us)
#ifndef MOBULA_TABLES_MINBLOCKS
#define MOBULA_TABLES_MINBLOCKS 5
#endif
template <bool DET, bool WIDE, int kTrChannels, bool GRP>
__global__ void __launch_bounds__(256, MOBULA_TABLES_MINBLOCKS) pull_tables_kernel(const PullParams p, const unsigned lblocks,
                                                          const unsigned tblocks) {
  extern __shared__ float s_t[];      // copy blocks: [kTrChannels][B | 1]
  ptx::chain_enter();
  const unsigned i = blockIdx.x;
  const unsigned few = min(lblocks, tblocks), many = max(lblocks, tblocks);
  bool is_few = false;
  unsigned idx = i;
  if (few > 0) {
    // one block of the smaller set every q blocks; q = many / few + 1 spreads it over the whole grid, a
    // smaller q (p.tables_q) runs it out earlier: the rest of the grid is blocks of the larger set only
    unsigned q = many / few + 1;
    if (lblocks <= tblocks && p.tables_q > 0 && (unsigned)p.tables_q < q) q = (unsigned)p.tables_q;
    const unsigned few_before = min(few, (i + q - 1) / q);
    is_few = i % q == 0 && i / q < few;
    idx = is_few ? i / q : i - few_before;
  }
  const bool list_role = (lblocks <= tblocks) == is_few;      // the smaller set is the list set iff lblocks <= tblocks
  if (few == 0 ? lblocks > 0 : list_role) {
    if (GRP) pull_list_block8(p, idx);
    else pull_list_block<DET, WIDE>(p, idx);
  }
  else if (p.widen_flat) pull_widen_block(p, idx);
  else pull_transpose_block<kTrChannels>(p, idx, s_t);
}

// ---- the main kernel ------------------------------------------------------------------------
// CTA = kPullRows x kPullCols pixels (8 x 8) x (32 * CPL) channels; warp w: row w / kPullWarpsPerRow,
// 8 consecutive pixels of it.
// The entries of a warp's eight pixels are one contiguous run of the list: the warp brings them
// to shared memory 64 at a time (coalesced) and reads them back as broadcasts, so the only
// global load on the dependent path of a term is its gradient line.
#ifndef MOBULA_PULL_THREADS
#define MOBULA_PULL_THREADS 256
#endif
constexpr int kPullThreads = MOBULA_PULL_THREADS;
// resident CTAs per SM the main kernel is compiled for: the merged-tap kernel is bound by the latency
// of its gradient loads, a fourth CTA (<= 64 registers, a few spilled words) is worth 5-14 %; the
// canonical-order kernel keeps more state per term and loses with anything tighter than three
// (measured 3 / 4 / 5 CTAs: C2 150 / 143 / 156 us, C3 2.40 / 2.07 / 2.59 ms, C4 3.51 / 3.30 / 3.56 ms)
constexpr int pull_min_blocks(bool det) { return (det ? 3 : 4) * 256 / kPullThreads > 0 ? (det ? 3 : 4) * 256 / kPullThreads : 1; }
// Tile shape: the pixels of a CTA re-read the gradient lines of the bins they share out of L1, and a
// square tile meets fewer bins per ROI than a strip (bins of 6 x 6 pixels: 5.4 against 8.4 lines per
// ROI).  Measured 2 x 32 / 4 x 16 / 8 x 8 on the whole backward: C2 143 / 140 / 137 us,
// C3 1.96 / 1.90 / 1.79 ms, C4 3.18 / 3.08 / 3.01 ms.
#ifndef MOBULA_PULL_ROWS
#define MOBULA_PULL_ROWS 8
#endif
constexpr int kPullRows = MOBULA_PULL_ROWS;                     // rows of a CTA's pixel tile: 2, 4 or 8
constexpr int kPullWarpsPerRow = (kPullThreads / 32) / kPullRows;
constexpr int kPullCols = 8 * kPullWarpsPerRow;                 // columns of the tile (a warp = 8 pixels of one row)
constexpr int kEntChunk = 64;

template <int CPL>
__device__ __forceinline__ void load_vec(const float* ptr, float (&g)[CPL]) {
  if (CPL == 4) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(ptr));
    g[0] = t.x;
    g[1 % CPL] = t.y;
    g[2 % CPL] = t.z;
    g[3 % CPL] = t.w;
  } else if (CPL == 2) {
    const float2 t = __ldg(reinterpret_cast<const float2*>(ptr));
    g[0] = t.x;
    g[1 % CPL] = t.y;
  } else {
    g[0] = __ldg(ptr);
  }
}

template <int CPL, bool DET>
__device__ __forceinline__ void add_term(float (&acc)[CPL], const float (&g)[CPL], float w, float count, float inv,
                                         int pow2) {
#pragma unroll
  for (int k = 0; k < CPL; ++k) {
    if (DET) {
      const float t = __fmul_rn(g[k], w);                  // dtop * w, then / count (ROIAlign.cpp:143-146)
      acc[k] = __fadd_rn(acc[k], pow2 ? __fmul_rn(t, inv) : __fdiv_rn(t, count));
    } else {
      acc[k] = fmaf(g[k], w, acc[k]);
    }
  }
}

__device__ __forceinline__ uint2 zero_entry(uint2*) { return make_uint2(0u, 0u); }
__device__ __forceinline__ uint4 zero_entry(uint4*) { return make_uint4(0u, 0u, 0u, 0u); }
// sample count of the term's ROI: a kernel constant for fixed sampling ratios, in the entry otherwise
__device__ __forceinline__ void count_of(const uint2&, float& count, float& inv, int& pow2) {
  (void)count;
  (void)inv;
  (void)pow2;
}
__device__ __forceinline__ void count_of(const uint4& en, float& count, float& inv, int& pow2) {
  count = __uint_as_float(en.z);
  inv = __uint_as_float(en.w);
  pow2 = en.w != 0u;
}

template <int CPL, bool DET, bool ACC, int U, bool WIDE>
__global__ void __launch_bounds__(kPullThreads, pull_min_blocks(DET)) pull_bwd_kernel(const PullParams p, const float count0,
                                                                const float inv0, const int pow20) {
  typedef typename EntryOf<WIDE>::type Entry;
  constexpr int CG = 32 * CPL;                        // channels per CTA
  // per warp: its 8 pixels x CG channels, [pixel][channel slot], and a ring of list entries
  extern __shared__ __align__(16) unsigned char s_pull[];
  float (*s_tile)[8 * CG] = reinterpret_cast<float (*)[8 * CG]>(s_pull);
  Entry (*s_ent)[2 * kEntChunk] =
      reinterpret_cast<Entry (*)[2 * kEntChunk]>(s_pull + sizeof(float) * (kPullThreads / 32) * 8 * CG);
  ptx::chain_enter();
  const Entry* entries = static_cast<const Entry*>(p.entries);
  float count = count0, inv = inv0;
  int pow2 = pow20;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int segs = (p.width + kPullCols - 1) / kPullCols, bands = (p.height + kPullRows - 1) / kPullRows;
  const int seg = blockIdx.x % segs;
  const int band = (blockIdx.x / segs) % bands;
  const int n = blockIdx.x / (segs * bands);
  const int c0 = blockIdx.y * CG;
  const int y = band * kPullRows + warp / kPullWarpsPerRow;
  const int xw = seg * kPullCols + (warp % kPullWarpsPerRow) * 8;
  if (y >= p.height || xw >= p.width) return;         // no block-wide barrier below: warps are independent
  const bool lane_on = c0 + lane * CPL < p.cw;
  const char* dyt = reinterpret_cast<const char*>(p.dyt + (lane_on ? c0 + lane * CPL : 0));
  float* tile = s_tile[warp];
  Entry* ring = s_ent[warp];

  const long long pix0 = ((long long)n * p.height + y) * p.width + xw;
  const int npx = min(8, p.width - xw);
  // The entries of the warp's pixels are ONE contiguous run [e0, end_all) of the entry array.  It
  // is walked in groups of U terms whatever the pixel boundaries (U gradient loads in flight even
  // where a pixel has only a few entries); an entry carries its pixel (low three bits) and a flag
  // on the last entry of the pixel, where the running sums go to the tile.
  uint2 my_ptr = make_uint2(0u, 0u);
  if (lane < npx) my_ptr = __ldg(p.ptr + pix0 + lane);
  const unsigned e0 = __shfl_sync(0xffffffffu, my_ptr.x, 0);
  const unsigned end_all = __shfl_sync(0xffffffffu, my_ptr.x + my_ptr.y, npx - 1);
  // tile[pixel][slot]: the lane's CPL channels as one vector; pixels 4 .. 7 swap the halves of the
  // row (slot ^ 16 channels) so that the read-out below is conflict-free.  Zero first: a pixel
  // without entries is never stored.
  float* tlane = tile + CPL * lane;
  constexpr int kSwap = CPL * (16 / CPL);           // 16 channels, in floats
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float* trow = tlane + i * CG;
    if (CPL == 4) *reinterpret_cast<float4*>(trow) = make_float4(0.f, 0.f, 0.f, 0.f);
    else if (CPL == 2) *reinterpret_cast<float2*>(trow) = make_float2(0.f, 0.f);
    else trow[0] = 0.f;
  }
  float acc[CPL];
#pragma unroll
  for (int k = 0; k < CPL; ++k) acc[k] = 0.0f;
  auto flush = [&](unsigned bits) {
    const int px = (int)(bits & 7u);
    float* trow = tile + px * CG + ((CPL * lane) ^ ((px & 4) ? kSwap : 0));
    if (CPL == 4) *reinterpret_cast<float4*>(trow) = make_float4(acc[0], acc[1 % CPL], acc[2 % CPL], acc[3 % CPL]);
    else if (CPL == 2) *reinterpret_cast<float2*>(trow) = make_float2(acc[0], acc[1 % CPL]);
    else trow[0] = acc[0];
#pragma unroll
    for (int k = 0; k < CPL; ++k) acc[k] = 0.0f;
  };
  // the ring holds two chunks of entries; the next chunk is fetched into registers while the
  // current one is consumed
  Entry nxt[kEntChunk / 32];
#pragma unroll
  for (int k = 0; k < kEntChunk / 32; ++k) {
    const unsigned i = e0 + k * 32 + lane;
    if (i < end_all) ring[k * 32 + lane] = __ldg(entries + i);
  }
  __syncwarp();
  int buf = 0;
  for (unsigned cbase = e0; cbase < end_all; cbase += kEntChunk, buf ^= 1) {
    const unsigned cn = min((unsigned)kEntChunk, end_all - cbase);
#pragma unroll
    for (int k = 0; k < kEntChunk / 32; ++k) {
      const unsigned i = cbase + kEntChunk + k * 32 + lane;
      nxt[k] = i < end_all ? __ldg(entries + i) : zero_entry(static_cast<Entry*>(nullptr));
    }
    const Entry* se = ring + buf * kEntChunk;
    unsigned t = 0;
    for (; t + U <= cn; t += U) {
      Entry en[U];
      float g[U][CPL];
#pragma unroll
      for (int u = 0; u < U; ++u) en[u] = se[t + u];
#pragma unroll
      for (int u = 0; u < U; ++u) load_vec<CPL>(reinterpret_cast<const float*>(dyt + (en[u].x & ~15u)), g[u]);
#pragma unroll
      for (int u = 0; u < U; ++u) {
        count_of(en[u], count, inv, pow2);
        add_term<CPL, DET>(acc, g[u], __uint_as_float(en[u].y), count, inv, pow2);
        if (en[u].x & 8u) flush(en[u].x);
      }
    }
    for (; t < cn; ++t) {
      const Entry en = se[t];
      float g[CPL];
      load_vec<CPL>(reinterpret_cast<const float*>(dyt + (en.x & ~15u)), g);
      count_of(en, count, inv, pow2);
      add_term<CPL, DET>(acc, g, __uint_as_float(en.y), count, inv, pow2);
      if (en.x & 8u) flush(en.x);
    }
    __syncwarp();
#pragma unroll
    for (int k = 0; k < kEntChunk / 32; ++k) ring[(buf ^ 1) * kEntChunk + k * 32 + lane] = nxt[k];
    __syncwarp();
  }
  __syncwarp();
  // read-out by the warp itself: every dX element of its 8 pixels is written once
  const size_t plane = (size_t)p.height * p.width;
  const int nch = min(CG, p.c_work - c0);
  float* base = p.dx + ((size_t)n * p.channels + c0) * plane + (size_t)y * p.width + xw;
  if (npx == 8 && (p.width & 3) == 0 && (reinterpret_cast<uintptr_t>(p.dx) & 15u) == 0) {
    // a step = 16 channels x 8 pixels: lane = (channel k, pixel quad h); four conflict-free reads
    // (banks k ^ 16 h), one 16-byte store per lane = sixteen full 32-byte sectors per step
    const int k = lane >> 1, h = lane & 1;
#pragma unroll 2
    for (int cb = 0; cb < CG; cb += 16) {
      const int c = cb + k;
      if (c >= nch) continue;
      const float* t = tile + (4 * h) * CG + (c ^ (16 * h));
      float4 v = make_float4(t[0], t[CG], t[2 * CG], t[3 * CG]);
      float4* dst = reinterpret_cast<float4*>(base + (size_t)c * plane + 4 * h);
      if (ACC) {
        const float4 o = *dst;
        v.x = __fadd_rn(o.x, v.x);
        v.y = __fadd_rn(o.y, v.y);
        v.z = __fadd_rn(o.z, v.z);
        v.w = __fadd_rn(o.w, v.w);
      }
      *dst = v;
    }
  } else {
    // ragged group / unaligned rows: lane = (channel, pixel) pairs, four channels per step
    const int px = lane & 7, k = lane >> 3;
    if (px < npx)
      for (int c = k; c < nch; c += 4) {
        const float v = tile[px * CG + (c ^ ((px & 4) ? 16 : 0))];
        float* dst = base + (size_t)c * plane + px;
        *dst = ACC ? __fadd_rn(*dst, v) : v;
      }
  }
}

// ---- the main kernel on grouped lists ---------------------------------------------------------
// Same tiling (warp = 8 consecutive pixels x 32 * CPL channels).  An entry is one gradient line and
// 8 weights: the line is loaded once (16 bytes per lane) and multiplied into the 8 pixels' sums,
// held in registers as pairs of neighbouring pixels so that one FFMA2 serves two of them; a pair
// whose two weights are zero is skipped (warp-uniform test on the entry's mask).
constexpr int kG8Chunk = 64;                          // entries per half of a warp's ring (64 x 48 bytes)

// U = gradient lines per batch, MINB = resident CTAs per SM the kernel is compiled for, PIPE = the
// loads of the next batch are issued before the current batch is consumed
template <int CPL, bool ACC, int U, int MINB, bool PIPE>
__global__ void __launch_bounds__(kPullThreads, MINB) pull8_bwd_kernel(const PullParams p) {
  constexpr int CG = 32 * CPL;
  constexpr int kHalf = 3 * kG8Chunk;                 // uint4 words of one half of the ring
  constexpr int kWarpBytes = (2 * kHalf * 16 > 8 * CG * 4) ? 2 * kHalf * 16 : 8 * CG * 4;
  // per warp: the ring of entries; once the list is consumed, the same bytes hold the 8 x CG tile
  extern __shared__ __align__(16) unsigned char s_pull[];
  ptx::chain_enter();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int segs = (p.width + kPullCols - 1) / kPullCols, bands = (p.height + kPullRows - 1) / kPullRows;
  const int seg = blockIdx.x % segs;
  const int band = (blockIdx.x / segs) % bands;
  const int n = blockIdx.x / (segs * bands);
  const int c0 = blockIdx.y * CG;
  const int y = band * kPullRows + warp / kPullWarpsPerRow;
  const int xw = seg * kPullCols + (warp % kPullWarpsPerRow) * 8;
  if (y >= p.height || xw >= p.width) return;         // warps are independent: no block-wide barrier below
  const bool lane_on = c0 + lane * CPL < p.cw;
  const char* dyt = reinterpret_cast<const char*>(p.dyt + (lane_on ? c0 + lane * CPL : 0));
  uint4* ring = reinterpret_cast<uint4*>(s_pull + (size_t)warp * kWarpBytes);
  float* tile = reinterpret_cast<float*>(ring);
  const int npx = min(8, p.width - xw);
  const int segs8 = (p.width + 7) >> 3;
  const uint2 run = __ldg(p.ptr + ((long long)n * p.height + y) * segs8 + (xw >> 3));
  const unsigned e0 = run.x, end_all = run.x + run.y;
  const uint4* entries = static_cast<const uint4*>(p.entries);      // three words per entry

  float2 acc[4][CPL];                                 // [pixel pair][channel]: .x = pixel 2q, .y = pixel 2q + 1
#pragma unroll
  for (int q = 0; q < 4; ++q)
#pragma unroll
    for (int k = 0; k < CPL; ++k) acc[q][k] = make_float2(0.0f, 0.0f);

  // the ring is filled with 16-byte cp.async copies (no registers held across the loop)
  auto fill = [&](unsigned first, int half) {
    const unsigned long long w0 = 3ull * first, wend = 3ull * end_all;
#pragma unroll
    for (int k = 0; k < kHalf / 32; ++k) {
      const unsigned long long i = w0 + k * 32 + lane;
      if (i < wend) ptx::cp_async16(ring + half * kHalf + k * 32 + lane, entries + i);
    }
  };
  fill(e0, 0);
  ptx::cp_async_wait_all();
  __syncwarp();
  auto term = [&](const float (&g)[CPL], unsigned hdr, const uint4* e) {
    const float4 wa = *reinterpret_cast<const float4*>(e + 1), wb = *reinterpret_cast<const float4*>(e + 2);
    if (hdr & 1u) {
#pragma unroll
      for (int k = 0; k < CPL; ++k) acc[0][k] = __ffma2_rn(make_float2(g[k], g[k]), make_float2(wa.x, wa.y), acc[0][k]);
    }
    if (hdr & 2u) {
#pragma unroll
      for (int k = 0; k < CPL; ++k) acc[1][k] = __ffma2_rn(make_float2(g[k], g[k]), make_float2(wa.z, wa.w), acc[1][k]);
    }
    if (hdr & 4u) {
#pragma unroll
      for (int k = 0; k < CPL; ++k) acc[2][k] = __ffma2_rn(make_float2(g[k], g[k]), make_float2(wb.x, wb.y), acc[2][k]);
    }
    if (hdr & 8u) {
#pragma unroll
      for (int k = 0; k < CPL; ++k) acc[3][k] = __ffma2_rn(make_float2(g[k], g[k]), make_float2(wb.z, wb.w), acc[3][k]);
    }
  };
  int buf = 0;
  for (unsigned cbase = e0; cbase < end_all; cbase += kG8Chunk, buf ^= 1) {
    const unsigned cn = min((unsigned)kG8Chunk, end_all - cbase);
    fill(cbase + kG8Chunk, buf ^ 1);
    const uint4* se = ring + buf * kHalf;
    // a batch = U entries; entries behind the end of the chunk have header 0 (no pair, offset 0)
    auto issue = [&](unsigned t, unsigned (&hdr)[U], float (&g)[U][CPL]) {
#pragma unroll
      for (int u = 0; u < U; ++u) hdr[u] = t + u < cn ? se[3 * (t + u)].x : 0u;
#pragma unroll
      for (int u = 0; u < U; ++u) load_vec<CPL>(reinterpret_cast<const float*>(dyt + (hdr[u] & ~15u)), g[u]);
    };
    auto consume = [&](unsigned t, const unsigned (&hdr)[U], const float (&g)[U][CPL]) {
#pragma unroll
      for (int u = 0; u < U; ++u) term(g[u], hdr[u], se + 3 * (t + u));
    };
    if (PIPE) {
      unsigned ha[U], hb[U];
      float ga[U][CPL], gb[U][CPL];
      issue(0, ha, ga);
      for (unsigned t = 0; t < cn; t += 2 * U) {
        if (t + U < cn) issue(t + U, hb, gb);
        consume(t, ha, ga);
        if (t + U < cn) {
          if (t + 2 * U < cn) issue(t + 2 * U, ha, ga);
          consume(t + U, hb, gb);
        }
      }
    } else {
      for (unsigned t = 0; t < cn; t += U) {
        unsigned ha[U];
        float ga[U][CPL];
        issue(t, ha, ga);
        consume(t, ha, ga);
      }
    }
    ptx::cp_async_wait_all();
    __syncwarp();
  }
  __syncwarp();
  // tile[pixel][slot] as in pull_bwd_kernel: pixels 4 .. 7 swap the halves of the row
  constexpr int kSwap = CPL * (16 / CPL);
#pragma unroll
  for (int px = 0; px < 8; ++px) {
    float* trow = tile + px * CG + ((CPL * lane) ^ ((px & 4) ? kSwap : 0));
    float v[CPL];
#pragma unroll
    for (int k = 0; k < CPL; ++k) v[k] = (px & 1) ? acc[px >> 1][k].y : acc[px >> 1][k].x;
    if (CPL == 4) *reinterpret_cast<float4*>(trow) = make_float4(v[0], v[1 % CPL], v[2 % CPL], v[3 % CPL]);
    else if (CPL == 2) *reinterpret_cast<float2*>(trow) = make_float2(v[0], v[1 % CPL]);
    else trow[0] = v[0];
  }
  __syncwarp();
  const size_t plane = (size_t)p.height * p.width;
  const int nch = min(CG, p.c_work - c0);
  float* base = p.dx + ((size_t)n * p.channels + c0) * plane + (size_t)y * p.width + xw;
  if (npx == 8 && (p.width & 3) == 0 && (reinterpret_cast<uintptr_t>(p.dx) & 15u) == 0) {
    const int k = lane >> 1, h = lane & 1;
#pragma unroll 2
    for (int cb = 0; cb < CG; cb += 16) {
      const int c = cb + k;
      if (c >= nch) continue;
      const float* t = tile + (4 * h) * CG + (c ^ (16 * h));
      float4 v = make_float4(t[0], t[CG], t[2 * CG], t[3 * CG]);
      float4* dst = reinterpret_cast<float4*>(base + (size_t)c * plane + 4 * h);
      if (ACC) {
        const float4 o = *dst;
        v.x = __fadd_rn(o.x, v.x);
        v.y = __fadd_rn(o.y, v.y);
        v.z = __fadd_rn(o.z, v.z);
        v.w = __fadd_rn(o.w, v.w);
      }
      *dst = v;
    }
  } else {
    const int px = lane & 7, k = lane >> 3;
    if (px < npx)
      for (int c = k; c < nch; c += 4) {
        const float v = tile[px * CG + (c ^ ((px & 4) ? 16 : 0))];
        float* dst = base + (size_t)c * plane + px;
        *dst = ACC ? __fadd_rn(*dst, v) : v;
      }
  }
}

// grouped lists for the merged (non-deterministic) mode; MOBULA_PULL_GROUP=0 / 1 overrides the default
#ifndef MOBULA_PULL8_VARIANT_DEFAULT
#define MOBULA_PULL8_VARIANT_DEFAULT 0
#endif
#ifndef MOBULA_PULL_GROUP_DEFAULT
#define MOBULA_PULL_GROUP_DEFAULT 0
#endif
static bool pull_grouped_enabled() {      // read on every call (one getenv): the tests switch it inside one process
  const char* e = getenv("MOBULA_PULL_GROUP");
  return e ? atoi(e) != 0 : MOBULA_PULL_GROUP_DEFAULT != 0;
}

struct PullLayout {
  size_t img_start, order, gbox, yidx, xidx, ytap, xtap, rowmask, ptr, cursor, need, dyt, entries, total;
  int caps, capy, capx, tw, cw, geom_warps;
  size_t entry_bytes;
  unsigned long long entry_cap;
  long long pixels;
};

inline size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

// `entries`: capacity of the entry array (fixed sampling ratios: ignored, the exact bound
// R * capy * capx is used)
PullLayout pull_layout(const RoiAlignArgs& a, bool deterministic, unsigned long long entries) {
  PullLayout L;
  const int G = a.sampling_ratio;
  const bool adaptive = G <= 0;
  const size_t R = (size_t)a.num_rois;
  const int cwork = a.c_count > 0 ? a.c_count : a.channels;
  if (adaptive) {
    // valid samples per axis: spacing bin / ceil(bin) > 1/2 over [-1, extent], or `pooled` samples when bin <= 1
    const int sy = 2 * a.height + a.pooled_h + 4, sx = 2 * a.width + a.pooled_w + 4;
    L.caps = sy > sx ? sy : sx;
    L.capy = 2 * sy;
    L.capx = 2 * sx;
    L.entry_cap = entries;
  } else {
    const int sy = a.pooled_h * G, sx = a.pooled_w * G;
    L.caps = sy > sx ? sy : sx;
    L.capy = 2 * sy;
    L.capx = 2 * sx;
    L.entry_cap = (unsigned long long)R * L.capy * L.capx;
  }
  L.entry_bytes = (deterministic && adaptive) ? sizeof(uint4) : sizeof(uint2);
  // grouped lists: never more entries than per-pixel ones (an entry stands for >= 1 of them)
  if (!deterministic && pull_grouped_enabled()) L.entry_bytes = sizeof(GroupEntry);
  const size_t per_warp = axis_scratch_bytes(L.caps);
  L.geom_warps = (int)((200 * 1024) / per_warp);
  if (L.geom_warps > kMaxGeomWarps) L.geom_warps = kMaxGeomWarps;
  if (L.geom_warps < 1) L.geom_warps = 1;
  L.tw = (a.num_rois >> 5) + a.num_images + 1;
  L.cw = (cwork + 3) & ~3;
  L.pixels = (long long)a.num_images * a.height * a.width;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    const size_t at = off;
    off += align256(bytes);
    return at;
  };
  L.img_start = take((size_t)(a.num_images + 2) * sizeof(int));
  L.order = take(R * sizeof(int));
  L.gbox = take(R * sizeof(uint4));
  L.yidx = take(R * (size_t)(a.height + 2) * sizeof(unsigned short));
  L.xidx = take(R * (size_t)(a.width + 2) * sizeof(unsigned short));
  L.ytap = take(R * (size_t)L.capy * sizeof(uint2));
  L.xtap = take(R * (size_t)L.capx * sizeof(uint2));
  L.rowmask = take((size_t)a.height * L.tw * sizeof(unsigned));
  L.ptr = take((size_t)L.pixels * sizeof(uint2));
  L.cursor = take(sizeof(unsigned));
  L.need = take(sizeof(unsigned long long));
  L.dyt = take(R * (size_t)a.pooled_h * a.pooled_w * L.cw * sizeof(float));
  L.entries = take((size_t)L.entry_cap * L.entry_bytes);
  L.total = off;
  return L;
}

}  // namespace

bool pull_supported(const RoiAlignArgs& a) {
  const int G = a.sampling_ratio;
  if (a.num_images < 1 || a.num_images > 65535) return false;
  if (a.pooled_h > 64 || a.pooled_w > 64) return false;
  if (G > 0) {
    if (a.pooled_h * G > kMaxFixedSamples || a.pooled_w * G > kMaxFixedSamples) return false;
    if (a.height > 4095 || a.width > 4095) return false;
    if ((long long)a.num_rois * 4 * a.pooled_h * G * a.pooled_w * G >= (1LL << 31)) return false;
  } else {
    if (a.height > kMaxAdaptiveExtent || a.width > kMaxAdaptiveExtent) return false;
  }
  if (a.pooled_h * a.pooled_w > 1600) return false;       // the transpose block: 32 channels x bins in shared memory
  const int cwork = a.c_count > 0 ? a.c_count : a.channels;
  if (cwork < 1) return false;
  // entry offsets and list positions are 32-bit
  if ((long long)a.num_rois * a.pooled_h * a.pooled_w * ((cwork + 3) & ~3) * 4 >= (1LL << 32)) return false;
  const long long segs = ((long long)a.width + kPullCols - 1) / kPullCols, bands = ((long long)a.height + kPullRows - 1) / kPullRows;
  if (segs * bands * a.num_images >= (1LL << 31)) return false;
  return true;
}

size_t pull_bwd_workspace(const RoiAlignArgs& a, int deterministic, unsigned long long entries) {
  return pull_layout(a, deterministic != 0, entries).total;
}

// Fixed sampling ratios: everything is launched, nothing is read back.  Adaptive grids: the number
// of entries is only known once the geometry kernel has run; the host waits for it (one 8-byte
// read-back).  If the workspace has no room for that many entries, nothing else is launched,
// *need_entries says how many are needed and the caller comes back with a larger workspace
// (~0ull: more than this formulation can index).
cudaError_t launch_bwd_pull(const RoiAlignArgs& a, void* workspace, unsigned long long entry_cap, const float* dy,
                            float* dx, int accumulate, int deterministic, int num_sms, cudaStream_t stream,
                            unsigned long long* need_entries) {
  *need_entries = 0;
  const int cwork = a.c_count > 0 ? a.c_count : a.channels;
  if (cwork <= 0 || a.num_images <= 0) return cudaSuccess;
  const bool det = deterministic != 0, adaptive = a.sampling_ratio <= 0;
  const PullLayout L = pull_layout(a, det, entry_cap);
  char* base = static_cast<char*>(workspace);
  PullParams p;
  p.rois = a.rois;
  p.num_rois = a.num_rois;
  p.num_images = a.num_images;
  p.channels = a.channels;
  p.c_work = cwork;
  p.cw = L.cw;
  p.height = a.height;
  p.width = a.width;
  p.pooled_h = a.pooled_h;
  p.pooled_w = a.pooled_w;
  p.grid = a.sampling_ratio;
  p.scale = a.scale;
  p.aligned = a.aligned;
  p.accumulate = accumulate;
  p.caps = L.caps;
  p.capy = L.capy;
  p.capx = L.capx;
  p.geom_warps = L.geom_warps;
  p.tw = L.tw;
  p.pixels = L.pixels;
  int* img_start = reinterpret_cast<int*>(base + L.img_start);
  int* order = reinterpret_cast<int*>(base + L.order);
  p.img_start = img_start;
  p.order = order;
  p.gbox = reinterpret_cast<uint4*>(base + L.gbox);
  p.yidx = reinterpret_cast<unsigned short*>(base + L.yidx);
  p.xidx = reinterpret_cast<unsigned short*>(base + L.xidx);
  p.ytap = reinterpret_cast<uint2*>(base + L.ytap);
  p.xtap = reinterpret_cast<uint2*>(base + L.xtap);
  p.rowmask = reinterpret_cast<unsigned*>(base + L.rowmask);
  p.ptr = reinterpret_cast<uint2*>(base + L.ptr);
  p.cursor = reinterpret_cast<unsigned*>(base + L.cursor);
  p.need = reinterpret_cast<unsigned long long*>(base + L.need);
  p.entries = base + L.entries;
  p.entry_cap = (unsigned)(L.entry_cap > 0xffffffffull ? 0xffffffffull : L.entry_cap);
  p.dyt = reinterpret_cast<float*>(base + L.dyt);
  p.dy = dy;
  p.dy_dtype = a.top_dtype;
  p.dy_layout = a.top_layout;
  p.dx = dx;

  cudaError_t e = cudaMemsetAsync(p.need, 0, sizeof(unsigned long long), stream);
  if (e != cudaSuccess) return e;
  e = launch_roi_group(a, img_start, order, stream);
  if (e != cudaSuccess) return e;
  const int B = a.pooled_h * a.pooled_w;
  if (a.num_rois > 0) {
    const unsigned gblocks = (unsigned)((a.num_rois + L.geom_warps - 1) / L.geom_warps);
    const size_t gsmem = (size_t)L.geom_warps * axis_scratch_bytes(L.caps);
    if (det) {
      if (gsmem > 48 * 1024)
        cudaFuncSetAttribute(pull_geom_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gsmem);
      chained_launch(pull_geom_kernel<true>, gblocks, L.geom_warps * 32, gsmem, stream, p);
    } else {
      if (gsmem > 48 * 1024)
        cudaFuncSetAttribute(pull_geom_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gsmem);
      chained_launch(pull_geom_kernel<false>, gblocks, L.geom_warps * 32, gsmem, stream, p);
    }
    count_launch();
  }
  if (adaptive && a.num_rois > 0) {
    unsigned long long need = 0;
    e = cudaMemcpyAsync(&need, p.need, sizeof(need), cudaMemcpyDeviceToHost, stream);
    if (e != cudaSuccess) return e;
    e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) return e;
    if (need >= 0xfffffff0ull) {
      *need_entries = ~0ull;
      return cudaSuccess;
    }
    if (need > L.entry_cap) {
      *need_entries = need;
      return cudaSuccess;
    }
  }
  {
    const long long warps = (long long)L.tw * ((a.height + 31) / 32);
    chained_launch(pull_rowmask_kernel, (unsigned)((warps + 7) / 8), 256, 0, stream, p);
    count_launch();
  }
  const long long segs = (a.width + 31) / 32;
  const bool wide = det && adaptive;
  const bool grouped = !det && pull_grouped_enabled();
  {
    // per-pixel lists + channel-last copy of dy, one grid
    const long long lwarps = (long long)a.num_images * a.height * segs;
    const unsigned lblocks = (unsigned)((lwarps + 7) / 8);
    const int tc = B <= 64 ? 128 : 32;
    unsigned tblocks = (unsigned)(a.num_rois * ((L.cw + tc - 1) / tc));
    size_t tsmem = a.num_rois > 0 ? (size_t)tc * (B | 1) * sizeof(float) : 0;
    // a channel-last dy with whole rows (cw == C == the channels worked on): fp32 is the copy the
    // main kernel wants -- it reads dy in place, no copy blocks at all; half precision is widened flat
    p.widen_flat = 0;
    // List blocks are long (a dependent walk over the ROIs of the row), copy blocks short.  While the list
    // blocks are few (up to three waves of resident CTAs) they go first, so that none of them starts late
    // and is left running alone at the end (C2 backward 137 -> 129 us, C3 1.61 -> 1.51 ms, 8 images of
    // C4: 413 -> 404 us); a long grid keeps them spread over the copy blocks (16 images: 782 against
    // 787 us, C4: 2.99 against 3.07 ms with the list blocks first; profiles/r2_knob_experiments.txt).
    // MOBULA_TABLES_Q = cap of the interleaving period (0 = none), see pull_tables_kernel.
    const int q_env = [] { const char* e = getenv("MOBULA_TABLES_Q"); return e ? atoi(e) : -1; }();   // per call: the tests switch it
    p.tables_q = q_env >= 0 ? q_env : (lblocks <= 3u * (unsigned)(num_sms > 0 ? num_sms : 148) * MOBULA_TABLES_MINBLOCKS ? 1 : 0);
    if (a.top_layout != 0 && L.cw == a.channels && cwork == a.channels) {
      if (a.top_dtype == 0 && (reinterpret_cast<uintptr_t>(dy) & 15u) == 0) {
        p.dyt = const_cast<float*>(dy);
        tblocks = 0;
      } else if (a.top_dtype != 0) {
        p.widen_flat = 1;
        const unsigned long long elems = (unsigned long long)a.num_rois * B * L.cw;
        tblocks = (unsigned)((elems + 2047) / 2048);
      }
      if (!tblocks || p.widen_flat) tsmem = 0;
    }
#define PULL_TABLES(DET, WIDE, TC, GRP)                                                                            \
  do {                                                                                                             \
    auto kk = pull_tables_kernel<DET, WIDE, TC, GRP>;                                                              \
    if (tsmem > 48 * 1024) /* per device, cheap: set on every call */                                              \
      cudaFuncSetAttribute(kk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsmem);                           \
    chained_launch(kk, lblocks + tblocks, 256, tsmem, stream, p, lblocks, tblocks);                                \
  } while (0)
#define PULL_TABLES_TC(TC)                        \
  do {                                            \
    if (wide) PULL_TABLES(true, true, TC, false);        \
    else if (det) PULL_TABLES(true, false, TC, false);   \
    else if (grouped) PULL_TABLES(false, false, TC, true); \
    else PULL_TABLES(false, false, TC, false);           \
  } while (0)
    if (tc == 128) PULL_TABLES_TC(128);
    else PULL_TABLES_TC(32);
#undef PULL_TABLES_TC
#undef PULL_TABLES
    count_launch();
  }

  const int G = a.sampling_ratio, cnt_i = adaptive ? 1 : G * G;
  const float count = (float)cnt_i;
  const int pow2 = (cnt_i & (cnt_i - 1)) == 0 ? 1 : 0;
  const float inv = 1.0f / count;
  const int cpl = (L.cw >= 128) ? 4 : (L.cw >= 64 ? 2 : 1);
  const long long bands = (a.height + kPullRows - 1) / kPullRows, tile_cols = (a.width + kPullCols - 1) / kPullCols;
  const dim3 grid((unsigned)(tile_cols * bands * a.num_images), (unsigned)((L.cw + 32 * cpl - 1) / (32 * cpl)));
  // MOBULA_PULL_UNROLL = terms per group (4 or 8): gradient loads in flight per warp
  static const int unroll = [] { const char* e = getenv("MOBULA_PULL_UNROLL"); return e ? atoi(e) : 8; }();
  // MOBULA_PULL_CARVEOUT = shared-memory share of the L1 / shared-memory array in percent (experiment knob)
  static const int carveout = [] { const char* e = getenv("MOBULA_PULL_CARVEOUT"); return e ? atoi(e) : -1; }();
#define PULL_LAUNCH(CPL, DET, ACC, WIDE)                                                  \
  do {                                                                                    \
    auto k8 = pull_bwd_kernel<CPL, DET, ACC, 8, WIDE>;                                    \
    auto k4 = pull_bwd_kernel<CPL, DET, ACC, 4, WIDE>;                                    \
    auto kk = unroll == 4 ? k4 : k8;                                                      \
    const size_t smem = (size_t)(kPullThreads / 32) * (8 * 32 * CPL * sizeof(float) +    \
                                                       2 * kEntChunk * (WIDE ? 16 : 8));  \
    if (smem > 48 * 1024) cudaFuncSetAttribute(kk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    if (carveout >= 0) cudaFuncSetAttribute(kk, cudaFuncAttributePreferredSharedMemoryCarveout, carveout); \
    chained_launch(kk, grid, dim3(kPullThreads), smem, stream, p, count, inv, pow2);      \
  } while (0)
#define PULL_MODE(CPL)                                              \
  do {                                                              \
    if (wide) {                                                     \
      if (accumulate) PULL_LAUNCH(CPL, true, true, true);           \
      else PULL_LAUNCH(CPL, true, false, true);                     \
    } else if (det) {                                               \
      if (accumulate) PULL_LAUNCH(CPL, true, true, false);          \
      else PULL_LAUNCH(CPL, true, false, false);                    \
    } else {                                                        \
      if (accumulate) PULL_LAUNCH(CPL, false, true, false);         \
      else PULL_LAUNCH(CPL, false, false, false);                   \
    }                                                               \
  } while (0)
  // MOBULA_PULL8_VARIANT: 0 = 4 lines per batch, 1 = 8 lines, 2 = 4 lines pipelined (two CTAs per SM each),
  // 3 = 4 lines, three CTAs per SM
  const int variant = [] { const char* e = getenv("MOBULA_PULL8_VARIANT"); return e ? atoi(e) : MOBULA_PULL8_VARIANT_DEFAULT; }();
#define PULL8_KERNEL(CPL, ACC) \
  (variant == 1 ? pull8_bwd_kernel<CPL, ACC, 8, 2, false> : variant == 2 ? pull8_bwd_kernel<CPL, ACC, 4, 2, true> : \
   variant == 3 ? pull8_bwd_kernel<CPL, ACC, 4, 3, false> : pull8_bwd_kernel<CPL, ACC, 4, 2, false>)
#define PULL8_LAUNCH(CPL)                                                                         \
  do {                                                                                            \
    const size_t ring = 2 * kG8Chunk * sizeof(GroupEntry), tile = 8 * 32 * CPL * sizeof(float);   \
    const size_t smem = (size_t)(kPullThreads / 32) * (ring > tile ? ring : tile);                \
    auto kk = accumulate ? PULL8_KERNEL(CPL, true) : PULL8_KERNEL(CPL, false);                    \
    if (smem > 48 * 1024) cudaFuncSetAttribute(kk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    chained_launch(kk, grid, dim3(kPullThreads), smem, stream, p);                                \
  } while (0)
  if (grouped) {
    if (cpl == 4) PULL8_LAUNCH(4);
    else if (cpl == 2) PULL8_LAUNCH(2);
    else PULL8_LAUNCH(1);
  } else if (cpl == 4) PULL_MODE(4);
  else if (cpl == 2) PULL_MODE(2);
  else PULL_MODE(1);
#undef PULL8_LAUNCH
#undef PULL8_KERNEL
#undef PULL_MODE
#undef PULL_LAUNCH
  count_launch();
  return cudaGetLastError();
}

}  // namespace mobula_b200
