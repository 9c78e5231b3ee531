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

constexpr int kMaxSamples = 64;               // samples per axis: PH * G, PW * G
constexpr int kMaxTaps = 2 * kMaxSamples;     // taps per axis
constexpr int kGeomWarps = 4;

struct PullParams {
  const float* rois;
  int num_rois, num_images;
  int channels;          // channel stride of dy / dx
  int c_work;            // channels this call works on
  int cw;                // channels per row of the channel-last copy of dy (c_work rounded up to 4)
  int height, width, pooled_h, pooled_w, grid;
  float scale;
  int aligned;
  int accumulate;
  int capy, capx;        // taps per ROI and axis: 2 * PH * G, 2 * PW * G
  int tw;                // row-mask words per feature-map row
  long long pixels;      // N * H * W
  const int* img_start;  // [N + 2]
  const int* order;      // [R] ROI rows grouped by image, ascending inside an image
  uint2* gbox;           // [R] by grouped position: rows min | max << 16, columns min | max << 16
  unsigned char* yidx;   // [R][H + 2] by grouped position: first tap of row (ymin + q)
  unsigned char* xidx;   // [R][W + 2]
  uint2* ytap;           // [R][capy] taps sorted by row: x = row | info << 16, y = weight
  uint2* xtap;           // [R][capx]
  unsigned* rowmask;     // [H][tw]
  uint2* ptr;            // [pixels] first entry and number of entries of every pixel
  unsigned* cursor;      // entries handed out so far (zeroed by the geometry kernel)
  uint2* entries;        // x = byte offset of the bin's channel row in dyt, ((roi * PH + ph) * PW + pw) * cw * 4 (a
                         // multiple of 16) | pixel & 7 | 8 on the last entry of the pixel; y = weight
  float* dyt;            // [R * PH * PW][cw] channel-last copy of dy
  const float* dy;
  float* dx;
};

// first row-mask word of image n: its ROIs sit at grouped positions [img_start[n], img_start[n + 1])
// and use the words (position >> 5) + n, so that two images never share a word
__device__ __forceinline__ int mask_word0(const int* __restrict__ img_start, int n) {
  return (__ldg(img_start + n) >> 5) + n;
}

// ---- geometry: one warp per ROI (grouped position) ------------------------------------------
// One axis: the taps of the P * G samples, sorted by (pixel, sample, low/high) [DET] or merged per
// (pixel, bin) [!DET], and the index "pixel -> first tap".  Returns min | max << 16 of the touched
// pixels (0xffff | 0: none).
template <bool DET>
__device__ unsigned pull_axis(float start, float bin, int pooled, int grid, int extent, float wscale, int lane,
                              unsigned* s_key, float* s_w, unsigned* s_skey, float* s_sw, uint2* __restrict__ taps,
                              unsigned char* __restrict__ idx) {
  const int ns = pooled * grid;
  // raw taps: sample s -> slots 2 s (low) and 2 s + 1 (high); key = pixel << 8 | s << 1 | high
  for (int s = lane; s < kMaxSamples; s += 32) {
    unsigned k0 = 0xffffffffu, k1 = 0xffffffffu;
    float w0 = 0.0f, w1 = 0.0f;
    if (s < ns) {
      const int p = s / grid, i = s - p * grid;
      const AxisTap t = axis_decode(sample_pos(start, p, bin, i, grid), extent);
      if (t.valid) {
        k0 = ((unsigned)t.lo << 8) | ((unsigned)s << 1);
        k1 = ((unsigned)t.hi << 8) | ((unsigned)s << 1) | 1u;
        w0 = t.wh;      // weight of the low tap (1 - fraction)
        w1 = t.wl;      // weight of the high tap (the fraction)
        if (!DET && t.hi == t.lo) k1 = 0xffffffffu;      // clamped sample: its high tap adds 0 to the same pixel
      }
    }
    s_key[2 * s] = k0;
    s_key[2 * s + 1] = k1;
    s_w[2 * s] = w0;
    s_w[2 * s + 1] = w1;
  }
  __syncwarp();
  // rank sort (valid keys are unique)
  const int nall = 2 * ns;
  int nvalid = 0;
  for (int e0 = 0; e0 < nall; e0 += 32) {
    const int e = e0 + lane;
    const unsigned key = e < nall ? s_key[e] : 0xffffffffu;
    const bool valid = key != 0xffffffffu;
    if (valid) {
      int rank = 0;
      for (int k = 0; k < nall; ++k) rank += s_key[k] < key ? 1 : 0;
      s_skey[rank] = key;
      s_sw[rank] = s_w[e];
    }
    nvalid += __popc(__ballot_sync(0xffffffffu, valid));
  }
  __syncwarp();
  int nfinal = nvalid;
  if (DET) {
    for (int e = lane; e < nvalid; e += 32) {
      const unsigned key = s_skey[e];
      taps[e] = make_uint2((key >> 8) | ((key & 0xffu) << 16), __float_as_uint(s_sw[e]));
    }
  } else {
    // runs of one (pixel, bin) become one tap with the summed weight
    int out_base = 0;
    for (int e0 = 0; e0 < nvalid; e0 += 32) {
      const int e = e0 + lane;
      bool head = false;
      unsigned id = 0;
      if (e < nvalid) {
        const unsigned key = s_skey[e];
        id = ((key >> 8) << 8) | (((key & 0xffu) >> 1) / (unsigned)grid);
        if (e == 0) head = true;
        else {
          const unsigned pk = s_skey[e - 1];
          head = (((pk >> 8) << 8) | (((pk & 0xffu) >> 1) / (unsigned)grid)) != id;
        }
      }
      const unsigned heads = __ballot_sync(0xffffffffu, head);
      if (head) {
        float w = s_sw[e];
        for (int k = e + 1; k < nvalid; ++k) {
          const unsigned nk = s_skey[k];
          if ((((nk >> 8) << 8) | (((nk & 0xffu) >> 1) / (unsigned)grid)) != id) break;
          w += s_sw[k];
        }
        const int pos = out_base + __popc(heads & ((1u << lane) - 1u));
        taps[pos] = make_uint2((id >> 8) | ((id & 0xffu) << 16), __float_as_uint(w * wscale));
        s_key[pos] = id >> 8;           // pixel of the merged tap (for the index below)
      }
      out_base += __popc(heads);
    }
    nfinal = out_base;
    __syncwarp();
  }
  if (nfinal == 0) return 0xffffu;
  const unsigned* pixkey = DET ? s_skey : s_key;
  const int shift = DET ? 8 : 0;
  const int pmin = (int)(pixkey[0] >> shift), pmax = (int)(pixkey[nfinal - 1] >> shift);
  // idx[q] = first tap whose pixel is >= pmin + q, q = 0 .. pmax - pmin + 1
  for (int q = lane; q <= pmax - pmin + 1; q += 32) {
    const unsigned want = (unsigned)(pmin + q);
    int lo = 0, hi = nfinal;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if ((pixkey[mid] >> shift) < want) lo = mid + 1;
      else hi = mid;
    }
    idx[q] = (unsigned char)lo;
  }
  __syncwarp();
  return (unsigned)pmin | ((unsigned)pmax << 16);
}

template <bool DET>
__global__ void __launch_bounds__(kGeomWarps * 32) pull_geom_kernel(const PullParams p) {
  __shared__ unsigned s_key[kGeomWarps][kMaxTaps];
  __shared__ float s_w[kGeomWarps][kMaxTaps];
  __shared__ unsigned s_skey[kGeomWarps][kMaxTaps];
  __shared__ float s_sw[kGeomWarps][kMaxTaps];
  ptx::chain_enter();
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int pos = blockIdx.x * kGeomWarps + wib;
  if (blockIdx.x == 0 && threadIdx.x == 0) *p.cursor = 0u;
  if (pos >= p.num_rois) return;
  uint2 box = make_uint2(0xffffu, 0xffffu);
  if (pos < __ldg(p.img_start + p.num_images)) {       // rows with a batch index outside [0, N) scatter nothing
    const int r = __ldg(p.order + pos);
    const RoiGeom g = roi_decode(p.rois + (size_t)r * 5, p.scale, p.pooled_h, p.pooled_w, p.grid, p.aligned != 0);
    const float inv_count = __fdiv_rn(1.0f, g.count);
    const unsigned yr = pull_axis<DET>(g.start_h, g.bin_h, p.pooled_h, p.grid, p.height, 1.0f, lane, s_key[wib],
                                       s_w[wib], s_skey[wib], s_sw[wib], p.ytap + (size_t)pos * p.capy,
                                       p.yidx + (size_t)pos * (p.height + 2));
    __syncwarp();
    const unsigned xr = pull_axis<DET>(g.start_w, g.bin_w, p.pooled_w, p.grid, p.width, inv_count, lane, s_key[wib],
                                       s_w[wib], s_skey[wib], s_sw[wib], p.xtap + (size_t)pos * p.capx,
                                       p.xidx + (size_t)pos * (p.width + 2));
    if (yr != 0xffffu && xr != 0xffffu) box = make_uint2(yr, xr);
  }
  if (lane == 0) p.gbox[pos] = box;
}

// ---- row mask: one warp per (feature-map row, mask word), lane = ROI ---------------------------
__global__ void __launch_bounds__(256) pull_rowmask_kernel(const PullParams p) {
  ptx::chain_enter();
  const int lane = threadIdx.x & 31;
  const long long wid = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (wid >= (long long)p.height * p.tw) return;
  const int y = (int)(wid / p.tw), wg = (int)(wid - (long long)y * p.tw);
  // image of this word: the last n with mask_word0(n) <= wg
  int lo = 0, hi = p.num_images;        // answer in [lo, hi]; hi == num_images: behind the last image
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (mask_word0(p.img_start, mid) <= wg) lo = mid;
    else hi = mid - 1;
  }
  const int n = lo;
  bool hit = false;
  if (n < p.num_images) {
    const int pos = ((wg - n) << 5) + lane;
    if (pos >= __ldg(p.img_start + n) && pos < __ldg(p.img_start + n + 1)) {
      const unsigned yr = __ldg(&p.gbox[pos].x);
      hit = (int)(yr & 0xffffu) <= y && y <= (int)(yr >> 16);
    }
  }
  const unsigned m = __ballot_sync(0xffffffffu, hit);
  if (lane == 0) p.rowmask[wid] = m;
}

// ---- per-pixel lists: one thread per pixel, a warp = 32 consecutive pixels of one row ----------
// walk<FILL = false> counts the entries of the lane's pixel, walk<true> writes them: both visit
// the ROIs of the row in ascending order, so a list is written in its final order by one thread.
template <bool FILL, bool DET>
__device__ __forceinline__ unsigned pull_walk(const PullParams& p, int n, int y, int x0, int lane, uint2* out) {
  const int x = x0 + lane;
  const bool in = x < p.width;
  const int w_lo = mask_word0(p.img_start, n), w_hi = mask_word0(p.img_start, n + 1);
  const int B = p.pooled_h * p.pooled_w, G = p.grid;
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
      unsigned my_r = 0, my_x = 0;               // ra | rb << 8, xmin | xmax << 16
      bool cand = (m >> lane) & 1u;
      if (cand) {
        const int pos = ((wg - n) << 5) + lane;
        const uint2 box = __ldg(p.gbox + pos);
        const int xmin = (int)(box.y & 0xffffu), xmax = (int)(box.y >> 16);
        cand = !(xmax < x0 || xmin > x0 + 31);
        if (cand) {
          const unsigned char* yi = p.yidx + (size_t)pos * (p.height + 2) + (y - (int)(box.x & 0xffffu));
          const unsigned ra = __ldg(yi), rb = __ldg(yi + 1);
          cand = ra != rb;
          my_r = ra | (rb << 8);
          my_x = box.y;
        }
      }
      unsigned live = __ballot_sync(0xffffffffu, cand);
      while (live) {
        const int b = __ffs(live) - 1;
        live &= live - 1;
        const int pos = ((wg - n) << 5) + b;
        const unsigned rr = __shfl_sync(0xffffffffu, my_r, b), xx = __shfl_sync(0xffffffffu, my_x, b);
        const int ra = (int)(rr & 0xffu), rb = (int)(rr >> 8);
        const int xmin = (int)(xx & 0xffffu), xmax = (int)(xx >> 16);
        int ca = 0, cb = 0;
        if (in && x >= xmin && x <= xmax) {
          const unsigned char* xi = p.xidx + (size_t)pos * (p.width + 2) + (x - xmin);
          ca = __ldg(xi);
          cb = __ldg(xi + 1);
        }
        if (!FILL) {
          total += (unsigned)((rb - ra) * (cb - ca));
          continue;
        }
        if (ca == cb) continue;
        const unsigned base = (unsigned)__ldg(p.order + pos) * (unsigned)B;
        const uint2* yt = p.ytap + (size_t)pos * p.capy;
        const uint2* xt = p.xtap + (size_t)pos * p.capx;
        const unsigned cwb = (unsigned)(p.cw * 4), pxbits = (unsigned)(x & 7);     // cwb is a multiple of 16
        if (!DET) {
          for (int a = ra; a < rb; ++a) {
            const uint2 ya = __ldg(yt + a);
            const unsigned ph = ya.x >> 16;
            const float wy = __uint_as_float(ya.y);
            for (int c = ca; c < cb; ++c) {
              const uint2 xc = __ldg(xt + c);
              *out++ = make_uint2((base + ph * (unsigned)p.pooled_w + (xc.x >> 16)) * cwb | pxbits,
                                  __float_as_uint(wy * __uint_as_float(xc.y)));
            }
          }
        } else {
          // the reference's order inside one ROI: bin row, bin column, sample row, sample column,
          // then the four taps (low/low, low/high, high/low, high/high); ROIAlign.cpp:126-157
          int a0 = ra;
          while (a0 < rb) {
            const unsigned ph = ((__ldg(&yt[a0].x) >> 17) & 0x7fu) / (unsigned)G;
            int a1 = a0 + 1;
            while (a1 < rb && ((__ldg(&yt[a1].x) >> 17) & 0x7fu) / (unsigned)G == ph) ++a1;
            int c0 = ca;
            while (c0 < cb) {
              const unsigned pw = ((__ldg(&xt[c0].x) >> 17) & 0x7fu) / (unsigned)G;
              int c1 = c0 + 1;
              while (c1 < cb && ((__ldg(&xt[c1].x) >> 17) & 0x7fu) / (unsigned)G == pw) ++c1;
              const unsigned id = (base + ph * (unsigned)p.pooled_w + pw) * cwb | pxbits;
              int a = a0;
              while (a < a1) {                                   // taps of one sample row: 1 or 2
                const unsigned sa = __ldg(&yt[a].x) >> 17;
                const int a2 = (a + 1 < a1 && (__ldg(&yt[a + 1].x) >> 17) == sa) ? a + 2 : a + 1;
                int c = c0;
                while (c < c1) {                                 // taps of one sample column
                  const unsigned sc = __ldg(&xt[c].x) >> 17;
                  const int c2 = (c + 1 < c1 && (__ldg(&xt[c + 1].x) >> 17) == sc) ? c + 2 : c + 1;
                  for (int aa = a; aa < a2; ++aa) {
                    const float wy = __uint_as_float(__ldg(&yt[aa].y));
                    for (int cc = c; cc < c2; ++cc)
                      *out++ = make_uint2(id, __float_as_uint(__fmul_rn(wy, __uint_as_float(__ldg(&xt[cc].y)))));
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
template <bool DET>
__device__ __forceinline__ void pull_list_block(const PullParams& p, unsigned block) {
  const int lane = threadIdx.x & 31;
  const int segs = (p.width + 31) >> 5;
  const long long wid = (long long)block * 8 + (threadIdx.x >> 5);
  if (wid >= (long long)p.num_images * p.height * segs) return;
  const int seg = (int)(wid % segs);
  const int row = (int)(wid / segs);               // n * H + y
  const int n = row / p.height, y = row - n * p.height;
  const int x0 = seg * 32;
  const unsigned total = pull_walk<false, DET>(p, n, y, x0, lane, nullptr);
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
  const unsigned begin = base + incl - total;
  if (x0 + lane < p.width) p.ptr[(long long)row * p.width + x0 + lane] = make_uint2(begin, total);
  if (warp_total) pull_walk<true, DET>(p, n, y, x0, lane, p.entries + begin);
  if (total) p.entries[begin + total - 1].x |= 8u;       // the last entry of the pixel (written by this thread)
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
  const float* src = p.dy + ((size_t)r * p.channels + c0) * B;
  const int total = nc * B;
  if (pitch == B && (reinterpret_cast<uintptr_t>(src) & 15u) == 0) {
    // odd B: the shared-memory image is the source run itself, copied 16 bytes at a time
    const int quads = total >> 2;
    float4* dst4 = reinterpret_cast<float4*>(s_t);
    for (int q = threadIdx.x; q < quads; q += 256) dst4[q] = __ldg(reinterpret_cast<const float4*>(src) + q);
    for (int i = (quads << 2) + threadIdx.x; i < total; i += 256) s_t[i] = __ldg(src + i);
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
  float* dst = p.dyt + (size_t)r * B * p.cw + c0;
  const int c = threadIdx.x & (kTrChannels - 1);
  if (c < ncw) {
    const float* mine = s_t + c * pitch;
    const bool real = c < nc;
    for (int bb = threadIdx.x / kTrChannels; bb < B; bb += 256 / kTrChannels)
      dst[(size_t)bb * p.cw + c] = real ? mine[bb] : 0.0f;
  }
}

// The lists and the channel-last copy in ONE grid: building lists is latency-bound index work,
// the copy is pure bandwidth, and neither depends on the other -- side by side the copy is free.
// Block roles are interleaved (`lblocks` list blocks, `tblocks` copy blocks) so both progress.
template <bool DET, int kTrChannels>
__global__ void __launch_bounds__(256) pull_tables_kernel(const PullParams p, const unsigned lblocks,
                                                          const unsigned tblocks) {
  extern __shared__ float s_t[];      // copy blocks: [kTrChannels][B | 1]
  ptx::chain_enter();
  const unsigned i = blockIdx.x;
  const unsigned few = min(lblocks, tblocks), many = max(lblocks, tblocks);
  bool is_few = false;
  unsigned idx = i;
  if (few > 0) {
    const unsigned q = many / few + 1;               // one block of the smaller set every q blocks
    const unsigned few_before = min(few, (i + q - 1) / q);
    is_few = i % q == 0 && i / q < few;
    idx = is_few ? i / q : i - few_before;
  }
  const bool list_role = (lblocks <= tblocks) == is_few;      // the smaller set is the list set iff lblocks <= tblocks
  if (few == 0 ? lblocks > 0 : list_role) pull_list_block<DET>(p, idx);
  else pull_transpose_block<kTrChannels>(p, idx, s_t);
}

// ---- the main kernel ------------------------------------------------------------------------
// CTA = 2 rows x 32 pixels x (32 * CPL) channels; warp w: row w / 4, pixels 8 (w % 4) .. + 7.
// The entries of a warp's eight pixels are one contiguous run of the list: the warp brings them
// to shared memory 64 at a time (coalesced) and reads them back as broadcasts, so the only
// global load on the dependent path of a term is its gradient line.
constexpr int kPullThreads = 256;
constexpr int kPullRows = 2;
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

// One group of U terms: entries from the warp's shared-memory ring, every gradient load issued
// before the first use.
template <int CPL, bool DET, int U>
__device__ __forceinline__ void term_group(float (&acc)[CPL], const uint2* se, const char* dyt, float count,
                                           float inv, int pow2) {
  uint2 en[U];
  float g[U][CPL];
#pragma unroll
  for (int u = 0; u < U; ++u) en[u] = se[u];
#pragma unroll
  for (int u = 0; u < U; ++u) load_vec<CPL>(reinterpret_cast<const float*>(dyt + en[u].x), g[u]);
#pragma unroll
  for (int u = 0; u < U; ++u) add_term<CPL, DET>(acc, g[u], __uint_as_float(en[u].y), count, inv, pow2);
}

template <int CPL, bool DET, bool ACC, int U>
__global__ void __launch_bounds__(kPullThreads) pull_bwd_kernel(const PullParams p, const float count, const float inv,
                                                                const int pow2) {
  constexpr int CG = 32 * CPL;                        // channels per CTA
  // per warp: its 8 pixels x CG channels, [pixel][channel slot], and a ring of list entries
  __shared__ __align__(16) float s_tile[kPullThreads / 32][8 * CG];
  __shared__ uint2 s_ent[kPullThreads / 32][2 * kEntChunk];
  ptx::chain_enter();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int segs = (p.width + 31) >> 5, bands = (p.height + kPullRows - 1) / kPullRows;
  const int seg = blockIdx.x % segs;
  const int band = (blockIdx.x / segs) % bands;
  const int n = blockIdx.x / (segs * bands);
  const int c0 = blockIdx.y * CG;
  const int y = band * kPullRows + (warp >> 2);
  const int xw = seg * 32 + (warp & 3) * 8;
  if (y >= p.height || xw >= p.width) return;         // no block-wide barrier below: warps are independent
  const bool lane_on = c0 + lane * CPL < p.cw;
  const char* dyt = reinterpret_cast<const char*>(p.dyt + (lane_on ? c0 + lane * CPL : 0));
  float* tile = s_tile[warp];
  uint2* ring = s_ent[warp];

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
  uint2 nxt[kEntChunk / 32];
#pragma unroll
  for (int k = 0; k < kEntChunk / 32; ++k) {
    const unsigned i = e0 + k * 32 + lane;
    if (i < end_all) ring[k * 32 + lane] = __ldg(p.entries + i);
  }
  __syncwarp();
  int buf = 0;
  for (unsigned cbase = e0; cbase < end_all; cbase += kEntChunk, buf ^= 1) {
    const unsigned cn = min((unsigned)kEntChunk, end_all - cbase);
#pragma unroll
    for (int k = 0; k < kEntChunk / 32; ++k) {
      const unsigned i = cbase + kEntChunk + k * 32 + lane;
      nxt[k] = i < end_all ? __ldg(p.entries + i) : make_uint2(0u, 0u);
    }
    const uint2* se = ring + buf * kEntChunk;
    unsigned t = 0;
    for (; t + U <= cn; t += U) {
      uint2 en[U];
      float g[U][CPL];
#pragma unroll
      for (int u = 0; u < U; ++u) en[u] = se[t + u];
#pragma unroll
      for (int u = 0; u < U; ++u) load_vec<CPL>(reinterpret_cast<const float*>(dyt + (en[u].x & ~15u)), g[u]);
#pragma unroll
      for (int u = 0; u < U; ++u) {
        add_term<CPL, DET>(acc, g[u], __uint_as_float(en[u].y), count, inv, pow2);
        if (en[u].x & 8u) flush(en[u].x);
      }
    }
    for (; t < cn; ++t) {
      const uint2 en = se[t];
      float g[CPL];
      load_vec<CPL>(reinterpret_cast<const float*>(dyt + (en.x & ~15u)), g);
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

struct PullLayout {
  size_t img_start, order, gbox, yidx, xidx, ytap, xtap, rowmask, ptr, cursor, entries, dyt, total;
  int capy, capx, tw, cw;
  long long pixels;
};

inline size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

PullLayout pull_layout(const RoiAlignArgs& a) {
  PullLayout L;
  const int G = a.sampling_ratio;
  const size_t R = (size_t)a.num_rois;
  const int cwork = a.c_count > 0 ? a.c_count : a.channels;
  L.capy = 2 * a.pooled_h * G;
  L.capx = 2 * a.pooled_w * G;
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
  L.gbox = take(R * sizeof(uint2));
  L.yidx = take(R * (size_t)(a.height + 2));
  L.xidx = take(R * (size_t)(a.width + 2));
  L.ytap = take(R * (size_t)L.capy * sizeof(uint2));
  L.xtap = take(R * (size_t)L.capx * sizeof(uint2));
  L.rowmask = take((size_t)a.height * L.tw * sizeof(unsigned));
  L.ptr = take((size_t)L.pixels * sizeof(uint2));
  L.cursor = take(sizeof(unsigned));
  L.entries = take(R * (size_t)L.capy * L.capx * sizeof(uint2));
  L.dyt = take(R * (size_t)a.pooled_h * a.pooled_w * L.cw * sizeof(float));
  L.total = off;
  return L;
}

}  // namespace

bool pull_supported(const RoiAlignArgs& a) {
  const int G = a.sampling_ratio;
  if (G < 1 || a.num_images < 1 || a.num_images > 65535) return false;
  if (a.pooled_h * G > kMaxSamples || a.pooled_w * G > kMaxSamples) return false;
  if (a.height > 4095 || a.width > 4095) return false;
  if (a.pooled_h * a.pooled_w > 1600) return false;       // the transpose block: 32 channels x bins in shared memory
  const int cwork = a.c_count > 0 ? a.c_count : a.channels;
  if (cwork < 1) return false;
  // entry ids and list positions are 32-bit
  if ((long long)a.num_rois * a.pooled_h * a.pooled_w * ((cwork + 3) & ~3) * 4 >= (1LL << 32)) return false;
  if ((long long)a.num_rois * 4 * a.pooled_h * G * a.pooled_w * G >= (1LL << 31)) return false;
  const long long segs = ((long long)a.width + 31) / 32, bands = ((long long)a.height + kPullRows - 1) / kPullRows;
  if (segs * bands * a.num_images >= (1LL << 31)) return false;
  return true;
}

size_t pull_bwd_workspace(const RoiAlignArgs& a) { return pull_layout(a).total; }

cudaError_t launch_bwd_pull(const RoiAlignArgs& a, void* workspace, const float* dy, float* dx, int accumulate,
                            int deterministic, int num_sms, cudaStream_t stream) {
  (void)num_sms;
  const int cwork = a.c_count > 0 ? a.c_count : a.channels;
  if (cwork <= 0 || a.num_images <= 0) return cudaSuccess;
  const PullLayout L = pull_layout(a);
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
  p.capy = L.capy;
  p.capx = L.capx;
  p.tw = L.tw;
  p.pixels = L.pixels;
  int* img_start = reinterpret_cast<int*>(base + L.img_start);
  int* order = reinterpret_cast<int*>(base + L.order);
  p.img_start = img_start;
  p.order = order;
  p.gbox = reinterpret_cast<uint2*>(base + L.gbox);
  p.yidx = reinterpret_cast<unsigned char*>(base + L.yidx);
  p.xidx = reinterpret_cast<unsigned char*>(base + L.xidx);
  p.ytap = reinterpret_cast<uint2*>(base + L.ytap);
  p.xtap = reinterpret_cast<uint2*>(base + L.xtap);
  p.rowmask = reinterpret_cast<unsigned*>(base + L.rowmask);
  p.ptr = reinterpret_cast<uint2*>(base + L.ptr);
  p.cursor = reinterpret_cast<unsigned*>(base + L.cursor);
  p.entries = reinterpret_cast<uint2*>(base + L.entries);
  p.dyt = reinterpret_cast<float*>(base + L.dyt);
  p.dy = dy;
  p.dx = dx;

  cudaError_t e = launch_roi_group(a, img_start, order, stream);
  if (e != cudaSuccess) return e;
  const bool det = deterministic != 0;
  const int B = a.pooled_h * a.pooled_w;
  if (a.num_rois > 0) {
    const unsigned gblocks = (unsigned)((a.num_rois + kGeomWarps - 1) / kGeomWarps);
    if (det) chained_launch(pull_geom_kernel<true>, gblocks, kGeomWarps * 32, 0, stream, p);
    else chained_launch(pull_geom_kernel<false>, gblocks, kGeomWarps * 32, 0, stream, p);
    count_launch();
  }
  {
    const long long warps = (long long)a.height * L.tw;
    chained_launch(pull_rowmask_kernel, (unsigned)((warps + 7) / 8), 256, 0, stream, p);
    count_launch();
  }
  const long long segs = (a.width + 31) / 32;
  {
    // per-pixel lists + channel-last copy of dy, one grid
    const long long lwarps = (long long)a.num_images * a.height * segs;
    const unsigned lblocks = (unsigned)((lwarps + 7) / 8);
    const int tc = B <= 64 ? 128 : 32;
    const unsigned tblocks = (unsigned)(a.num_rois * ((L.cw + tc - 1) / tc));
    const size_t tsmem = a.num_rois > 0 ? (size_t)tc * (B | 1) * sizeof(float) : 0;
#define PULL_TABLES(DET, TC)                                                                                       \
  do {                                                                                                             \
    if (tsmem > 48 * 1024) /* per device, cheap: set on every call */                                              \
      cudaFuncSetAttribute(pull_tables_kernel<DET, TC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsmem);  \
    chained_launch(pull_tables_kernel<DET, TC>, lblocks + tblocks, 256, tsmem, stream, p, lblocks, tblocks);       \
  } while (0)
    if (tc == 128) {
      if (det) PULL_TABLES(true, 128);
      else PULL_TABLES(false, 128);
    } else {
      if (det) PULL_TABLES(true, 32);
      else PULL_TABLES(false, 32);
    }
#undef PULL_TABLES
    count_launch();
  }

  const int G = a.sampling_ratio, cnt_i = G * G;
  const float count = (float)cnt_i;
  const int pow2 = (cnt_i & (cnt_i - 1)) == 0 ? 1 : 0;
  const float inv = 1.0f / count;
  const int cpl = (L.cw >= 128) ? 4 : (L.cw >= 64 ? 2 : 1);
  const long long bands = (a.height + kPullRows - 1) / kPullRows;
  const dim3 grid((unsigned)(segs * bands * a.num_images), (unsigned)((L.cw + 32 * cpl - 1) / (32 * cpl)));
  // MOBULA_PULL_UNROLL = terms per group (4 or 8): gradient loads in flight per warp
  static const int unroll = [] { const char* e = getenv("MOBULA_PULL_UNROLL"); return e ? atoi(e) : 8; }();
  static const int carveout = [] { const char* e = getenv("MOBULA_PULL_CARVEOUT"); return e ? atoi(e) : -1; }();
#define PULL_LAUNCH(CPL, DET, ACC)                                                                                   \
  do {                                                                                                               \
    auto k8 = pull_bwd_kernel<CPL, DET, ACC, 8>;                                                                     \
    auto k4 = pull_bwd_kernel<CPL, DET, ACC, 4>;                                                                     \
    auto kk = unroll == 4 ? k4 : k8;                                                                                 \
    if (carveout >= 0) cudaFuncSetAttribute(kk, cudaFuncAttributePreferredSharedMemoryCarveout, carveout);           \
    chained_launch(kk, grid, dim3(kPullThreads), 0, stream, p, count, inv, pow2);                                    \
  } while (0)
#define PULL_MODE(CPL)                                       \
  do {                                                       \
    if (det) {                                               \
      if (accumulate) PULL_LAUNCH(CPL, true, true);          \
      else PULL_LAUNCH(CPL, true, false);                    \
    } else {                                                 \
      if (accumulate) PULL_LAUNCH(CPL, false, true);         \
      else PULL_LAUNCH(CPL, false, false);                   \
    }                                                        \
  } while (0)
  if (cpl == 4) PULL_MODE(4);
  else if (cpl == 2) PULL_MODE(2);
  else PULL_MODE(1);
#undef PULL_MODE
#undef PULL_LAUNCH
  count_launch();
  return cudaGetLastError();
}

}  // namespace mobula_b200
