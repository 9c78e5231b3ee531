// C-ABI of libmobula_roialign_sm100.so (see include/mobula_roialign.h).
//
// This file owns everything that is not a kernel: argument validation, the device
// guard (reference: mobula/cpp/include/context/hip_ctx.h:95-103), the per-device
// workspace, the host-buffer staging path, algorithm selection and the drop-in
// symbols the reference's ctypes caller binds (mobula/func.py:155-156).
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>

#include "../../include/mobula_roialign.h"
#include "roialign_kernels.h"
#include "roialign_common.cuh"

#define MOBULA_EXPORT extern "C" __attribute__((visibility("default")))

namespace mobula_b200 {

static std::atomic<uint64_t> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

static thread_local std::string t_error;
// last kernel family used, process wide (autograd runs the backward on another thread)
static std::atomic<const char*> t_algo[2] = {{"none"}, {"none"}};

static int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  t_error = buf;
  return code;
}

#define CUDA_TRY(expr)                                                                       \
  do {                                                                                       \
    cudaError_t e_ = (expr);                                                                 \
    if (e_ != cudaSuccess)                                                                   \
      return fail(MOBULA_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_),   \
                  __FILE__, __LINE__);                                                       \
  } while (0)

// Switches to `device` for the lifetime of the object and restores the previous one.
class DeviceGuard {
 public:
  explicit DeviceGuard(int device) : prev_(-1), status_(cudaSuccess) {
    status_ = cudaGetDevice(&prev_);
    if (status_ == cudaSuccess && device >= 0 && device != prev_) {
      status_ = cudaSetDevice(device);
      switched_ = status_ == cudaSuccess;
    }
  }
  ~DeviceGuard() {
    if (switched_) cudaSetDevice(prev_);
  }
  cudaError_t status() const { return status_; }
  int current(int requested) const { return requested >= 0 ? requested : prev_; }

 private:
  int prev_;
  cudaError_t status_;
  bool switched_ = false;
};

// ---- per-device state ------------------------------------------------------------
constexpr int kMaxDevices = 64;

struct Buffer {
  void* ptr = nullptr;
  size_t bytes = 0;
  // Grow-only; the old block is released on the stream so in-flight work finishes first.
  cudaError_t reserve(size_t need, cudaStream_t stream) {
    if (need <= bytes) return cudaSuccess;
    if (ptr) {
      cudaError_t e = cudaStreamSynchronize(stream);
      if (e != cudaSuccess) return e;
      e = cudaFree(ptr);
      if (e != cudaSuccess) return e;
      ptr = nullptr;
      bytes = 0;
    }
    const size_t rounded = (need + (size_t(1) << 20) - 1) & ~((size_t(1) << 20) - 1);
    cudaError_t e = cudaMalloc(&ptr, rounded);
    if (e == cudaSuccess) bytes = rounded;
    return e;
  }
  void release() {
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    bytes = 0;
  }
};

constexpr int kHostChunks = 32;   // at most; see host_chunks()

struct DeviceState {
  std::mutex mu;
  bool probed = false;
  int num_sms = 0;
  int smem_optin = 0;
  Buffer workspace;            // ROI tables of eager calls (one call at a time)
  // Workspaces donated to CUDA graphs: a capture bakes the workspace address into the graph, so
  // the buffer a capture used is never freed, grown or handed to an eager call again (it lives
  // until mobula_roialign_release); the eager path allocates a fresh one on its next call.
  Buffer graph_ws[8];
  int graph_ws_count = 0;
  cudaEvent_t ws_done = nullptr;   // recorded after the last kernel that reads the workspace
  cudaStream_t ws_stream = nullptr;
  bool ws_used = false;
  unsigned long long pull_entries_hint = 0;   // adaptive-grid pull backward: list entries the last call needed
  Buffer stage[4];             // host-buffer path: data/dy, rois, out/dx, spare
  // host-buffer path, pipelined over channel chunks: copy-in / copy-out streams and events
  cudaStream_t s_in = nullptr, s_out = nullptr;
  cudaEvent_t ev_start = nullptr, ev_in[kHostChunks] = {nullptr}, ev_done[kHostChunks] = {nullptr};
};
static DeviceState g_dev[kMaxDevices];

static int probe(DeviceState& st, int device) {
  if (st.probed) return MOBULA_OK;
  CUDA_TRY(cudaDeviceGetAttribute(&st.num_sms, cudaDevAttrMultiProcessorCount, device));
  CUDA_TRY(cudaDeviceGetAttribute(&st.smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
  st.probed = true;
  return MOBULA_OK;
}

static int check_common(const void* a, const void* b, const void* c, int num_rois, int channels,
                        int height, int width, int pooled_h, int pooled_w) {
  if (num_rois < 0 || channels < 0 || height <= 0 || width <= 0 || pooled_h <= 0 || pooled_w <= 0)
    return fail(MOBULA_ERR_INVALID_ARGUMENT,
                "bad shape: rois=%d channels=%d plane=%dx%d pooled=%dx%d", num_rois, channels,
                height, width, pooled_h, pooled_w);
  if ((long long)height * width > 0x7fffffffLL / 4)
    return fail(MOBULA_ERR_UNSUPPORTED, "feature-map plane %dx%d too large", height, width);
  if (num_rois > 0 && channels > 0 && (!a || !b || !c))
    return fail(MOBULA_ERR_INVALID_ARGUMENT, "null tensor pointer");
  return MOBULA_OK;
}

// Largest batch index + 1 over host-resident rois: the reference ABI never passes N
// (ROIAlign.cpp:9-16), the host-buffer path needs it to size its copies.
static int images_from_host_rois(const float* rois, int num_rois) {
  int n = 0;
  for (int r = 0; r < num_rois; ++r) {
    const int b = (int)rois[(size_t)r * 5];
    if (b + 1 > n) n = b + 1;
  }
  return n;
}

// MOBULA_ROIALIGN_PREFER_SWEEP=0 keeps the automatic choice on the resident kernels
static bool prefer_sweep() {
  static const bool v = [] { const char* e = getenv("MOBULA_ROIALIGN_PREFER_SWEEP"); return e ? atoi(e) != 0 : false; }();
  return v;
}

// MOBULA_ROIALIGN_PREFER_PULL=0 keeps the automatic choice of the backward on the resident kernels
// The pull kernels get their parallelism from the pixels (one warp per eight pixels and 128 channels): small
// planes with many channels (the reference benchmark's 14 x 14 x 1024) stay on the resident kernels.
static bool prefer_pull(const RoiAlignArgs& a) {
  static const bool v = [] { const char* e = getenv("MOBULA_ROIALIGN_PREFER_PULL"); return e ? atoi(e) != 0 : true; }();
  return v && (long long)a.num_images * a.height * a.width >= 32768;
}

static unsigned algo_of(unsigned flags) {
  return (flags & MOBULA_ROIALIGN_ALGO_MASK) >> MOBULA_ROIALIGN_ALGO_SHIFT;
}

// ---- device-pointer implementations --------------------------------------------------
static bool is_capturing(cudaStream_t stream) {
  cudaStreamCaptureStatus status = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(stream, &status) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return status == cudaStreamCaptureStatusActive;
}

// Hands a workspace to a call on `stream`.  Eager calls share one per-device buffer: a call on
// another stream waits for the previous user to finish with it, and it grows on demand.  While
// `stream` is being captured into a CUDA graph nothing may be allocated and eager events may
// not be waited on: the capture takes over ("is donated") the current eager buffer, which must
// already be large enough (run the same call once before capturing).  A donated buffer is never
// freed, grown or used by an eager call again, so replays cannot read freed or overwritten
// tables; ordering replays of DIFFERENT graphs against each other is the caller's business.
static int acquire_workspace(DeviceState& st, size_t bytes, cudaStream_t stream, void** out) {
  if (is_capturing(stream)) {
    Buffer* donated = st.graph_ws_count ? &st.graph_ws[st.graph_ws_count - 1] : nullptr;
    if (!donated || donated->bytes < bytes) {
      if (st.workspace.bytes < bytes || st.graph_ws_count >= 8)
        return fail(MOBULA_ERR_UNSUPPORTED,
                    "stream capture: the %zu-byte workspace of this call does not exist yet; run the "
                    "same call once outside the capture", bytes);
      st.graph_ws[st.graph_ws_count] = st.workspace;     // donate; eager calls get a new buffer
      donated = &st.graph_ws[st.graph_ws_count++];
      st.workspace = Buffer();
      st.ws_used = false;
    }
    *out = donated->ptr;
    return MOBULA_OK;
  }
  if (st.ws_used && st.ws_stream != stream) CUDA_TRY(cudaStreamWaitEvent(stream, st.ws_done, 0));
  if (bytes > st.workspace.bytes && st.ws_used) CUDA_TRY(cudaEventSynchronize(st.ws_done));
  CUDA_TRY(st.workspace.reserve(bytes, stream));
  *out = st.workspace.ptr;
  return MOBULA_OK;
}

// Builds the ROI tables for this call in the per-device workspace.
static int prepare_tables(DeviceState& st, const RoiAlignArgs& a, unsigned need, int pitch,
                          unsigned flags, cudaStream_t stream, RoiTables* tables) {
  roi_tables_set_arena_override((flags & MOBULA_ROIALIGN_DEBUG_NO_TABLES) ? 0 : -1);
  const size_t bytes = roi_tables_bytes(a, need);
  void* ws = nullptr;
  const int rc = acquire_workspace(st, bytes, stream, &ws);
  if (rc) {
    roi_tables_set_arena_override(-1);
    return rc;
  }
  *tables = roi_tables_carve(a, need, pitch, ws);
  roi_tables_set_arena_override(-1);
  CUDA_TRY(launch_roi_tables(a, *tables, need, stream));
  return MOBULA_OK;
}

static int release_tables(DeviceState& st, cudaStream_t stream) {
  if (is_capturing(stream)) return MOBULA_OK;   // an event recorded here could not be waited on later
  if (!st.ws_done) CUDA_TRY(cudaEventCreateWithFlags(&st.ws_done, cudaEventDisableTiming));
  CUDA_TRY(cudaEventRecord(st.ws_done, stream));
  st.ws_stream = stream;
  st.ws_used = true;
  return MOBULA_OK;
}

// Workspace of the resident kernels: same buffer, same hand-over rules as prepare_tables.
static int prepare_workspace(DeviceState& st, size_t bytes, cudaStream_t stream, void** ws) {
  return acquire_workspace(st, bytes, stream, ws);
}

static int forward_device(DeviceState& st, int device, cudaStream_t stream, const float* data,
                          const float* rois, float* out, const RoiAlignArgs& a, unsigned flags) {
  (void)device;
  (void)rois;
  const int accumulate = (flags & MOBULA_ROIALIGN_ACCUMULATE) ? 1 : 0;
  unsigned algo = algo_of(flags);
  if (algo > MOBULA_ROIALIGN_ALGO_PULL)
    return fail(MOBULA_ERR_INVALID_ARGUMENT, "unknown forward algorithm %u", algo);
  if (algo == MOBULA_ROIALIGN_ALGO_PULL) algo = MOBULA_ROIALIGN_ALGO_AUTO;     // a backward formulation
  const bool no_tables = (flags & MOBULA_ROIALIGN_DEBUG_NO_TABLES) != 0;
  // a half-precision or channel-last output: the sampling_ratio == 2 resident kernels (type only)
  // and the sweep kernels write it; nothing else does
  const bool typed = a.top_dtype != 0 || a.top_layout != 0;
  if (typed && accumulate)
    return fail(MOBULA_ERR_UNSUPPORTED, "MOBULA_ROIALIGN_ACCUMULATE needs an fp32 [R, C, PH, PW] output");
  ResidentPlan rplan;
  bool resident_ok = !no_tables && resident_supported(a, data, st.smem_optin, &rplan);
  if (typed && resident_ok && (!rplan.fwd_two || a.top_layout != 0)) resident_ok = false;
  if (typed && (algo == MOBULA_ROIALIGN_ALGO_GATHER || algo == MOBULA_ROIALIGN_ALGO_PLANE))
    return fail(MOBULA_ERR_UNSUPPORTED, "the gather / plane kernels write fp32 [R, C, PH, PW] only");
  {
    // sweep kernels: on request, and automatically wherever the resident kernels do not apply
    // (adaptive grids, planes beyond one SM's shared memory) instead of the global-memory gather
    SweepPlan splan;
    const bool want = algo == MOBULA_ROIALIGN_ALGO_SWEEP || a.aligned || (typed && !resident_ok) ||
                      (algo == MOBULA_ROIALIGN_ALGO_AUTO && (prefer_sweep() || !resident_ok));
    const bool sweep_ok = want && !no_tables && sweep_supported(a, data, st.smem_optin, st.num_sms, &splan);
    if ((algo == MOBULA_ROIALIGN_ALGO_SWEEP || a.aligned || (typed && !resident_ok)) && !sweep_ok)
      return fail(MOBULA_ERR_UNSUPPORTED,
                  "sweep algorithm (and MOBULA_ROIALIGN_ALIGNED, and fp16 / bf16 / channel-last outputs "
                  "outside the sampling_ratio 2 resident kernels) needs num_images in [1, 4096], pooled_w <= 16 and at least four "
                  "%d-pixel rows of 32 channels within %d bytes of shared memory", a.width, st.smem_optin);
    if (sweep_ok) {
      t_algo[0] = "sweep";
      void* ws = nullptr;
      int rc = prepare_workspace(st, sweep_fwd_workspace(a, splan), stream, &ws);
      if (rc) return rc;
      CUDA_TRY(launch_fwd_sweep(a, splan, ws, data, out, accumulate, st.num_sms, stream));
      return release_tables(st, stream);
    }
  }
  if (algo == MOBULA_ROIALIGN_ALGO_RESIDENT && !resident_ok)
    return fail(MOBULA_ERR_UNSUPPORTED,
                "resident algorithm needs num_images in [0, 4096], sampling_ratio in [1, 4], at most "
                "64 samples per ROI axis pair and a padded %dx%d fp32 plane within %d bytes of shared "
                "memory", a.height, a.width, st.smem_optin);
  if (resident_ok && (algo == MOBULA_ROIALIGN_ALGO_RESIDENT || algo == MOBULA_ROIALIGN_ALGO_AUTO)) {
    t_algo[0] = "resident";
    void* ws = nullptr;
    int rc = prepare_workspace(st, resident_fwd_workspace(a, rplan), stream, &ws);
    if (rc) return rc;
    CUDA_TRY(launch_fwd_resident(a, rplan, ws, data, out, accumulate, st.num_sms, stream));
    return release_tables(st, stream);
  }
  if (algo == MOBULA_ROIALIGN_ALGO_PLANE) {
    PlanePlan plan;
    if (!plane_supported(a, data, st.smem_optin, &plan))
      return fail(MOBULA_ERR_UNSUPPORTED,
                  "plane algorithm needs num_images in [0, 4096] and a %dx%d fp32 plane within %d "
                  "bytes of shared memory", a.height, a.width, st.smem_optin);
    t_algo[0] = "plane";
    RoiTables tables;
    int rc = prepare_tables(st, a, 0u, plan.pitch, flags, stream, &tables);
    if (rc) return rc;
    CUDA_TRY(launch_fwd_plane(a, plan, tables, data, out, accumulate, st.num_sms, stream));
    return release_tables(st, stream);
  }
  t_algo[0] = "gather";      // any shape, any sampling grid
  CUDA_TRY(launch_fwd_gather(a, data, out, accumulate, stream));
  return MOBULA_OK;
}

static int backward_device(DeviceState& st, int device, cudaStream_t stream, const float* dy,
                           const float* rois, float* dx, const RoiAlignArgs& a, int mode,
                           unsigned flags) {
  (void)device;
  (void)rois;
  const int accumulate = (flags & MOBULA_ROIALIGN_ACCUMULATE) ? 1 : 0;
  unsigned algo = algo_of(flags);
  if (mode != MOBULA_ROIALIGN_BWD_ATOMIC && mode != MOBULA_ROIALIGN_BWD_DETERMINISTIC)
    return fail(MOBULA_ERR_INVALID_ARGUMENT, "unknown backward mode %d", mode);
  if (algo > MOBULA_ROIALIGN_ALGO_PULL)
    return fail(MOBULA_ERR_INVALID_ARGUMENT, "unknown backward algorithm %u", algo);
  const bool want_tile = algo == MOBULA_ROIALIGN_ALGO_SWEEP || a.aligned;
  if (!accumulate && a.num_images < 0)
    return fail(MOBULA_ERR_INVALID_ARGUMENT,
                "backward without MOBULA_ROIALIGN_ACCUMULATE needs num_images >= 0");
  const bool deterministic = mode == MOBULA_ROIALIGN_BWD_DETERMINISTIC;
  // a half-precision or channel-last dy is read by the pull kernels only (their copy stage widens it)
  const bool typed = a.top_dtype != 0 || a.top_layout != 0;
  if (typed && algo != MOBULA_ROIALIGN_ALGO_AUTO && algo != MOBULA_ROIALIGN_ALGO_PULL)
    return fail(MOBULA_ERR_UNSUPPORTED, "an fp16 / bf16 / channel-last dy needs the pull backward");
  if (deterministic && algo == MOBULA_ROIALIGN_ALGO_GATHER)
    return fail(MOBULA_ERR_UNSUPPORTED, "the gather backward uses global atomics: not deterministic");
  auto run_tile = [&](const TilePlan& tplan) -> int {
    t_algo[1] = deterministic ? "tile-canonical" : "tile";
    void* ws = nullptr;
    int rc = prepare_workspace(st, tile_bwd_workspace(a, tplan), stream, &ws);
    if (rc) return rc;
    CUDA_TRY(launch_bwd_tile(a, tplan, ws, dy, dx, accumulate, deterministic, st.num_sms, stream));
    return release_tables(st, stream);
  };
  {
    // pull kernels (per-pixel contribution lists, lane = channel): fixed sampling ratios
    // (adaptive grids read one counter back: not inside a stream capture)
    const bool adaptive = a.sampling_ratio <= 0;
    const bool pull_ok = !(flags & MOBULA_ROIALIGN_DEBUG_NO_TABLES) && a.num_images >= 1 && pull_supported(a) &&
                         !(adaptive && is_capturing(stream));
    if ((algo == MOBULA_ROIALIGN_ALGO_PULL || typed) && !pull_ok)
      return fail(MOBULA_ERR_UNSUPPORTED,
                  "pull backward needs num_images >= 1, pooled sizes <= 64, pooled_h * pooled_w <= 1600 and either a "
                  "fixed sampling_ratio with at most 64 samples per ROI axis on a plane within 4095 x 4095, or an "
                  "adaptive grid on a plane within 2048 x 2048 outside stream capture");
    if (pull_ok && (algo == MOBULA_ROIALIGN_ALGO_PULL || typed || (algo == MOBULA_ROIALIGN_ALGO_AUTO && prefer_pull(a)))) {
      // adaptive grids: start from what the last call on this device needed (else 256 entries per ROI and bin)
      unsigned long long entries = 0;
      if (adaptive)
        entries = st.pull_entries_hint ? st.pull_entries_hint
                                       : (unsigned long long)a.num_rois * a.pooled_h * a.pooled_w * 256ull;
      bool done = false;
      for (int attempt = 0; attempt < 2 && !done; ++attempt) {
        void* ws = nullptr;
        int rc = prepare_workspace(st, pull_bwd_workspace(a, deterministic, entries), stream, &ws);
        if (rc) return rc;
        unsigned long long need = 0;
        CUDA_TRY(launch_bwd_pull(a, ws, entries, dy, dx, accumulate, deterministic, st.num_sms, stream, &need));
        if (need == 0) {
          done = true;
        } else if (need == ~0ull) {
          break;                                   // lists beyond 32 bits: the row-tile kernel below
        } else {
          entries = need + need / 8;               // the tables are rebuilt in the larger workspace
          st.pull_entries_hint = entries;
        }
      }
      if (done) {
        t_algo[1] = deterministic ? "pull-canonical" : "pull";
        return release_tables(st, stream);
      }
      if (algo == MOBULA_ROIALIGN_ALGO_PULL || typed)
        return fail(MOBULA_ERR_UNSUPPORTED, "pull backward: the contribution lists of this call exceed 2^32 entries");
      { const int rc = release_tables(st, stream); if (rc) return rc; }
    }
  }
  TilePlan tplan;
  const bool tile_ok = !(flags & MOBULA_ROIALIGN_DEBUG_NO_TABLES) && a.num_images >= 0 &&
                       tile_supported(a, st.smem_optin, st.num_sms, &tplan);
  if (want_tile) {
    if (!tile_ok)
      return fail(MOBULA_ERR_UNSUPPORTED,
                  "sweep (row tile) backward needs num_images in [1, 4096], pooled_w <= 16 and two %d-pixel "
                  "rows of 32 channels within %d bytes of shared memory", a.width, st.smem_optin);
    return run_tile(tplan);
  }
  {
    // resident kernels: fast (merged weights) or deterministic (canonical order) variant
    ResidentPlan rplan;
    const bool resident_ok = !(flags & MOBULA_ROIALIGN_DEBUG_NO_TABLES) && a.num_images >= 0 &&
                             resident_supported(a, dx, st.smem_optin, &rplan) && rplan.bwd_ok;
    if (algo == MOBULA_ROIALIGN_ALGO_RESIDENT && !resident_ok)
      return fail(MOBULA_ERR_UNSUPPORTED,
                  "resident backward needs num_images in [0, 4096], sampling_ratio 2, pooled_w <= 16 and "
                  "a %dx%d fp32 plane slab within %d bytes of shared memory", a.height, a.width,
                  st.smem_optin);
    if (resident_ok && (algo == MOBULA_ROIALIGN_ALGO_RESIDENT || algo == MOBULA_ROIALIGN_ALGO_AUTO)) {
      t_algo[1] = deterministic ? "resident-canonical" : "resident";
      void* ws = nullptr;
      int rc = prepare_workspace(st, resident_bwd_workspace(a, rplan, deterministic), stream, &ws);
      if (rc) return rc;
      CUDA_TRY(launch_bwd_resident(a, rplan, ws, dy, dx, accumulate, deterministic, st.num_sms, stream));
      return release_tables(st, stream);
    }
  }
  // any sampling grid, any plane that leaves room for two tile rows: the row-tile kernel
  if (tile_ok && algo == MOBULA_ROIALIGN_ALGO_AUTO) return run_tile(tplan);
  if (deterministic || algo == MOBULA_ROIALIGN_ALGO_PLANE) {
    PlanePlan plan;
    if (!plane_supported(a, dx, st.smem_optin, &plan))
      return fail(MOBULA_ERR_UNSUPPORTED,
                  "%s needs num_images in [0, 4096] and a %dx%d fp32 plane within %d bytes of shared "
                  "memory", deterministic ? "deterministic backward" : "plane algorithm", a.height,
                  a.width, st.smem_optin);
    // owner-computes plane kernel, contributions in the canonical (single-thread) order
    t_algo[1] = "plane-canonical";
    RoiTables tables;
    int rc = prepare_tables(st, a, kNeedCells, plan.bwd_pitch, flags, stream, &tables);
    if (rc) return rc;
    CUDA_TRY(launch_bwd_plane(a, plan, tables, dy, dx, accumulate, st.num_sms, stream));
    return release_tables(st, stream);
  }
  t_algo[1] = "gather";
  if (!accumulate) {
    const size_t bytes = (size_t)a.num_images * a.channels * a.height * a.width * sizeof(float);
    if (bytes) CUDA_TRY(cudaMemsetAsync(dx, 0, bytes, stream));
  }
  CUDA_TRY(launch_bwd_gather(a, dy, dx, stream));
  return MOBULA_OK;
}

static RoiAlignArgs make_args(const float* rois, int num_rois, int num_images, int channels,
                              int height, int width, int pooled_h, int pooled_w, float scale,
                              int sampling_ratio) {
  RoiAlignArgs a;
  a.rois = rois;
  a.num_rois = num_rois;
  a.num_images = num_images;
  a.channels = channels;
  a.c_count = 0;
  a.height = height;
  a.width = width;
  a.pooled_h = pooled_h;
  a.pooled_w = pooled_w;
  a.scale = scale;
  a.sampling_ratio = sampling_ratio;
  a.aligned = 0;
  a.top_dtype = 0;
  a.top_layout = 0;
  a.pre_img_start = nullptr;
  a.pre_order = nullptr;
  return a;
}

}  // namespace mobula_b200

using namespace mobula_b200;

// ---------------------------------------------------------------------------------------
// v2 entry points
// ---------------------------------------------------------------------------------------
MOBULA_EXPORT int mobula_roialign_abi_version(void) { return MOBULA_ROIALIGN_ABI_VERSION; }

MOBULA_EXPORT const char* mobula_last_error(void) { return t_error.c_str(); }

MOBULA_EXPORT uint64_t mobula_roialign_launch_count(void) {
  return g_launches.load(std::memory_order_relaxed);
}

MOBULA_EXPORT const char* mobula_roialign_last_algo(int backward) { return t_algo[backward ? 1 : 0].load(); }

// ---- host-buffer path, pipelined -------------------------------------------------------
// The staged tensors are cut into channel chunks; chunk k+1 is copied in and chunk k-1 copied
// out (two more streams, both directions of the PCIe link) while chunk k is computed.  Needs
// the resident kernels (they take a channel range); anything else runs unchunked.
static int pipeline_setup(DeviceState& st) {
  if (st.s_in) return MOBULA_OK;
  CUDA_TRY(cudaStreamCreateWithFlags(&st.s_in, cudaStreamNonBlocking));
  CUDA_TRY(cudaStreamCreateWithFlags(&st.s_out, cudaStreamNonBlocking));
  CUDA_TRY(cudaEventCreateWithFlags(&st.ev_start, cudaEventDisableTiming));
  for (int k = 0; k < kHostChunks; ++k) {
    CUDA_TRY(cudaEventCreateWithFlags(&st.ev_in[k], cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&st.ev_done[k], cudaEventDisableTiming));
  }
  return MOBULA_OK;
}

// Chunks of the host-buffer pipeline: enough for the copies to hide the kernels and for the
// exposed first / last chunk to be short, few enough for every copy and kernel to stay large.
// Measured on the C2 shapes (B200, PCIe 5 x16): the forward is best with 16 (its tail is one
// kernel and one output chunk), the backward with 8 (its kernel needs >= 16 channel pairs to fill
// the SMs, and its head is one kernel).  MOBULA_HOST_CHUNKS_FWD / _BWD override.
static int host_chunks(bool backward) {
  static const int chunks[2] = {
      [] { const char* e = getenv("MOBULA_HOST_CHUNKS_FWD"); return e ? atoi(e) : 16; }(),
      [] { const char* e = getenv("MOBULA_HOST_CHUNKS_BWD"); return e ? atoi(e) : 8; }()};
  const int v = chunks[backward ? 1 : 0];
  return v < 1 ? 1 : v > kHostChunks ? kHostChunks : v;
}

// channels per chunk (even, so that the backward's channel pairs stay whole); 0: do not chunk
static int host_chunk(int channels, size_t tensor_bytes, bool backward) {
  int chunks = host_chunks(backward);
  if (tensor_bytes < (size_t(16) << 20)) return 0;
  while (chunks > 1 && channels < 8 * chunks) chunks >>= 1;     // at least 8 channels per chunk
  if (chunks < 2) return 0;
  int cc = (channels + chunks - 1) / chunks;
  return (cc + 1) & ~1;
}

static bool resident_forward_applies(DeviceState& st, const RoiAlignArgs& a, const void* data, unsigned flags) {
  const unsigned algo = algo_of(flags);
  ResidentPlan plan;
  return !(flags & MOBULA_ROIALIGN_DEBUG_NO_TABLES) && !a.aligned &&
         (algo == MOBULA_ROIALIGN_ALGO_AUTO || algo == MOBULA_ROIALIGN_ALGO_RESIDENT) &&
         resident_supported(a, data, st.smem_optin, &plan);
}

static bool resident_backward_applies(DeviceState& st, const RoiAlignArgs& a, const void* dx, int mode,
                                      unsigned flags) {
  const unsigned algo = algo_of(flags);
  ResidentPlan plan;
  (void)mode;
  // (the pull kernels take a channel range too)
  if (!(flags & MOBULA_ROIALIGN_DEBUG_NO_TABLES) && a.num_images >= 1 && pull_supported(a) &&
      (algo == MOBULA_ROIALIGN_ALGO_PULL || (algo == MOBULA_ROIALIGN_ALGO_AUTO && prefer_pull(a))))
    return true;
  return !(flags & MOBULA_ROIALIGN_DEBUG_NO_TABLES) && !a.aligned &&
         (algo == MOBULA_ROIALIGN_ALGO_AUTO || algo == MOBULA_ROIALIGN_ALGO_RESIDENT) &&
         resident_supported(a, dx, st.smem_optin, &plan) && plan.bwd_ok;
}

static int check_top(int dtype, int layout) {
  if (dtype < MOBULA_DTYPE_F32 || dtype > MOBULA_DTYPE_BF16 || layout < MOBULA_LAYOUT_RCHW || layout > MOBULA_LAYOUT_RHWC)
    return fail(MOBULA_ERR_INVALID_ARGUMENT, "unknown element type %d / layout %d of the pooled tensor", dtype, layout);
  return MOBULA_OK;
}

// `top_data` is typed by `top_dtype`; the float* is only a carrier
static int forward_impl(int device_id, void* stream_, const float* bottom_data, const float* bottom_rois,
                        float* top_data, int top_dtype, int top_layout, int num_rois, int num_images,
                        int channels, int height, int width, int pooled_height, int pooled_width,
                        float spatial_scale, int sampling_ratio, unsigned flags) {
  int rc = check_top(top_dtype, top_layout);
  if (rc) return rc;
  const size_t esz = top_elem_bytes(top_dtype);
  rc = check_common(bottom_data, bottom_rois, top_data, num_rois, channels, height, width,
                        pooled_height, pooled_width);
  if (rc) return rc;
  if (num_rois == 0 || channels == 0) return MOBULA_OK;
  DeviceGuard guard(device_id);
  if (guard.status() != cudaSuccess)
    return fail(MOBULA_ERR_CUDA, "cannot select device %d: %s", device_id,
                cudaGetErrorString(guard.status()));
  const int device = guard.current(device_id);
  if (device < 0 || device >= kMaxDevices) return fail(MOBULA_ERR_INVALID_ARGUMENT, "device %d", device);
  DeviceState& st = g_dev[device];
  std::lock_guard<std::mutex> lock(st.mu);
  if ((rc = probe(st, device))) return rc;
  cudaStream_t stream = (cudaStream_t)stream_;

  if (!(flags & MOBULA_ROIALIGN_HOST_BUFFERS)) {
    RoiAlignArgs a = make_args(bottom_rois, num_rois, num_images, channels, height, width,
                               pooled_height, pooled_width, spatial_scale, sampling_ratio);
    a.aligned = (flags & MOBULA_ROIALIGN_ALIGNED) ? 1 : 0;
    a.top_dtype = top_dtype;
    a.top_layout = top_layout;
    return forward_device(st, device, stream, bottom_data, bottom_rois, top_data, a, flags);
  }

  // host buffers: stage in, run, stage out, synchronise
  const int n_img = num_images >= 0 ? num_images : images_from_host_rois(bottom_rois, num_rois);
  const size_t data_bytes = (size_t)n_img * channels * height * width * sizeof(float);
  const size_t rois_bytes = (size_t)num_rois * 5 * sizeof(float);
  const size_t out_bytes = (size_t)num_rois * channels * pooled_height * pooled_width * esz;
  CUDA_TRY(st.stage[0].reserve(data_bytes, stream));
  CUDA_TRY(st.stage[1].reserve(rois_bytes, stream));
  CUDA_TRY(st.stage[2].reserve(out_bytes, stream));
  float* d_data = (float*)st.stage[0].ptr;
  float* d_rois = (float*)st.stage[1].ptr;
  float* d_out = (float*)st.stage[2].ptr;
  RoiAlignArgs a = make_args(d_rois, num_rois, n_img, channels, height, width, pooled_height,
                             pooled_width, spatial_scale, sampling_ratio);
  a.aligned = (flags & MOBULA_ROIALIGN_ALIGNED) ? 1 : 0;
  a.top_dtype = top_dtype;
  a.top_layout = top_layout;
  // (channel chunks are runs of an [R, C, PH, PW] output; a channel-last one is staged whole)
  const int cc = (top_layout == 0 && resident_forward_applies(st, a, d_data, flags)) ? host_chunk(channels, data_bytes, false) : 0;
  if (cc > 0) {
    if ((rc = pipeline_setup(st))) return rc;
    const size_t hw = (size_t)height * width, bins = (size_t)pooled_height * pooled_width;
    const size_t row_pitch = (size_t)channels * bins * esz;
    char* const h_top = reinterpret_cast<char*>(top_data);
    char* const d_top = reinterpret_cast<char*>(d_out);
    const size_t img_pitch = (size_t)channels * hw * sizeof(float);
    CUDA_TRY(cudaMemcpyAsync(d_rois, bottom_rois, rois_bytes, cudaMemcpyHostToDevice, stream));
    CUDA_TRY(cudaEventRecord(st.ev_start, stream));
    CUDA_TRY(cudaStreamWaitEvent(st.s_in, st.ev_start, 0));
    CUDA_TRY(cudaStreamWaitEvent(st.s_out, st.ev_start, 0));
    int k = 0;
    for (int c0 = 0; c0 < channels; c0 += cc, ++k) {
      const int cn = channels - c0 < cc ? channels - c0 : cc;
      // the chunk's channels of every image: n_img runs of cn planes, one image apart
      if (n_img > 0)
        CUDA_TRY(cudaMemcpy2DAsync(d_data + c0 * hw, img_pitch, bottom_data + c0 * hw, img_pitch,
                                   (size_t)cn * hw * sizeof(float), n_img, cudaMemcpyHostToDevice, st.s_in));
      if ((flags & MOBULA_ROIALIGN_ACCUMULATE) && num_rois > 0)
        CUDA_TRY(cudaMemcpy2DAsync(d_top + c0 * bins * esz, row_pitch, h_top + c0 * bins * esz, row_pitch,
                                   (size_t)cn * bins * esz, num_rois, cudaMemcpyHostToDevice, st.s_in));
      CUDA_TRY(cudaEventRecord(st.ev_in[k], st.s_in));
      CUDA_TRY(cudaStreamWaitEvent(stream, st.ev_in[k], 0));
      a.c_count = cn;
      if ((rc = forward_device(st, device, stream, d_data + c0 * hw, d_rois,
                               reinterpret_cast<float*>(d_top + c0 * bins * esz), a, flags)))
        return rc;
      CUDA_TRY(cudaEventRecord(st.ev_done[k], stream));
      CUDA_TRY(cudaStreamWaitEvent(st.s_out, st.ev_done[k], 0));
      if (num_rois > 0)
        CUDA_TRY(cudaMemcpy2DAsync(h_top + c0 * bins * esz, row_pitch, d_top + c0 * bins * esz, row_pitch,
                                   (size_t)cn * bins * esz, num_rois, cudaMemcpyDeviceToHost, st.s_out));
    }
    CUDA_TRY(cudaStreamSynchronize(st.s_out));
    CUDA_TRY(cudaStreamSynchronize(stream));
    return MOBULA_OK;
  }
  if (data_bytes) CUDA_TRY(cudaMemcpyAsync(d_data, bottom_data, data_bytes, cudaMemcpyHostToDevice, stream));
  CUDA_TRY(cudaMemcpyAsync(d_rois, bottom_rois, rois_bytes, cudaMemcpyHostToDevice, stream));
  if (flags & MOBULA_ROIALIGN_ACCUMULATE)
    CUDA_TRY(cudaMemcpyAsync(d_out, top_data, out_bytes, cudaMemcpyHostToDevice, stream));
  if ((rc = forward_device(st, device, stream, d_data, d_rois, d_out, a, flags))) return rc;
  CUDA_TRY(cudaMemcpyAsync(top_data, d_out, out_bytes, cudaMemcpyDeviceToHost, stream));
  CUDA_TRY(cudaStreamSynchronize(stream));
  return MOBULA_OK;
}

MOBULA_EXPORT int mobula_roialign_forward_f32(int device_id, void* stream, const float* bottom_data,
                                              const float* bottom_rois, float* top_data,
                                              int num_rois, int num_images, int channels, int height,
                                              int width, int pooled_height, int pooled_width,
                                              float spatial_scale, int sampling_ratio,
                                              unsigned flags) {
  return forward_impl(device_id, stream, bottom_data, bottom_rois, top_data, MOBULA_DTYPE_F32, MOBULA_LAYOUT_RCHW,
                      num_rois, num_images, channels, height, width, pooled_height, pooled_width, spatial_scale,
                      sampling_ratio, flags);
}

MOBULA_EXPORT int mobula_roialign_forward_ex(int device_id, void* stream, const float* bottom_data,
                                             const float* bottom_rois, void* top_data, int top_dtype,
                                             int top_layout, int num_rois, int num_images, int channels,
                                             int height, int width, int pooled_height, int pooled_width,
                                             float spatial_scale, int sampling_ratio, unsigned flags) {
  return forward_impl(device_id, stream, bottom_data, bottom_rois, static_cast<float*>(top_data), top_dtype,
                      top_layout, num_rois, num_images, channels, height, width, pooled_height, pooled_width,
                      spatial_scale, sampling_ratio, flags);
}

static int backward_impl(int device_id, void* stream_, const float* top_diff, int top_dtype, int top_layout,
                         const float* bottom_rois, float* bottom_diff, int num_rois, int num_images, int channels,
                         int height, int width, int pooled_height, int pooled_width, float spatial_scale,
                         int sampling_ratio, int mode, unsigned flags) {
  int rc = check_top(top_dtype, top_layout);
  if (rc) return rc;
  const size_t esz = top_elem_bytes(top_dtype);
  rc = check_common(top_diff, bottom_rois, bottom_diff, num_rois, channels, height, width,
                        pooled_height, pooled_width);
  if (rc) return rc;
  const bool accumulate = flags & MOBULA_ROIALIGN_ACCUMULATE;
  if (channels == 0) return MOBULA_OK;
  if (num_rois == 0 && (accumulate || num_images <= 0)) return MOBULA_OK;
  if (!bottom_diff) return fail(MOBULA_ERR_INVALID_ARGUMENT, "null tensor pointer");
  DeviceGuard guard(device_id);
  if (guard.status() != cudaSuccess)
    return fail(MOBULA_ERR_CUDA, "cannot select device %d: %s", device_id,
                cudaGetErrorString(guard.status()));
  const int device = guard.current(device_id);
  if (device < 0 || device >= kMaxDevices) return fail(MOBULA_ERR_INVALID_ARGUMENT, "device %d", device);
  DeviceState& st = g_dev[device];
  std::lock_guard<std::mutex> lock(st.mu);
  if ((rc = probe(st, device))) return rc;
  cudaStream_t stream = (cudaStream_t)stream_;

  if (!(flags & MOBULA_ROIALIGN_HOST_BUFFERS)) {
    RoiAlignArgs a = make_args(bottom_rois, num_rois, num_images, channels, height, width,
                               pooled_height, pooled_width, spatial_scale, sampling_ratio);
    a.aligned = (flags & MOBULA_ROIALIGN_ALIGNED) ? 1 : 0;
    a.top_dtype = top_dtype;
    a.top_layout = top_layout;
    return backward_device(st, device, stream, top_diff, bottom_rois, bottom_diff, a, mode, flags);
  }

  const int n_img = num_images >= 0 ? num_images : images_from_host_rois(bottom_rois, num_rois);
  const size_t dx_bytes = (size_t)n_img * channels * height * width * sizeof(float);
  const size_t rois_bytes = (size_t)num_rois * 5 * sizeof(float);
  const size_t dy_bytes = (size_t)num_rois * channels * pooled_height * pooled_width * esz;
  CUDA_TRY(st.stage[0].reserve(dy_bytes, stream));
  CUDA_TRY(st.stage[1].reserve(rois_bytes, stream));
  CUDA_TRY(st.stage[2].reserve(dx_bytes, stream));
  float* d_dy = (float*)st.stage[0].ptr;
  float* d_rois = (float*)st.stage[1].ptr;
  float* d_dx = (float*)st.stage[2].ptr;
  RoiAlignArgs a = make_args(d_rois, num_rois, n_img, channels, height, width, pooled_height,
                             pooled_width, spatial_scale, sampling_ratio);
  a.aligned = (flags & MOBULA_ROIALIGN_ALIGNED) ? 1 : 0;
  a.top_dtype = top_dtype;
  a.top_layout = top_layout;
  const int cc = (top_layout == 0 && resident_backward_applies(st, a, d_dx, mode, flags)) ? host_chunk(channels, dx_bytes, true) : 0;
  if (cc > 0) {
    if ((rc = pipeline_setup(st))) return rc;
    const size_t hw = (size_t)height * width, bins = (size_t)pooled_height * pooled_width;
    const size_t row_pitch = (size_t)channels * bins * esz;
    const char* const h_top = reinterpret_cast<const char*>(top_diff);
    char* const d_top = reinterpret_cast<char*>(d_dy);
    const size_t img_pitch = (size_t)channels * hw * sizeof(float);
    if (rois_bytes) CUDA_TRY(cudaMemcpyAsync(d_rois, bottom_rois, rois_bytes, cudaMemcpyHostToDevice, stream));
    CUDA_TRY(cudaEventRecord(st.ev_start, stream));
    CUDA_TRY(cudaStreamWaitEvent(st.s_in, st.ev_start, 0));
    CUDA_TRY(cudaStreamWaitEvent(st.s_out, st.ev_start, 0));
    int k = 0;
    for (int c0 = 0; c0 < channels; c0 += cc, ++k) {
      const int cn = channels - c0 < cc ? channels - c0 : cc;
      if (num_rois > 0)
        CUDA_TRY(cudaMemcpy2DAsync(d_top + c0 * bins * esz, row_pitch, h_top + c0 * bins * esz, row_pitch,
                                   (size_t)cn * bins * esz, num_rois, cudaMemcpyHostToDevice, st.s_in));
      if (accumulate && n_img > 0)
        CUDA_TRY(cudaMemcpy2DAsync(d_dx + c0 * hw, img_pitch, bottom_diff + c0 * hw, img_pitch,
                                   (size_t)cn * hw * sizeof(float), n_img, cudaMemcpyHostToDevice, st.s_in));
      CUDA_TRY(cudaEventRecord(st.ev_in[k], st.s_in));
      CUDA_TRY(cudaStreamWaitEvent(stream, st.ev_in[k], 0));
      a.c_count = cn;
      if ((rc = backward_device(st, device, stream, reinterpret_cast<const float*>(d_top + c0 * bins * esz), d_rois,
                                d_dx + c0 * hw, a, mode, flags)))
        return rc;
      CUDA_TRY(cudaEventRecord(st.ev_done[k], stream));
      CUDA_TRY(cudaStreamWaitEvent(st.s_out, st.ev_done[k], 0));
      if (n_img > 0)
        CUDA_TRY(cudaMemcpy2DAsync(bottom_diff + c0 * hw, img_pitch, d_dx + c0 * hw, img_pitch,
                                   (size_t)cn * hw * sizeof(float), n_img, cudaMemcpyDeviceToHost, st.s_out));
    }
    CUDA_TRY(cudaStreamSynchronize(st.s_out));
    CUDA_TRY(cudaStreamSynchronize(stream));
    return MOBULA_OK;
  }
  if (dy_bytes) CUDA_TRY(cudaMemcpyAsync(d_dy, top_diff, dy_bytes, cudaMemcpyHostToDevice, stream));
  if (rois_bytes) CUDA_TRY(cudaMemcpyAsync(d_rois, bottom_rois, rois_bytes, cudaMemcpyHostToDevice, stream));
  if (accumulate && dx_bytes)
    CUDA_TRY(cudaMemcpyAsync(d_dx, bottom_diff, dx_bytes, cudaMemcpyHostToDevice, stream));
  if ((rc = backward_device(st, device, stream, d_dy, d_rois, d_dx, a, mode, flags))) return rc;
  if (dx_bytes) CUDA_TRY(cudaMemcpyAsync(bottom_diff, d_dx, dx_bytes, cudaMemcpyDeviceToHost, stream));
  CUDA_TRY(cudaStreamSynchronize(stream));
  return MOBULA_OK;
}

MOBULA_EXPORT int mobula_roialign_backward_f32(int device_id, void* stream, const float* top_diff,
                                               const float* bottom_rois, float* bottom_diff,
                                               int num_rois, int num_images, int channels,
                                               int height, int width, int pooled_height,
                                               int pooled_width, float spatial_scale,
                                               int sampling_ratio, int mode, unsigned flags) {
  return backward_impl(device_id, stream, top_diff, MOBULA_DTYPE_F32, MOBULA_LAYOUT_RCHW, bottom_rois, bottom_diff,
                       num_rois, num_images, channels, height, width, pooled_height, pooled_width, spatial_scale,
                       sampling_ratio, mode, flags);
}

MOBULA_EXPORT int mobula_roialign_backward_ex(int device_id, void* stream, const void* top_diff, int top_dtype,
                                              int top_layout, const float* bottom_rois, float* bottom_diff,
                                              int num_rois, int num_images, int channels, int height, int width,
                                              int pooled_height, int pooled_width, float spatial_scale,
                                              int sampling_ratio, int mode, unsigned flags) {
  return backward_impl(device_id, stream, static_cast<const float*>(top_diff), top_dtype, top_layout, bottom_rois,
                       bottom_diff, num_rois, num_images, channels, height, width, pooled_height, pooled_width,
                       spatial_scale, sampling_ratio, mode, flags);
}

// ---- prepared ROI tables -------------------------------------------------------------------
// What depends on the ROI table alone -- its validation, a private copy and the stable partition
// of its rows by image -- done once and shared by the forward and the backward of a step.
struct RoiPlan {
  unsigned magic;
  int device;
  RoiAlignArgs a;        // rois -> the plan's own copy
  float* rois;           // [R, 5]
  int* img_start;        // [N + 2]
  int* order;            // [R]
};
constexpr unsigned kPlanMagic = 0x524f4950u;   // "ROIP"

static RoiPlan* plan_of(void* handle) {
  RoiPlan* p = static_cast<RoiPlan*>(handle);
  return (p && p->magic == kPlanMagic) ? p : nullptr;
}

MOBULA_EXPORT int mobula_roialign_prepare_f32(int device_id, void* stream_, const float* bottom_rois, int num_rois,
                                              int num_images, int height, int width, int pooled_height,
                                              int pooled_width, float spatial_scale, int sampling_ratio,
                                              unsigned flags, void** plan_out) {
  if (!plan_out) return fail(MOBULA_ERR_INVALID_ARGUMENT, "null plan pointer");
  *plan_out = nullptr;
  if (flags & MOBULA_ROIALIGN_HOST_BUFFERS)
    return fail(MOBULA_ERR_UNSUPPORTED, "prepared ROI tables take device pointers");
  if (num_images < 0) return fail(MOBULA_ERR_INVALID_ARGUMENT, "prepare needs num_images >= 0");
  int rc = check_common(bottom_rois, bottom_rois, bottom_rois, num_rois, 1, height, width, pooled_height, pooled_width);
  if (rc) return rc;
  DeviceGuard guard(device_id);
  if (guard.status() != cudaSuccess)
    return fail(MOBULA_ERR_CUDA, "cannot select device %d: %s", device_id, cudaGetErrorString(guard.status()));
  const int device = guard.current(device_id);
  if (device < 0 || device >= kMaxDevices) return fail(MOBULA_ERR_INVALID_ARGUMENT, "device %d", device);
  cudaStream_t stream = (cudaStream_t)stream_;
  RoiPlan* p = new RoiPlan();
  p->magic = kPlanMagic;
  p->device = device;
  p->rois = nullptr;
  p->img_start = p->order = nullptr;
  // one allocation: rois copy | img_start | order
  const size_t rois_bytes = ((size_t)num_rois * 5 * sizeof(float) + 255) & ~(size_t)255;
  const size_t start_bytes = ((size_t)(num_images + 2) * sizeof(int) + 255) & ~(size_t)255;
  const size_t order_bytes = ((size_t)num_rois * sizeof(int) + 255) & ~(size_t)255;
  char* base = nullptr;
  cudaError_t e = cudaMalloc(&base, rois_bytes + start_bytes + order_bytes + 256);
  if (e != cudaSuccess) {
    delete p;
    return fail(MOBULA_ERR_CUDA, "cudaMalloc: %s", cudaGetErrorString(e));
  }
  p->rois = reinterpret_cast<float*>(base);
  p->img_start = reinterpret_cast<int*>(base + rois_bytes);
  p->order = reinterpret_cast<int*>(base + rois_bytes + start_bytes);
  p->a = make_args(p->rois, num_rois, num_images, 0, height, width, pooled_height, pooled_width, spatial_scale,
                   sampling_ratio);
  p->a.aligned = (flags & MOBULA_ROIALIGN_ALIGNED) ? 1 : 0;
  if (num_rois > 0)
    e = cudaMemcpyAsync(p->rois, bottom_rois, (size_t)num_rois * 5 * sizeof(float), cudaMemcpyDeviceToDevice, stream);
  if (e == cudaSuccess) e = launch_roi_group(p->a, p->img_start, p->order, stream);
  if (e != cudaSuccess) {
    cudaFree(base);
    delete p;
    return fail(MOBULA_ERR_CUDA, "prepare: %s", cudaGetErrorString(e));
  }
  p->a.pre_img_start = p->img_start;
  p->a.pre_order = p->order;
  *plan_out = p;
  return MOBULA_OK;
}

MOBULA_EXPORT int mobula_roialign_forward_prepared_f32(void* plan, void* stream_, const float* bottom_data,
                                                       float* top_data, int channels, unsigned flags) {
  RoiPlan* p = plan_of(plan);
  if (!p) return fail(MOBULA_ERR_INVALID_ARGUMENT, "not a plan of mobula_roialign_prepare_f32");
  if (flags & MOBULA_ROIALIGN_HOST_BUFFERS) return fail(MOBULA_ERR_UNSUPPORTED, "prepared calls take device pointers");
  RoiAlignArgs a = p->a;
  a.channels = channels;
  int rc = check_common(bottom_data, a.rois, top_data, a.num_rois, channels, a.height, a.width, a.pooled_h, a.pooled_w);
  if (rc) return rc;
  if (a.num_rois == 0 || channels == 0) return MOBULA_OK;
  DeviceGuard guard(p->device);
  if (guard.status() != cudaSuccess) return fail(MOBULA_ERR_CUDA, "cannot select device %d", p->device);
  DeviceState& st = g_dev[p->device];
  std::lock_guard<std::mutex> lock(st.mu);
  if ((rc = probe(st, p->device))) return rc;
  return forward_device(st, p->device, (cudaStream_t)stream_, bottom_data, a.rois, top_data, a, flags);
}

MOBULA_EXPORT int mobula_roialign_backward_prepared_f32(void* plan, void* stream_, const float* top_diff,
                                                        float* bottom_diff, int channels, int mode, unsigned flags) {
  RoiPlan* p = plan_of(plan);
  if (!p) return fail(MOBULA_ERR_INVALID_ARGUMENT, "not a plan of mobula_roialign_prepare_f32");
  if (flags & MOBULA_ROIALIGN_HOST_BUFFERS) return fail(MOBULA_ERR_UNSUPPORTED, "prepared calls take device pointers");
  RoiAlignArgs a = p->a;
  a.channels = channels;
  int rc = check_common(top_diff, a.rois, bottom_diff, a.num_rois, channels, a.height, a.width, a.pooled_h, a.pooled_w);
  if (rc) return rc;
  if (channels == 0) return MOBULA_OK;
  if (a.num_rois == 0 && ((flags & MOBULA_ROIALIGN_ACCUMULATE) || a.num_images <= 0)) return MOBULA_OK;
  if (!bottom_diff) return fail(MOBULA_ERR_INVALID_ARGUMENT, "null tensor pointer");
  DeviceGuard guard(p->device);
  if (guard.status() != cudaSuccess) return fail(MOBULA_ERR_CUDA, "cannot select device %d", p->device);
  DeviceState& st = g_dev[p->device];
  std::lock_guard<std::mutex> lock(st.mu);
  if ((rc = probe(st, p->device))) return rc;
  return backward_device(st, p->device, (cudaStream_t)stream_, top_diff, a.rois, bottom_diff, a, mode, flags);
}

MOBULA_EXPORT void mobula_roialign_plan_destroy(void* plan) {
  RoiPlan* p = plan_of(plan);
  if (!p) return;
  DeviceGuard guard(p->device);
  cudaFree(p->rois);       // (synchronises with the device: in-flight users finish first)
  p->magic = 0;
  delete p;
}

MOBULA_EXPORT int mobula_roialign_bookkeeping_f32(int device_id, void* stream_, const float* bottom_rois,
                                                  int num_rois, int height, int width,
                                                  int pooled_height, int pooled_width,
                                                  float spatial_scale, int sampling_ratio,
                                                  int32_t* geom_i, float* geom_f, int max_grid,
                                                  int32_t* tap_i, float* tap_f) {
  if (num_rois < 0 || height <= 0 || width <= 0 || pooled_height <= 0 || pooled_width <= 0)
    return fail(MOBULA_ERR_INVALID_ARGUMENT, "bad shape");
  if (num_rois == 0) return MOBULA_OK;
  if (!bottom_rois || !geom_i || !geom_f || (max_grid > 0 && (!tap_i || !tap_f)))
    return fail(MOBULA_ERR_INVALID_ARGUMENT, "null pointer");
  DeviceGuard guard(device_id);
  if (guard.status() != cudaSuccess)
    return fail(MOBULA_ERR_CUDA, "cannot select device %d: %s", device_id,
                cudaGetErrorString(guard.status()));
  const RoiAlignArgs a = make_args(bottom_rois, num_rois, -1, 0, height, width, pooled_height,
                                   pooled_width, spatial_scale, sampling_ratio);
  CUDA_TRY(launch_bookkeeping(a, geom_i, geom_f, max_grid, tap_i, tap_f, (cudaStream_t)stream_));
  return MOBULA_OK;
}

MOBULA_EXPORT void* mobula_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  return p;
}

MOBULA_EXPORT void mobula_host_free(void* ptr) {
  if (ptr) cudaFreeHost(ptr);
}

MOBULA_EXPORT void mobula_roialign_release(void) {
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess) {
    cudaGetLastError();
    return;
  }
  int prev = 0;
  cudaGetDevice(&prev);
  for (int d = 0; d < count && d < kMaxDevices; ++d) {
    DeviceState& st = g_dev[d];
    std::lock_guard<std::mutex> lock(st.mu);
    if (!st.workspace.ptr && !st.graph_ws_count && !st.stage[0].ptr && !st.stage[1].ptr &&
        !st.stage[2].ptr && !st.stage[3].ptr)
      continue;
    cudaSetDevice(d);
    cudaDeviceSynchronize();
    st.workspace.release();
    for (int k = 0; k < st.graph_ws_count; ++k) st.graph_ws[k].release();
    st.graph_ws_count = 0;
    if (st.ws_done) cudaEventDestroy(st.ws_done);
    st.ws_done = nullptr;
    st.ws_used = false;
    for (Buffer& b : st.stage) b.release();
    if (st.s_in) {
      cudaStreamDestroy(st.s_in);
      cudaStreamDestroy(st.s_out);
      cudaEventDestroy(st.ev_start);
      for (int k = 0; k < kHostChunks; ++k) {
        cudaEventDestroy(st.ev_in[k]);
        cudaEventDestroy(st.ev_done[k]);
      }
      st.s_in = st.s_out = nullptr;
    }
  }
  cudaSetDevice(prev);
}

// ---------------------------------------------------------------------------------------
// drop-in symbols of the reference ABI
// ---------------------------------------------------------------------------------------
// The reference aborts the process on a native error (logging.h:23-26, hip_ctx.h:49-60).
static void die_like_reference(const char* where, int rc) {
  fprintf(stderr, "[mobula_roialign] %s failed (%d): %s\n", where, rc, mobula_last_error());
  fflush(stderr);
  exit(-1);
}

static bool host_context(int device_id) { return device_id < 0; }

MOBULA_EXPORT void roi_align_forward_ab78de3c(const int device_id, const int nthreads,
                                              const float* bottom_data, const float spatial_scale,
                                              const int channels, const int height, const int width,
                                              const int pooled_height, const int pooled_width,
                                              const int sampling_ratio, const float* bottom_rois,
                                              float* top_data) {
  const long long per_roi = (long long)channels * pooled_height * pooled_width;
  if (nthreads <= 0 || per_roi <= 0) return;
  const int num_rois = (int)(nthreads / per_roi);   // R is implicit in the reference ABI
  unsigned flags = host_context(device_id) ? MOBULA_ROIALIGN_HOST_BUFFERS : 0u;
  int rc = mobula_roialign_forward_f32(device_id, nullptr, bottom_data, bottom_rois, top_data,
                                       num_rois, -1, channels, height, width, pooled_height,
                                       pooled_width, spatial_scale, sampling_ratio, flags);
  if (rc == MOBULA_OK && !host_context(device_id)) {
    DeviceGuard guard(device_id);
    if (cudaStreamSynchronize(nullptr) != cudaSuccess) {
      fail(MOBULA_ERR_CUDA, "synchronize: %s", cudaGetErrorString(cudaGetLastError()));
      rc = MOBULA_ERR_CUDA;
    }
  }
  if (rc != MOBULA_OK) die_like_reference("roi_align_forward", rc);
}

MOBULA_EXPORT void roi_align_backward_f318a24e(const int device_id, const int nthreads,
                                               const float* top_diff, const float spatial_scale,
                                               const int channels, const int height, const int width,
                                               const int pooled_height, const int pooled_width,
                                               const int sampling_ratio, float* bottom_diff,
                                               const float* bottom_rois) {
  const long long per_roi = (long long)channels * pooled_height * pooled_width;
  if (nthreads <= 0 || per_roi <= 0) return;
  const int num_rois = (int)(nthreads / per_roi);
  // The reference kernel adds into a buffer its caller zeroed (ROIAlign.py:33-34).
  unsigned flags = MOBULA_ROIALIGN_ACCUMULATE;
  if (host_context(device_id)) flags |= MOBULA_ROIALIGN_HOST_BUFFERS;
  int rc = mobula_roialign_backward_f32(device_id, nullptr, top_diff, bottom_rois, bottom_diff,
                                        num_rois, -1, channels, height, width, pooled_height,
                                        pooled_width, spatial_scale, sampling_ratio,
                                        MOBULA_ROIALIGN_BWD_ATOMIC, flags);
  if (rc == MOBULA_OK && !host_context(device_id)) {
    DeviceGuard guard(device_id);
    if (cudaStreamSynchronize(nullptr) != cudaSuccess) {
      fail(MOBULA_ERR_CUDA, "synchronize: %s", cudaGetErrorString(cudaGetLastError()));
      rc = MOBULA_ERR_CUDA;
    }
  }
  if (rc != MOBULA_OK) die_like_reference("roi_align_backward", rc);
}

// ---------------------------------------------------------------------------------------
// im2col / col2im (opzoo Convolution)
// ---------------------------------------------------------------------------------------
static int conv_call(bool to_col, int device_id, void* stream_, const float* src, int channels, int height,
                     int width, int kernel_h, int kernel_w, int pad_h, int pad_w, int stride_h, int stride_w,
                     int dilation_h, int dilation_w, int height_col, int width_col, float* dst, unsigned flags) {
  if (channels < 0 || height <= 0 || width <= 0 || kernel_h <= 0 || kernel_w <= 0 || pad_h < 0 || pad_w < 0 ||
      stride_h <= 0 || stride_w <= 0 || dilation_h <= 0 || dilation_w <= 0 || height_col < 0 || width_col < 0)
    return fail(MOBULA_ERR_INVALID_ARGUMENT, "bad convolution geometry");
  const long long im_elems = (long long)channels * height * width;
  const long long col_elems = (long long)channels * kernel_h * kernel_w * height_col * width_col;
  if (im_elems > 0x7fffffffLL || col_elems > 0x7fffffffLL)      // the reference indexes with int n
    return fail(MOBULA_ERR_UNSUPPORTED, "im2col / col2im: more than 2^31 - 1 elements");
  const long long n_src = to_col ? im_elems : col_elems, n_dst = to_col ? col_elems : im_elems;
  if (n_dst == 0) return MOBULA_OK;
  if (!dst || (n_src > 0 && !src)) return fail(MOBULA_ERR_INVALID_ARGUMENT, "null tensor pointer");
  DeviceGuard guard(device_id);
  if (guard.status() != cudaSuccess)
    return fail(MOBULA_ERR_CUDA, "cannot select device %d: %s", device_id, cudaGetErrorString(guard.status()));
  const int device = guard.current(device_id);
  if (device < 0 || device >= kMaxDevices) return fail(MOBULA_ERR_INVALID_ARGUMENT, "device %d", device);
  DeviceState& st = g_dev[device];
  std::lock_guard<std::mutex> lock(st.mu);
  int rc = probe(st, device);
  if (rc) return rc;
  cudaStream_t stream = (cudaStream_t)stream_;
  const float* d_src = src;
  float* d_dst = dst;
  const bool host = (flags & MOBULA_ROIALIGN_HOST_BUFFERS) != 0;
  if (host) {
    CUDA_TRY(st.stage[0].reserve((size_t)(n_src > 0 ? n_src : 1) * sizeof(float), stream));
    CUDA_TRY(st.stage[2].reserve((size_t)n_dst * sizeof(float), stream));
    d_src = (const float*)st.stage[0].ptr;
    d_dst = (float*)st.stage[2].ptr;
    if (n_src > 0)
      CUDA_TRY(cudaMemcpyAsync(st.stage[0].ptr, src, (size_t)n_src * sizeof(float), cudaMemcpyHostToDevice, stream));
  }
  if (to_col)
    CUDA_TRY(launch_im2col(d_src, channels, height, width, kernel_h, kernel_w, pad_h, pad_w, stride_h, stride_w,
                           dilation_h, dilation_w, height_col, width_col, d_dst, st.num_sms, stream));
  else
    CUDA_TRY(launch_col2im(d_src, channels, height, width, kernel_h, kernel_w, pad_h, pad_w, stride_h, stride_w,
                           dilation_h, dilation_w, height_col, width_col, d_dst, st.num_sms, stream));
  if (host) {
    CUDA_TRY(cudaMemcpyAsync(dst, d_dst, (size_t)n_dst * sizeof(float), cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
  }
  return MOBULA_OK;
}

MOBULA_EXPORT int mobula_im2col_f32(int device_id, void* stream, const float* data_im, int channels, int height,
                                    int width, int kernel_h, int kernel_w, int pad_h, int pad_w, int stride_h,
                                    int stride_w, int dilation_h, int dilation_w, int height_col, int width_col,
                                    float* data_col, unsigned flags) {
  return conv_call(true, device_id, stream, data_im, channels, height, width, kernel_h, kernel_w, pad_h, pad_w,
                   stride_h, stride_w, dilation_h, dilation_w, height_col, width_col, data_col, flags);
}

MOBULA_EXPORT int mobula_col2im_f32(int device_id, void* stream, const float* data_col, int channels, int height,
                                    int width, int kernel_h, int kernel_w, int pad_h, int pad_w, int stride_h,
                                    int stride_w, int dilation_h, int dilation_w, int height_col, int width_col,
                                    float* data_im, unsigned flags) {
  return conv_call(false, device_id, stream, data_col, channels, height, width, kernel_h, kernel_w, pad_h, pad_w,
                   stride_h, stride_w, dilation_h, dilation_w, height_col, width_col, data_im, flags);
}

// the reference ABI passes the element count n of the OUTPUT and no channel count
static void conv_dropin(bool to_col, const char* where, int device_id, int n, const float* src, int height,
                        int width, int kernel_h, int kernel_w, int pad_h, int pad_w, int stride_h, int stride_w,
                        int dilation_h, int dilation_w, int height_col, int width_col, float* dst) {
  if (n <= 0) return;
  const long long per_channel = to_col ? (long long)kernel_h * kernel_w * height_col * width_col
                                       : (long long)height * width;
  int rc = MOBULA_OK;
  if (per_channel <= 0 || n % per_channel != 0) {
    rc = fail(MOBULA_ERR_INVALID_ARGUMENT, "n = %d is not a whole number of channels", n);
  } else {
    const unsigned flags = host_context(device_id) ? MOBULA_ROIALIGN_HOST_BUFFERS : 0u;
    rc = conv_call(to_col, device_id, nullptr, src, (int)(n / per_channel), height, width, kernel_h, kernel_w,
                   pad_h, pad_w, stride_h, stride_w, dilation_h, dilation_w, height_col, width_col, dst, flags);
    if (rc == MOBULA_OK && !host_context(device_id)) {
      DeviceGuard guard(device_id);
      if (cudaStreamSynchronize(nullptr) != cudaSuccess) {
        fail(MOBULA_ERR_CUDA, "synchronize: %s", cudaGetErrorString(cudaGetLastError()));
        rc = MOBULA_ERR_CUDA;
      }
    }
  }
  if (rc != MOBULA_OK) die_like_reference(where, rc);
}

MOBULA_EXPORT void im2col_341af5d3(const int device_id, const int n, const float* data_im, const int height,
                                   const int width, const int kernel_h, const int kernel_w, const int pad_h,
                                   const int pad_w, const int stride_h, const int stride_w, const int dilation_h,
                                   const int dilation_w, const int height_col, const int width_col,
                                   float* data_col) {
  conv_dropin(true, "im2col", device_id, n, data_im, height, width, kernel_h, kernel_w, pad_h, pad_w, stride_h,
              stride_w, dilation_h, dilation_w, height_col, width_col, data_col);
}

MOBULA_EXPORT void col2im_341af5d3(const int device_id, const int n, const float* data_col, const int height,
                                   const int width, const int kernel_h, const int kernel_w, const int pad_h,
                                   const int pad_w, const int stride_h, const int stride_w, const int dilation_h,
                                   const int dilation_w, const int height_col, const int width_col,
                                   float* data_im) {
  conv_dropin(false, "col2im", device_id, n, data_col, height, width, kernel_h, kernel_w, pad_h, pad_w, stride_h,
              stride_w, dilation_h, dilation_w, height_col, width_col, data_im);
}

MOBULA_EXPORT void set_device(const int device_id) {
  if (device_id < 0) return;
  int current = -1;
  if (cudaGetDevice(&current) != cudaSuccess || current != device_id) {
    if (cudaSetDevice(device_id) != cudaSuccess) {
      fail(MOBULA_ERR_CUDA, "cudaSetDevice(%d): %s", device_id,
           cudaGetErrorString(cudaGetLastError()));
      die_like_reference("set_device", MOBULA_ERR_CUDA);
    }
  }
}
