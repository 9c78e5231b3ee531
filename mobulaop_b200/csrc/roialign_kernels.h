// Internal interface between the C-ABI layer (roialign_api.cu) and the kernel files.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mobula_b200 {

struct RoiAlignArgs {
  const float* rois;   // [R,5] device
  int num_rois;
  int num_images;      // -1 = unknown
  int channels, height, width;
  int c_count;         // resident kernels: work on channels [0, c_count) only (0 = all); `channels`
                       // stays the stride, so a channel range is a pointer offset + c_count
  int pooled_h, pooled_w;
  float scale;
  int sampling_ratio;
  int aligned;         // half-pixel box coordinates (torchvision aligned=True); sweep / tile kernels only
  // the pooled tensor (forward output / backward dy): element type 0 fp32 | 1 fp16 | 2 bf16 and
  // layout 0 [R, C, PH, PW] (the reference) | 1 [R, PH, PW, C] (channel-last); anything but (0, 0)
  // runs on the sampling_ratio == 2 resident forward (type only), the sweep forward and the pull backward
  int top_dtype, top_layout;
  // grouping of the ROI rows by image prepared earlier (mobula_roialign_prepare_f32): when set,
  // launch_roi_group copies it instead of recomputing it
  const int* pre_img_start;
  const int* pre_order;
};

void count_launch();

// ---- gather family (roialign_gather.cu): any shape, taps from global memory --------------
cudaError_t launch_fwd_gather(const RoiAlignArgs& a, const float* data, float* out, int accumulate,
                              cudaStream_t stream);
cudaError_t launch_bwd_gather(const RoiAlignArgs& a, const float* dy, float* dx, cudaStream_t stream);
cudaError_t launch_bookkeeping(const RoiAlignArgs& a, int32_t* geom_i, float* geom_f, int max_grid,
                               int32_t* tap_i, float* tap_f, cudaStream_t stream);

// ---- ROI tables (roialign_tables.cu) --------------------------------------------------------
// Channel-independent decoding of every ROI, done once per call and shared by all the
// (image, channel) planes: ROIs grouped by image, per-axis sample tables, and for the
// backward the inverse map "feature-map row/column -> samples that touch it".

// One sample coordinate along one axis (ROIAlign.cpp:59-65 + bilinear.h:15-48).
struct __align__(16) AxisEntry {
  int lo, hi;       // pixel index of the two taps (y axis: premultiplied by the plane width);
                    // lo < 0: the coordinate is outside the plane, the sample contributes nothing
  float w_hi;       // weight of `hi` (the fraction l)
  float w_lo;       // weight of `lo` (1 - l)
};

// A feature-map row (or column) some sample of the ROI touches and the contiguous range of
// sample indices k = p*grid + i that touch it (as lo or as hi): the inverse map the canonical
// (deterministic) backward walks.
struct __align__(16) AxisCell {
  int pix;          // row (premultiplied by the plane pitch) or column
  int p_first;      // pooled index of sample k_first (k = p * grid + i)
  int k_first, k_last;
};

struct __align__(16) RoiRec {
  int roi;          // row of the caller's roi table
  int grid_h, grid_w;
  float count;      // (float)(grid_h * grid_w)
  int ny, nx;       // pooled_h * grid_h, pooled_w * grid_w
  int yoff, xoff;   // first AxisEntry of each axis in the arena; -1: not tabled (arena full)
  int ycell, xcell; // first AxisCell of each axis
  int nycell, nxcell;
  int y_first, y_last;   // touched rows (plain row index, inclusive); first > last: none
  int x_first, x_last;
};

constexpr int kBands = 16;   // row bands per plane = warps of the canonical backward's CTAs

enum : unsigned { kNeedCells = 1u };

struct RoiTables {
  int* img_start;    // [N+2]: grouped position of the first ROI of image n; segment N holds the
                     //        ROIs whose batch index is outside [0, N)
  int* order;        // [R]   roi rows, grouped by image, ascending inside an image
  RoiRec* recs;      // [R]   indexed by grouped position
  AxisEntry* axis;   // [arena]
  AxisCell* cells;   // [2 * arena]
  long long arena;   // capacity of `axis` in entries
  int pitch;         // shared-memory row pitch the y offsets are premultiplied with
  int regular;       // fixed sampling ratio and an exact arena: roi j owns entries
                     // [j*E, (j+1)*E), E = (PH+PW)*sampling_ratio, y entries first
};

size_t roi_tables_bytes(const RoiAlignArgs& a, unsigned need);
RoiTables roi_tables_carve(const RoiAlignArgs& a, unsigned need, int pitch, void* base);
cudaError_t launch_roi_tables(const RoiAlignArgs& a, const RoiTables& t, unsigned need,
                              cudaStream_t stream);
// testing aid: an arena of zero entries leaves every ROI untabled
void roi_tables_set_arena_override(long long entries /* < 0: none */);

// ---- plane family (roialign_plane.cu): one CTA owns a whole (image, channel) plane in
// shared memory -----------------------------------------------------------------------------
struct PlanePlan {
  // forward: rows `pitch` floats apart (padded against bank conflicts: lanes of a warp read
  // the same columns of several rows)
  int pitch;
  size_t plane_bytes;  // height * pitch * 4
  // backward: dense rows (a warp touches one row per instruction) + per-warp dy staging
  int bwd_pitch;
  size_t bwd_plane_bytes;
  int gslot_floats;    // staging floats per warp; 0: does not fit, dy is read in place
  int use_bulk;        // rows can be moved with cp.async.bulk (16-byte alignment holds)
  int ctas_per_sm;     // resident CTAs per SM (forward footprint)
  int bwd_ctas_per_sm;
};

bool plane_supported(const RoiAlignArgs& a, const void* plane_ptr, int smem_optin, PlanePlan* plan);
cudaError_t launch_fwd_plane(const RoiAlignArgs& a, const PlanePlan& plan, const RoiTables& t,
                             const float* data, float* out, int accumulate, int num_sms,
                             cudaStream_t stream);
// Owner-computes backward: every dX element is produced by exactly one thread and written
// once (zero-fill fused); contributions are added one by one in the order of the single-thread
// reference (bit-exact): the deterministic mode.
cudaError_t launch_bwd_plane(const RoiAlignArgs& a, const PlanePlan& plan, const RoiTables& t,
                             const float* dy, float* dx, int accumulate, int num_sms,
                             cudaStream_t stream);

// ---- resident family (roialign_resident.cu): fixed sampling ratio, padded plane resident in
// shared memory, lean tables; the default whenever it applies ------------------------------
// Stable partition of the ROI rows by image (roialign_tables.cu): img_start[N+2], order[R].
cudaError_t launch_roi_group(const RoiAlignArgs& a, int* img_start, int* order, cudaStream_t stream);

struct ResidentPlan {
  int grid;            // sampling ratio
  int pitch;           // floats between rows of the padded plane in shared memory
  int per_roi;         // sample entries per ROI: (PH + PW) * grid
  int fwd_two;         // the sampling_ratio == 2 forward (lane = sample column) applies
  int fwd_pair;        // ... on a channel pair (two padded planes fit shared memory)
  int ypairs, xslots;  // its table: row records (pairs) and column records per ROI
  size_t plane_bytes;  // (H + 1) * pitch * 4
  size_t fwd_smem;
  int bwd_ok;          // the resident backward applies
  int bwd_slabs, bwd_slab_rows, bwd_maxc, bwd_maxe, bwd_pitch2;   // bands per plane, rows per band, ...
  int bwd_ntx, bwd_tile_cols;   // column tiles per band, columns per tile
  size_t bwd_smem;
  int vec16;           // plane rows are 16-byte aligned in global memory: cp.async 16-byte loads
};

bool resident_supported(const RoiAlignArgs& a, const void* plane_ptr, int smem_optin, ResidentPlan* plan);
size_t resident_fwd_workspace(const RoiAlignArgs& a, const ResidentPlan& plan);
cudaError_t launch_fwd_resident(const RoiAlignArgs& a, const ResidentPlan& plan, void* workspace,
                                const float* data, float* out, int accumulate, int num_sms,
                                cudaStream_t stream);
size_t resident_bwd_workspace(const RoiAlignArgs& a, const ResidentPlan& plan, int deterministic);
// dX := scatter (accumulate = 0, zero-fill fused) or dX += scatter.  deterministic = 0: merged
// weights, reproducible run to run, within 1e-5 of the reference; 1: every contribution added
// separately in the reference's single-thread order, bit-identical dX.
cudaError_t launch_bwd_resident(const RoiAlignArgs& a, const ResidentPlan& plan, void* workspace,
                                const float* dy, float* dx, int accumulate, int deterministic, int num_sms,
                                cudaStream_t stream);

// ---- sweep family (roialign_sweep.cu): lane = channel, ring of full-width rows of 32 channels
// (odd channel pitch: bank-conflict-free taps), items sorted by row; any sampling grid, any
// pooled height, pooled width <= 16 ---------------------------------------------------------
struct SweepPlan {
  int pitch, rowf;            // floats between channels of a ring row (odd), floats per ring row
  int ring, span;             // ring rows; rows one item's taps may span
  int pwp;                    // staging pitch
  int segs;                   // row segments per image
  int colcap, rowcap;         // column / row records per ROI
  size_t smem;
};

bool sweep_supported(const RoiAlignArgs& a, const void* plane_ptr, int smem_optin, int num_sms, SweepPlan* plan);
size_t sweep_fwd_workspace(const RoiAlignArgs& a, const SweepPlan& plan);
cudaError_t launch_fwd_sweep(const RoiAlignArgs& a, const SweepPlan& plan, void* workspace, const float* data,
                             float* out, int accumulate, int num_sms, cudaStream_t stream);

// backward of the sweep family: dX row tiles in shared memory, strip-owner warps, ROIs visited in
// ascending order; deterministic = every tap added separately in the reference's order
struct TilePlan {
  int pitch, tr, strip, pwp, colcap, rowcap;
  size_t smem;
};
bool tile_supported(const RoiAlignArgs& a, int smem_optin, int num_sms, TilePlan* plan);
size_t tile_bwd_workspace(const RoiAlignArgs& a, const TilePlan& plan);
cudaError_t launch_bwd_tile(const RoiAlignArgs& a, const TilePlan& plan, void* workspace, const float* dy,
                            float* dx, int accumulate, int deterministic, int num_sms, cudaStream_t stream);

// ---- pull family (roialign_pull.cu): backward as a gather over per-pixel contribution lists,
// lane = channel; fixed sampling ratios -------------------------------------------------------
// lane = channel; fixed sampling ratios and adaptive grids.  `entries` = capacity of the list array (adaptive
// grids: the exact number is known only after the geometry kernel; launch_bwd_pull reports it in
// *need_entries -- and launches nothing else -- when the workspace is too small; ~0ull: beyond 32-bit lists)
bool pull_supported(const RoiAlignArgs& a);
size_t pull_bwd_workspace(const RoiAlignArgs& a, int deterministic, unsigned long long entries);
cudaError_t launch_bwd_pull(const RoiAlignArgs& a, void* workspace, unsigned long long entries, const float* dy,
                            float* dx, int accumulate, int deterministic, int num_sms, cudaStream_t stream,
                            unsigned long long* need_entries);

// ---- im2col / col2im (im2col.cu): the opzoo Convolution operator's kernels -----------------
cudaError_t launch_im2col(const float* im, int channels, int height, int width, int kernel_h, int kernel_w,
                          int pad_h, int pad_w, int stride_h, int stride_w, int dilation_h, int dilation_w,
                          int height_col, int width_col, float* col, int num_sms, cudaStream_t stream);
cudaError_t launch_col2im(const float* col, int channels, int height, int width, int kernel_h, int kernel_w,
                          int pad_h, int pad_w, int stride_h, int stride_w, int dilation_h, int dilation_w,
                          int height_col, int width_col, float* im, int num_sms, cudaStream_t stream);

}  // namespace mobula_b200
