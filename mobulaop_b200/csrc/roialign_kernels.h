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
  int pooled_h, pooled_w;
  float scale;
  int sampling_ratio;
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

// A feature-map row (or column) some sample of the ROI touches, and the contiguous range
// of sample indices k = p*grid + i that touch it (as lo or as hi).
struct __align__(16) AxisCell {
  int pix;          // row (premultiplied by the plane width) or column
  int k_first, k_last;
  int p_first;      // pooled index of sample k_first (k = p * grid + i)
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

struct RoiTables {
  int* img_start;    // [N+2]: grouped position of the first ROI of image n; segment N holds the
                     //        ROIs whose batch index is outside [0, N)
  int* order;        // [R]   roi rows, grouped by image, ascending inside an image
  RoiRec* recs;      // [R]   indexed by grouped position
  AxisEntry* axis;   // [arena]
  AxisCell* cells;   // [2 * arena]
  int* scan_tmp;     // [R + 1]
  long long arena;   // capacity of `axis` in entries
};

size_t roi_tables_bytes(const RoiAlignArgs& a);
RoiTables roi_tables_carve(const RoiAlignArgs& a, void* base);
// need_cells: also build the inverse maps (backward).
cudaError_t launch_roi_tables(const RoiAlignArgs& a, const RoiTables& t, bool need_cells,
                              cudaStream_t stream);
// testing aid: an arena of zero entries leaves every ROI untabled
void roi_tables_set_arena_override(long long entries /* < 0: none */);

// ---- plane family (roialign_plane.cu): one CTA owns a whole (image, channel) plane in
// shared memory -----------------------------------------------------------------------------
struct PlanePlan {
  size_t smem_bytes;   // dynamic shared memory per CTA
  int threads;
  int use_bulk;        // plane can be moved with cp.async.bulk (16-byte alignment holds)
  int ctas_per_sm;     // resident CTAs per SM at this shared-memory footprint
};

bool plane_supported(const RoiAlignArgs& a, const void* plane_ptr, int smem_optin, PlanePlan* plan);
cudaError_t launch_fwd_plane(const RoiAlignArgs& a, const PlanePlan& plan, const RoiTables& t,
                             const float* data, float* out, int accumulate, int num_sms,
                             cudaStream_t stream);
// Owner-computes backward: every dX element is produced by exactly one thread, in the
// canonical order of the single-thread reference, and written once (zero-fill fused).
cudaError_t launch_bwd_plane(const RoiAlignArgs& a, const PlanePlan& plan, const RoiTables& t,
                             const float* dy, float* dx, int accumulate, int num_sms,
                             cudaStream_t stream);

}  // namespace mobula_b200
