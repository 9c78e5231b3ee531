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

// ---- gather family (roialign_gather.cu) ----------------------------------------
cudaError_t launch_fwd_gather(const RoiAlignArgs& a, const float* data, float* out, int accumulate,
                              cudaStream_t stream);
cudaError_t launch_bwd_gather(const RoiAlignArgs& a, const float* dy, float* dx, cudaStream_t stream);
cudaError_t launch_bookkeeping(const RoiAlignArgs& a, int32_t* geom_i, float* geom_f, int max_grid,
                               int32_t* tap_i, float* tap_f, cudaStream_t stream);

// ---- plane family (roialign_plane.cu) --------------------------------------------
// One 16-byte entry per (roi, pooled index, grid index) and axis: byte offsets of the
// two pixels inside the shared-memory plane and their weights.
struct __align__(16) AxisEntry {
  int lo_off, hi_off;   // y: row * pitch * 4, x: col * 4; lo_off < 0 marks "outside"
  float w_hi, w_lo;     // weight of hi (fraction) / of lo (1 - fraction)
};

struct PlanePlan {
  int pitch;            // floats per shared-memory row
  int nbuf;             // plane buffers resident per CTA
  size_t smem_bytes;
  int use_bulk;         // rows can be moved with cp.async.bulk (16-byte alignment holds)
};

// Workspace the plane kernels need, all device pointers.
struct PlaneWorkspace {
  int* img_start;       // [N+1] start of each image's run in `order`
  int* order;           // [R]   roi indices grouped by image, ascending inside an image
  AxisEntry* ytab;      // [R][PH*sr]  indexed by position in `order`
  AxisEntry* xtab;      // [R][PW*sr]
};

bool plane_supported(const RoiAlignArgs& a, const void* p0, const void* p1, int device,
                     PlanePlan* plan);
size_t plane_workspace_bytes(const RoiAlignArgs& a);
PlaneWorkspace plane_workspace_carve(const RoiAlignArgs& a, void* base);
cudaError_t launch_plane_prepare(const RoiAlignArgs& a, const PlanePlan& plan,
                                 const PlaneWorkspace& ws, cudaStream_t stream);
cudaError_t launch_fwd_plane(const RoiAlignArgs& a, const PlanePlan& plan, const PlaneWorkspace& ws,
                             const float* data, float* out, int accumulate, int num_sms,
                             cudaStream_t stream);
cudaError_t launch_bwd_plane(const RoiAlignArgs& a, const PlanePlan& plan, const PlaneWorkspace& ws,
                             const float* dy, float* dx, int accumulate, int num_sms,
                             cudaStream_t stream);

// ---- deterministic backward (roialign_det.cu) -------------------------------------
cudaError_t launch_bwd_deterministic(const RoiAlignArgs& a, const float* dy, float* dx,
                                     int accumulate, int num_sms, cudaStream_t stream);

}  // namespace mobula_b200
