/*
 * oracle/ref_shim.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Exposes the UNMODIFIED reference kernels through the same extern "C" symbols
 * the reference's JIT code generator would emit
 * (/root/reference/mobula/op/templates/header.cpp:1-16 and
 *  kernel_code.cpp:1-5; names from mobula/func.py:13-50).  The reference
 * sources are compiled where they lie (see oracle/Makefile); nothing from
 * /root/reference is copied into this repository.  Output goes to oracle/_ref/.
 */
#include "mobula_op.h"
using namespace mobula;

#include "ROIAlign.cpp"  /* resolved through -I/root/reference/opzoo/ROIAlign */

extern "C" {

void roi_align_forward_ab78de3c(const int device_id, const int nthreads,
                                const float* bottom_data,
                                const float spatial_scale, const int channels,
                                const int height, const int width,
                                const int pooled_height, const int pooled_width,
                                const int sampling_ratio,
                                const float* bottom_rois, float* top_data) {
  KERNEL_RUN_BEGIN(device_id);
  KERNEL_RUN(roi_align_forward_kernel<float>)
  (nthreads, bottom_data, spatial_scale, channels, height, width, pooled_height,
   pooled_width, sampling_ratio, bottom_rois, top_data);
  KERNEL_RUN_END();
}

void roi_align_backward_f318a24e(const int device_id, const int nthreads,
                                 const float* top_diff,
                                 const float spatial_scale, const int channels,
                                 const int height, const int width,
                                 const int pooled_height,
                                 const int pooled_width,
                                 const int sampling_ratio, float* bottom_diff,
                                 const float* bottom_rois) {
  KERNEL_RUN_BEGIN(device_id);
  KERNEL_RUN(roi_align_backward_kernel<float>)
  (nthreads, top_diff, spatial_scale, channels, height, width, pooled_height,
   pooled_width, sampling_ratio, bottom_diff, bottom_rois);
  KERNEL_RUN_END();
}

int ref_host_threads(void) { return HOST_NUM_THREADS; }

}  // extern "C"
