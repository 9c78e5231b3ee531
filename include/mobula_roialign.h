/*
 * mobula_roialign.h -- C ABI of libmobula_roialign_sm100.so
 *
 * The B200 (sm_100a) replacement for the native side of MobulaOP's ROIAlign
 * operator.  Plain C: pointers, ints and floats only; no C++/torch types.
 *
 * Two groups of entry points:
 *
 *  1. DROP-IN symbols.  Exactly the extern "C" functions the reference's JIT code
 *     generator emits for opzoo/ROIAlign/ROIAlign.cpp and that
 *     mobula/func.py:155-156 calls through ctypes (wrapper template
 *     mobula/op/templates/kernel_code.cpp:1-5, symbol naming func.py:13-50,
 *     `set_device` from mobula/cpp/include/context/context.h:15-17).  Same names,
 *     same argument order, same synchronous semantics.
 *
 *  2. `mobula_roialign_*` v2 entry points: explicit CUDA stream, status return,
 *     number of images (needed to fuse the dX zero-fill of ROIAlign.py:33-34 into
 *     the backward kernel), backward mode and algorithm selection.  The Python
 *     host layer (mobulaop_b200/func.py) uses these.
 *
 * Tensor conventions (reference: opzoo/ROIAlign/ROIAlign.cpp:9-16, 78-83):
 *   bottom_data / bottom_diff : [N, C, H, W]   fp32, contiguous (NCHW)
 *   bottom_rois               : [R, 5]         fp32, rows = [batch_idx, x1, y1, x2, y2]
 *   top_data / top_diff       : [R, C, PH, PW] fp32, contiguous
 * The caller owns every buffer.  The library keeps a small per-device workspace
 * (ROI order tables, staging buffers for the host-buffer path) that is released by
 * mobula_roialign_release().
 */
#ifndef MOBULA_ROIALIGN_H_
#define MOBULA_ROIALIGN_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MOBULA_ROIALIGN_ABI_VERSION 1

/* status codes returned by the v2 entry points */
#define MOBULA_OK 0
#define MOBULA_ERR_INVALID_ARGUMENT 1
#define MOBULA_ERR_CUDA 2
#define MOBULA_ERR_UNSUPPORTED 3

/* backward `mode` */
#define MOBULA_ROIALIGN_BWD_ATOMIC 0        /* fastest; rounding order not reproducible   */
#define MOBULA_ROIALIGN_BWD_DETERMINISTIC 1 /* bit-exact: canonical ascending-index order */

/* `flags` (bit mask) */
#define MOBULA_ROIALIGN_ACCUMULATE 1u   /* fwd: top_data += result; bwd: do not zero-fill dX */
#define MOBULA_ROIALIGN_HOST_BUFFERS 2u /* all pointers are HOST memory: the library stages
                                           them to the device, runs, copies the result back
                                           and synchronises before returning               */
#define MOBULA_ROIALIGN_ALIGNED 4u      /* half-pixel box coordinates, boxes not forced to one pixel:
                                           torchvision.ops.roi_align(aligned=True) / detectron2
                                           ROIAlignV2 instead of the reference operator's convention.
                                           Served by the sweep / row-tile kernels (pooled_w <= 16)   */
/* testing aid: build no ROI tables, every kernel decodes ROIs on the fly (slow path) */
#define MOBULA_ROIALIGN_DEBUG_NO_TABLES 0x10000u
/* algorithm selection, bits 8..11 (0 = automatic) */
#define MOBULA_ROIALIGN_ALGO_SHIFT 8
#define MOBULA_ROIALIGN_ALGO_MASK (0xFu << MOBULA_ROIALIGN_ALGO_SHIFT)
#define MOBULA_ROIALIGN_ALGO_AUTO 0u
#define MOBULA_ROIALIGN_ALGO_GATHER 1u /* per-(roi,bin) threads, taps from global memory   */
#define MOBULA_ROIALIGN_ALGO_PLANE 2u  /* one CTA owns a whole (image, channel) plane in
                                          shared memory (TMA bulk copy in / out); handles
                                          adaptive sampling grids; its canonical-order
                                          backward is the deterministic mode                */
#define MOBULA_ROIALIGN_ALGO_RESIDENT 3u /* fixed sampling ratio: padded plane resident in
                                          shared memory, lean per-call ROI tables, work split
                                          evenly over the SMs                               */
#define MOBULA_ROIALIGN_ALGO_SWEEP 4u  /* lane = channel: ring of full-width rows of 32 channels
                                          in shared memory, bank-conflict-free taps, any
                                          sampling grid; needs pooled width <= 16 and a known
                                          num_images                                        */
#define MOBULA_ROIALIGN_ALGO_PULL 5u   /* backward only, fixed sampling ratio: per-pixel lists of
                                          (roi, bin, weight) contributions built once per call, one
                                          warp per pixel sums them over 32..128 channels per pass
                                          (lane = channel): no atomics, every dX element written
                                          once; merged-tap and bit-exact deterministic variants.
                                          As a forward algorithm it means AUTO                  */

/* ------------------------------------------------------------------------- *
 * 1. Drop-in symbols (reference ABI)                                         *
 * ------------------------------------------------------------------------- */

/* Replaces the generated wrapper around roi_align_forward_kernel<float>
 * (/root/reference/opzoo/ROIAlign/ROIAlign.cpp:9-16).  device_id >= 0: device
 * pointers on that GPU; device_id < 0 (the reference's CPU context,
 * mobula/func.py:139-141): HOST pointers, staged through the current GPU.
 * Synchronous: returns after the result is complete (reference:
 * mobula/cpp/include/context/hip_ctx.h:71-86).  Errors print to stderr and
 * abort like the reference's CHECK macros (logging.h:23-26). */
void roi_align_forward_ab78de3c(const int device_id, const int nthreads,
                                const float* bottom_data, const float spatial_scale,
                                const int channels, const int height, const int width,
                                const int pooled_height, const int pooled_width,
                                const int sampling_ratio, const float* bottom_rois,
                                float* top_data);

/* Replaces the wrapper around roi_align_backward_kernel<float>
 * (ROIAlign.cpp:78-83).  Adds into bottom_diff, which the caller has zeroed
 * (ROIAlign.py:33-34) -- note the argument order: bottom_diff BEFORE bottom_rois. */
void roi_align_backward_f318a24e(const int device_id, const int nthreads,
                                 const float* top_diff, const float spatial_scale,
                                 const int channels, const int height, const int width,
                                 const int pooled_height, const int pooled_width,
                                 const int sampling_ratio, float* bottom_diff,
                                 const float* bottom_rois);

/* mobula/cpp/src/context.cpp:9-16 */
void set_device(const int device_id);

/* ------------------------------------------------------------------------- *
 * 2. v2 entry points                                                         *
 * ------------------------------------------------------------------------- */

int mobula_roialign_abi_version(void);

/* Human-readable description of the last failure on the calling thread. */
const char* mobula_last_error(void);

/* Forward bilinear pool.  `stream` is a cudaStream_t (NULL = legacy default
 * stream); the call is asynchronous unless MOBULA_ROIALIGN_HOST_BUFFERS is set.
 * `num_images` = N, or -1 if unknown (disables the batch-index range check and the
 * PLANE algorithm).
 * Device-pointer calls may be captured into a CUDA graph: they do not synchronise,
 * allocate or read back once the per-device workspace has its size, so run the same
 * call once before the capture (a capture that would have to grow the workspace fails
 * with MOBULA_ERR_UNSUPPORTED).  The workspace is shared by all streams of a device;
 * eager calls order themselves on it, graph replays must be ordered against other
 * users of the library on the same device by the caller. */
int mobula_roialign_forward_f32(int device_id, void* stream, const float* bottom_data,
                                const float* bottom_rois, float* top_data, int num_rois,
                                int num_images, int channels, int height, int width,
                                int pooled_height, int pooled_width, float spatial_scale,
                                int sampling_ratio, unsigned flags);

/* Backward gradient scatter.  Unless MOBULA_ROIALIGN_ACCUMULATE is set the result
 * REPLACES bottom_diff (the zero-fill is fused), which requires num_images >= 0.
 * The ROI gradient is identically zero (ROIAlign.py:41-42) and is not produced. */
int mobula_roialign_backward_f32(int device_id, void* stream, const float* top_diff,
                                 const float* bottom_rois, float* bottom_diff, int num_rois,
                                 int num_images, int channels, int height, int width,
                                 int pooled_height, int pooled_width, float spatial_scale,
                                 int sampling_ratio, int mode, unsigned flags);

/* Half-precision and channel-last pooled tensors (the consumers of the
 * [R, C, PH, PW] tensor -- box-head FC, mask-head conv -- take fp16 / bf16, often channel-last).
 * The arithmetic stays fp32 and in the reference's order: the forward output is the fp32 result
 * rounded ONCE to nearest-even when it is stored, the backward widens dy exactly when it loads
 * it.  Feature map, ROIs and bottom_diff stay fp32 NCHW.
 *   top_dtype : MOBULA_DTYPE_F32 | _F16 | _BF16
 *   top_layout: MOBULA_LAYOUT_RCHW = [R, C, PH, PW] (the reference)
 *               MOBULA_LAYOUT_RHWC = [R, PH, PW, C] (channel-last)
 * Forward: sampling_ratio 2 resident kernels (any type, RCHW) or the sweep kernels (any type, either
 * layout; pooled width <= 16); MOBULA_ROIALIGN_ACCUMULATE needs (F32, RCHW).  Backward: the pull
 * kernels, whose copy stage reads any type / layout (a channel-last dy skips their transpose). */
#define MOBULA_DTYPE_F32 0
#define MOBULA_DTYPE_F16 1
#define MOBULA_DTYPE_BF16 2
#define MOBULA_LAYOUT_RCHW 0
#define MOBULA_LAYOUT_RHWC 1
int mobula_roialign_forward_ex(int device_id, void* stream, const float* bottom_data,
                               const float* bottom_rois, void* top_data, int top_dtype, int top_layout,
                               int num_rois, int num_images, int channels, int height, int width,
                               int pooled_height, int pooled_width, float spatial_scale,
                               int sampling_ratio, unsigned flags);
int mobula_roialign_backward_ex(int device_id, void* stream, const void* top_diff, int top_dtype,
                                int top_layout, const float* bottom_rois, float* bottom_diff,
                                int num_rois, int num_images, int channels, int height, int width,
                                int pooled_height, int pooled_width, float spatial_scale,
                                int sampling_ratio, int mode, unsigned flags);

/* Prepared ROI tables: what depends on the ROI table alone -- validation, a
 * private device copy of the [R, 5] rows and their stable partition by image -- is done once per
 * table and shared by the forward and the backward of a step (and by every call that pools the same
 * boxes).  The caller may reuse or free `bottom_rois` as soon as the stream has passed the prepare
 * call.  Device pointers only; `flags`: MOBULA_ROIALIGN_ALIGNED.  A plan belongs to its device;
 * destroy it with mobula_roialign_plan_destroy (which waits for the device).  The prepared calls
 * take the same flags as the v2 entry points (ACCUMULATE, algorithm selection) and obey the same
 * stream-capture rules. */
int mobula_roialign_prepare_f32(int device_id, void* stream, const float* bottom_rois, int num_rois,
                                int num_images, int height, int width, int pooled_height,
                                int pooled_width, float spatial_scale, int sampling_ratio,
                                unsigned flags, void** plan);
int mobula_roialign_forward_prepared_f32(void* plan, void* stream, const float* bottom_data,
                                         float* top_data, int channels, unsigned flags);
int mobula_roialign_backward_prepared_f32(void* plan, void* stream, const float* top_diff,
                                          float* bottom_diff, int channels, int mode, unsigned flags);
void mobula_roialign_plan_destroy(void* plan);

/* ROI bookkeeping as the kernels compute it, for bit-exact checking against the
 * oracle (same layout as oracle_roi_align_bookkeeping_f32):
 *   geom_i [R,4]  = batch, grid_h, grid_w, 0
 *   geom_f [R,5]  = start_w, start_h, bin_w, bin_h, count
 *   tap_i / tap_f [R, 2 (y,x), max(PH,PW), max_grid, 2] = (lo, hi) / (pos, frac);
 *   lo = hi = -1 for a coordinate outside the plane.  Device pointers. */
int mobula_roialign_bookkeeping_f32(int device_id, void* stream, const float* bottom_rois,
                                    int num_rois, int height, int width, int pooled_height,
                                    int pooled_width, float spatial_scale, int sampling_ratio,
                                    int32_t* geom_i, float* geom_f, int max_grid,
                                    int32_t* tap_i, float* tap_f);

/* Pinned host memory for callers that use MOBULA_ROIALIGN_HOST_BUFFERS. */
void* mobula_host_alloc(size_t bytes);
void mobula_host_free(void* ptr);

/* Number of kernel launches issued by this library since load (all threads). */
uint64_t mobula_roialign_launch_count(void);

/* Name of the kernel family the last forward / backward call of this process used
 * ("gather", "plane", "deterministic", ...). */
const char* mobula_roialign_last_algo(int backward);

/* Frees the per-device workspaces. */
void mobula_roialign_release(void);

/* ------------------------------------------------------------------------- *
 * 3. im2col / col2im of the opzoo Convolution operator     *
 * ------------------------------------------------------------------------- */

/* Drop-in symbols: replace the generated wrappers around im2col_kernel<float>
 * (/root/reference/opzoo/Convolution/Convolution.cpp:14-45) and col2im_kernel<float>
 * (:48-87).  Both kernels have the same C signature, hence the same hash
 * (mobula/func.py:13-50).  device_id < 0: host pointers, staged through the current
 * GPU; synchronous like the reference.
 *   im2col: n = C * KH * KW * OH * OW elements of data_col [C*KH*KW, OH, OW];
 *   col2im: n = C * H * W elements of data_im; every pixel is the sum of the column
 *           entries copied from it, added in ascending (h_col, w_col) order. */
void im2col_341af5d3(const int device_id, const int n, const float* data_im, const int height,
                     const int width, const int kernel_h, const int kernel_w, const int pad_h,
                     const int pad_w, const int stride_h, const int stride_w, const int dilation_h,
                     const int dilation_w, const int height_col, const int width_col,
                     float* data_col);
void col2im_341af5d3(const int device_id, const int n, const float* data_col, const int height,
                     const int width, const int kernel_h, const int kernel_w, const int pad_h,
                     const int pad_w, const int stride_h, const int stride_w, const int dilation_h,
                     const int dilation_w, const int height_col, const int width_col,
                     float* data_im);

/* v2: status code, explicit stream, asynchronous on device pointers.  `channels` is
 * explicit (the reference derives it from n).  flags: MOBULA_ROIALIGN_HOST_BUFFERS. */
int mobula_im2col_f32(int device_id, void* stream, const float* data_im, int channels, int height,
                      int width, int kernel_h, int kernel_w, int pad_h, int pad_w, int stride_h,
                      int stride_w, int dilation_h, int dilation_w, int height_col, int width_col,
                      float* data_col, unsigned flags);
int mobula_col2im_f32(int device_id, void* stream, const float* data_col, int channels, int height,
                      int width, int kernel_h, int kernel_w, int pad_h, int pad_w, int stride_h,
                      int stride_w, int dilation_h, int dilation_w, int height_col, int width_col,
                      float* data_im, unsigned flags);

#ifdef __cplusplus
}
#endif
#endif /* MOBULA_ROIALIGN_H_ */
