/*
 * oracle/conv_shim.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * The UNMODIFIED reference im2col_kernel / col2im_kernel
 * (/root/reference/opzoo/Convolution/Convolution.cpp:14-87, compiled where it lies, see
 * oracle/Makefile) behind the extern "C" symbols the reference's code generator would emit
 * (mobula/op/templates/kernel_code.cpp:1-5; names from mobula/func.py:13-50: both kernels have
 * the same C signature, hence the same hash).  Output goes to oracle/_ref/.
 */
#include "mobula_op.h"
using namespace mobula;

#include "Convolution.cpp" /* resolved through -I/root/reference/opzoo/Convolution */

extern "C" {

void im2col_341af5d3(const int device_id, const int n, const float* data_im, const int height,
                     const int width, const int kernel_h, const int kernel_w, const int pad_h,
                     const int pad_w, const int stride_h, const int stride_w,
                     const int dilation_h, const int dilation_w, const int height_col,
                     const int width_col, float* data_col) {
  KERNEL_RUN_BEGIN(device_id);
  KERNEL_RUN(im2col_kernel<float>)
  (n, data_im, height, width, kernel_h, kernel_w, pad_h, pad_w, stride_h, stride_w, dilation_h,
   dilation_w, height_col, width_col, data_col);
  KERNEL_RUN_END();
}

void col2im_341af5d3(const int device_id, const int n, const float* data_col, const int height,
                     const int width, const int kernel_h, const int kernel_w, const int pad_h,
                     const int pad_w, const int stride_h, const int stride_w,
                     const int dilation_h, const int dilation_w, const int height_col,
                     const int width_col, float* data_im) {
  KERNEL_RUN_BEGIN(device_id);
  KERNEL_RUN(col2im_kernel<float>)
  (n, data_col, height, width, kernel_h, kernel_w, pad_h, pad_w, stride_h, stride_w, dilation_h,
   dilation_w, height_col, width_col, data_im);
  KERNEL_RUN_END();
}

}  // extern "C"
