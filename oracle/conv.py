"""oracle/conv.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Checkers for im2col / col2im:

* ``numpy_im2col / numpy_col2im`` -- NumPy restatement of
  /root/reference/opzoo/Convolution/Convolution.cpp:14-45 (im2col_kernel: column element
  (c, kh, kw, oh, ow) = image (c, oh*sh - ph + kh*dh, ow*sw - pw + kw*dw) or 0) and :48-87
  (col2im_kernel: every image pixel sums the column entries copied from it, in ascending
  (h_col, w_col) order, starting from 0 -- fp32 adds, so the order is part of the result).
* ``RefConv`` -- ctypes binding of oracle/_ref/libmobula_ref_conv.so, the reference's own
  Convolution.cpp compiled where it lies (oracle/Makefile `make ref`).

Parity status: PINNED -- the restatement is bit-identical to the compiled reference on the seeded
cases of tests/golden/conv_*.npz (tests/test_conv.py), which were generated from the compiled
reference by tests/golden/gen_conv_golden.py.  Only tests/ may import this module.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def out_size(x, k, p, s, d):
    """Convolution.cpp:97-100 / Convolution.py:85-87."""
    return (x + 2 * p - (d * (k - 1) + 1)) // s + 1


def numpy_im2col(im, kernel, pad, stride, dilation):
    C, H, W = im.shape
    (KH, KW), (PH, PW), (SH, SW), (DH, DW) = kernel, pad, stride, dilation
    OH, OW = out_size(H, KH, PH, SH, DH), out_size(W, KW, PW, SW, DW)
    col = np.zeros((C, KH, KW, OH, OW), np.float32)
    rows = np.arange(OH) * SH - PH
    cols = np.arange(OW) * SW - PW
    for kh in range(KH):
        r = rows + kh * DH
        rv = (r >= 0) & (r < H)
        for kw in range(KW):
            c = cols + kw * DW
            cv = (c >= 0) & (c < W)
            block = im[:, np.clip(r, 0, H - 1)[:, None], np.clip(c, 0, W - 1)[None, :]]
            col[:, kh, kw] = np.where(rv[:, None] & cv[None, :], block, np.float32(0))
    return col.reshape(C * KH * KW, OH, OW)


def numpy_col2im(col, im_shape, kernel, pad, stride, dilation):
    C, H, W = im_shape
    (KH, KW), (PH, PW), (SH, SW), (DH, DW) = kernel, pad, stride, dilation
    OH, OW = out_size(H, KH, PH, SH, DH), out_size(W, KW, PW, SW, DW)
    col = col.reshape(C, KH, KW, OH, OW)
    im = np.zeros((C, H, W), np.float32)
    # ascending (h_col, w_col) per pixel (Convolution.cpp:67-83): the outer loops below visit
    # every (h_col, w_col) in that order and add its KH*KW entries to DISTINCT pixels, so each
    # pixel receives its terms in the reference's order
    for oh in range(OH):
        for ow in range(OW):
            for kh in range(KH):
                h = oh * SH - PH + kh * DH
                if h < 0 or h >= H:
                    continue
                for kw in range(KW):
                    w = ow * SW - PW + kw * DW
                    if 0 <= w < W:
                        im[:, h, w] = im[:, h, w] + col[:, kh, kw, oh, ow]
    return im


class RefConv:
    """The reference's own kernels (CPU, OpenMP); col2im is a gather, so threads do not change it."""

    PATH = os.path.join(_HERE, '_ref', 'libmobula_ref_conv.so')

    @classmethod
    def available(cls):
        return os.path.exists(cls.PATH)

    def __init__(self):
        self.lib = ctypes.CDLL(self.PATH)
        self.lib.im2col_341af5d3.restype = None
        self.lib.col2im_341af5d3.restype = None

    @staticmethod
    def _geom(H, W, kernel, pad, stride, dilation, OH, OW):
        vals = (H, W, kernel[0], kernel[1], pad[0], pad[1], stride[0], stride[1], dilation[0], dilation[1], OH, OW)
        return [ctypes.c_int(int(v)) for v in vals]

    def im2col(self, im, kernel, pad, stride, dilation):
        im = np.ascontiguousarray(im, np.float32)
        C, H, W = im.shape
        OH, OW = out_size(H, kernel[0], pad[0], stride[0], dilation[0]), out_size(W, kernel[1], pad[1], stride[1], dilation[1])
        col = np.empty((C * kernel[0] * kernel[1], OH, OW), np.float32)
        self.lib.im2col_341af5d3(ctypes.c_int(-1), ctypes.c_int(col.size), ctypes.c_void_p(im.ctypes.data),
                                 *self._geom(H, W, kernel, pad, stride, dilation, OH, OW),
                                 ctypes.c_void_p(col.ctypes.data))
        return col

    def col2im(self, col, im_shape, kernel, pad, stride, dilation):
        col = np.ascontiguousarray(col, np.float32)
        C, H, W = im_shape
        OH, OW = out_size(H, kernel[0], pad[0], stride[0], dilation[0]), out_size(W, kernel[1], pad[1], stride[1], dilation[1])
        im = np.empty((C, H, W), np.float32)
        self.lib.col2im_341af5d3(ctypes.c_int(-1), ctypes.c_int(im.size), ctypes.c_void_p(col.ctypes.data),
                                 *self._geom(H, W, kernel, pad, stride, dilation, OH, OW),
                                 ctypes.c_void_p(im.ctypes.data))
        return im
