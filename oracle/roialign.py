"""oracle/roialign.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Python access to the checkers for the ROIAlign hot path:

* ``COracle``      -- ctypes binding of oracle/roialign_oracle.c (our plain-C
                      restatement; single-thread canonical order and an OpenMP
                      variant), built by ``oracle/Makefile`` (``make oracle``).
* ``RefLib``       -- ctypes binding of oracle/_ref/libmobula_ref_{omp,st}.so, the
                      reference's own C++ sources compiled where they lie
                      (``make ref``; only possible where /root/reference exists,
                      the prebuilt files travel to the GPU box).
* ``numpy_forward / numpy_backward / numpy_bookkeeping`` -- a NumPy fp32
                      restatement used for ROI bookkeeping (batch index, grid
                      sizes, sample coordinates, tap indices) and for counting
                      the unique feature-map elements a workload touches.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl
reference`` legs may import this module.  Parity status: PINNED (see the header
of roialign_oracle.c and tests/test_oracle.py).

Reference being restated: /root/reference/opzoo/ROIAlign/ROIAlign.cpp:9-161 and
bilinear.h:10-112.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_F32P = ctypes.POINTER(ctypes.c_float)
_I32P = ctypes.POINTER(ctypes.c_int32)


def build(verbose=False):
    """Compile the C restatement and, where the reference tree exists, oracle/_ref."""
    out = subprocess.run(['make', '-C', _HERE, 'all'], capture_output=True, text=True)
    if verbose or out.returncode != 0:
        print(out.stdout, out.stderr)
    if out.returncode != 0:
        raise RuntimeError('oracle build failed')


def _fp(a):
    return a.ctypes.data_as(_F32P)


def _check(data, rois, pooled):
    assert data.dtype == np.float32 and data.flags.c_contiguous and data.ndim == 4
    assert rois.dtype == np.float32 and rois.flags.c_contiguous
    assert rois.ndim == 2 and rois.shape[1] == 5
    assert len(pooled) == 2


class COracle:
    """ctypes view of oracle/_lib/liboracle_roialign.so."""

    def __init__(self, path=None):
        path = path or os.path.join(_HERE, '_lib', 'liboracle_roialign.so')
        if not os.path.exists(path):
            build()
        self.lib = ctypes.CDLL(path)
        i64, i32, f32 = ctypes.c_int64, ctypes.c_int, ctypes.c_float
        common = [i64, _F32P, f32, i32, i32, i32, i32, i32, i32]
        self.lib.oracle_roi_align_forward_f32.argtypes = common + [_F32P, _F32P]
        self.lib.oracle_roi_align_backward_f32.argtypes = common + [_F32P, _F32P]
        self.lib.oracle_roi_align_forward_f32_mt.argtypes = common + [_F32P, _F32P, i32]
        self.lib.oracle_roi_align_backward_f32_mt.argtypes = common + [_F32P, _F32P, i32]
        self.lib.oracle_roi_align_forward_f32_mt.restype = i32
        self.lib.oracle_roi_align_backward_f32_mt.restype = i32
        self.lib.oracle_roi_align_bookkeeping_f32.argtypes = [
            i32, f32, i32, i32, i32, i32, i32, _F32P, _I32P, _F32P, i32, _I32P, _F32P]
        self.lib.oracle_max_threads.restype = i32

    def max_threads(self):
        return int(self.lib.oracle_max_threads())

    def forward(self, data, rois, pooled, scale, sr, threads=1):
        _check(data, rois, pooled)
        _, C, H, W = data.shape
        R = rois.shape[0]
        out = np.empty((R, C, pooled[0], pooled[1]), np.float32)
        if out.size == 0:
            return out
        args = (out.size, _fp(data), scale, C, H, W, pooled[0], pooled[1], sr, _fp(rois), _fp(out))
        if threads == 1:
            self.lib.oracle_roi_align_forward_f32(*args)
        else:
            self.lib.oracle_roi_align_forward_f32_mt(*args, threads)
        return out

    def backward(self, dy, rois, data_shape, scale, sr, threads=1, dx=None):
        """Returns dX (zero-filled here unless `dx` is given, in which case it is added to)."""
        R, C, PH, PW = dy.shape
        assert dy.dtype == np.float32 and dy.flags.c_contiguous
        assert rois.dtype == np.float32 and rois.flags.c_contiguous and rois.shape == (R, 5)
        N, C2, H, W = data_shape
        assert C2 == C
        if dx is None:
            dx = np.zeros(data_shape, np.float32)
        if dy.size == 0:
            return dx
        args = (dy.size, _fp(dy), scale, C, H, W, PH, PW, sr, _fp(dx), _fp(rois))
        if threads == 1:
            self.lib.oracle_roi_align_backward_f32(*args)
        else:
            self.lib.oracle_roi_align_backward_f32_mt(*args, threads)
        return dx

    def bookkeeping(self, rois, hw, pooled, scale, sr, max_grid):
        R = rois.shape[0]
        pmax = max(pooled)
        geom_i = np.zeros((R, 4), np.int32)
        geom_f = np.zeros((R, 5), np.float32)
        tap_i = np.full((R, 2, pmax, max(max_grid, 1), 2), -7, np.int32)
        tap_f = np.full((R, 2, pmax, max(max_grid, 1), 2), np.nan, np.float32)
        self.lib.oracle_roi_align_bookkeeping_f32(
            R, scale, hw[0], hw[1], pooled[0], pooled[1], sr, _fp(rois),
            geom_i.ctypes.data_as(_I32P), _fp(geom_f), max_grid,
            tap_i.ctypes.data_as(_I32P), _fp(tap_f))
        return geom_i, geom_f, tap_i, tap_f


class RefLib:
    """The reference's own kernels (oracle/_ref), called through its ABI.

    kind='st'  : -DHOST_NUM_THREADS=1, plain ascending loop, bit-reproducible.
    kind='omp' : -DHOST_NUM_THREADS=<cores>, OpenMP, `omp atomic` backward.
    """

    def __init__(self, kind='st'):
        path = os.path.join(_HERE, '_ref', 'libmobula_ref_%s.so' % kind)
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.kind = kind
        self.lib = ctypes.CDLL(path)
        i32, f32 = ctypes.c_int, ctypes.c_float
        common = [i32, i32, _F32P, f32, i32, i32, i32, i32, i32, i32]
        self.lib.roi_align_forward_ab78de3c.argtypes = common + [_F32P, _F32P]
        self.lib.roi_align_backward_f318a24e.argtypes = common + [_F32P, _F32P]
        self.lib.roi_align_forward_ab78de3c.restype = None
        self.lib.roi_align_backward_f318a24e.restype = None
        self.lib.ref_host_threads.restype = i32

    @staticmethod
    def available(kind='st'):
        return os.path.exists(os.path.join(_HERE, '_ref', 'libmobula_ref_%s.so' % kind))

    def host_threads(self):
        return int(self.lib.ref_host_threads())

    def forward(self, data, rois, pooled, scale, sr):
        _check(data, rois, pooled)
        _, C, H, W = data.shape
        out = np.empty((rois.shape[0], C, pooled[0], pooled[1]), np.float32)
        if out.size:
            self.lib.roi_align_forward_ab78de3c(
                -1, out.size, _fp(data), scale, C, H, W, pooled[0], pooled[1], sr,
                _fp(rois), _fp(out))
        return out

    def backward(self, dy, rois, data_shape, scale, sr):
        R, C, PH, PW = dy.shape
        _, _, H, W = data_shape
        dx = np.zeros(data_shape, np.float32)       # ROIAlign.py:33-34
        if dy.size:
            self.lib.roi_align_backward_f318a24e(
                -1, dy.size, _fp(dy), scale, C, H, W, PH, PW, sr, _fp(dx), _fp(rois))
        return dx


# ---------------------------------------------------------------------------
# NumPy fp32 restatement (bookkeeping + small-case values)
# ---------------------------------------------------------------------------
_f = np.float32


def numpy_geometry(rois, pooled, scale, sr):
    """Per-ROI quantities of ROIAlign.cpp:24-54, all in fp32, vectorised over ROIs."""
    rois = np.asarray(rois, np.float32)
    s = _f(scale)
    batch = rois[:, 0].astype(np.int32)                      # trunc toward zero (:25)
    sw, sh = rois[:, 1] * s, rois[:, 2] * s
    ew, eh = rois[:, 3] * s, rois[:, 4] * s
    rw = np.maximum(ew - sw, _f(1))
    rh = np.maximum(eh - sh, _f(1))
    bh = rh / _f(pooled[0])
    bw = rw / _f(pooled[1])
    if sr > 0:
        gh = np.full(len(rois), sr, np.int32)
        gw = np.full(len(rois), sr, np.int32)
    else:
        gh = np.ceil(rh / _f(pooled[0])).astype(np.int32)
        gw = np.ceil(rw / _f(pooled[1])).astype(np.int32)
    return dict(batch=batch, start_w=sw, start_h=sh, bin_w=bw, bin_h=bh,
                grid_h=gh, grid_w=gw, count=(gh * gw).astype(np.float32))


def numpy_axis(start, binsz, pooled, grid, extent):
    """All sample taps of ONE roi along one axis -> (pos, lo, hi, frac, valid), shape [pooled, grid]."""
    p = np.arange(pooled, dtype=np.float32)[:, None]
    i = np.arange(grid, dtype=np.float32)[None, :]
    pos = (_f(start) + p * _f(binsz)) + ((i + _f(0.5)) * _f(binsz)) / _f(grid)
    pos = pos.astype(np.float32)
    valid = ~((pos < _f(-1)) | (pos > _f(extent)))
    q = np.where(pos <= 0, _f(0), pos).astype(np.float32)
    lo = q.astype(np.int32)
    edge = lo >= extent - 1
    lo = np.where(edge, extent - 1, lo)
    hi = np.where(edge, extent - 1, lo + 1)
    q = np.where(edge, lo.astype(np.float32), q)
    frac = (q - lo.astype(np.float32)).astype(np.float32)
    return pos, lo, hi, frac, valid


def numpy_forward(data, rois, pooled, scale, sr):
    """fp32 forward with the reference's per-output summation order (iy, then ix)."""
    _check(data, rois, pooled)
    N, C, H, W = data.shape
    PH, PW = pooled
    g = numpy_geometry(rois, pooled, scale, sr)
    out = np.zeros((len(rois), C, PH, PW), np.float32)
    for r in range(len(rois)):
        plane = data[g['batch'][r]]                              # [C, H, W]
        _, ylo, yhi, ly, yv = numpy_axis(g['start_h'][r], g['bin_h'][r], PH, g['grid_h'][r], H)
        _, xlo, xhi, lx, xv = numpy_axis(g['start_w'][r], g['bin_w'][r], PW, g['grid_w'][r], W)
        hy, hx = _f(1) - ly, _f(1) - lx
        acc = np.zeros((C, PH, PW), np.float32)
        for iy in range(g['grid_h'][r]):
            for ix in range(g['grid_w'][r]):
                yl, yh = ylo[:, iy][:, None], yhi[:, iy][:, None]
                xl, xh = xlo[:, ix][None, :], xhi[:, ix][None, :]
                w1 = (hy[:, iy][:, None] * hx[:, ix][None, :]).astype(np.float32)
                w2 = (hy[:, iy][:, None] * lx[:, ix][None, :]).astype(np.float32)
                w3 = (ly[:, iy][:, None] * hx[:, ix][None, :]).astype(np.float32)
                w4 = (ly[:, iy][:, None] * lx[:, ix][None, :]).astype(np.float32)
                val = (w1 * plane[:, yl, xl] + w2 * plane[:, yl, xh]
                       + w3 * plane[:, yh, xl] + w4 * plane[:, yh, xh]).astype(np.float32)
                ok = (yv[:, iy][:, None] & xv[:, ix][None, :])
                acc = (acc + np.where(ok[None], val, _f(0))).astype(np.float32)
        out[r] = acc / g['count'][r]
    return out


def numpy_backward(dy, rois, data_shape, scale, sr):
    """fp32 backward in the single-thread canonical order (slow: small cases only)."""
    R, C, PH, PW = dy.shape
    N, _, H, W = data_shape
    g = numpy_geometry(rois, (PH, PW), scale, sr)
    dx = np.zeros(data_shape, np.float32)
    for r in range(R):
        b = g['batch'][r]
        _, ylo, yhi, ly, yv = numpy_axis(g['start_h'][r], g['bin_h'][r], PH, g['grid_h'][r], H)
        _, xlo, xhi, lx, xv = numpy_axis(g['start_w'][r], g['bin_w'][r], PW, g['grid_w'][r], W)
        hy, hx = _f(1) - ly, _f(1) - lx
        cnt = g['count'][r]
        for c in range(C):
            plane = dx[b, c]
            for ph in range(PH):
                for pw in range(PW):
                    d = dy[r, c, ph, pw]
                    for iy in range(g['grid_h'][r]):
                        for ix in range(g['grid_w'][r]):
                            if not (yv[ph, iy] and xv[pw, ix]):
                                continue
                            w = (hy[ph, iy] * hx[pw, ix], hy[ph, iy] * lx[pw, ix],
                                 ly[ph, iy] * hx[pw, ix], ly[ph, iy] * lx[pw, ix])
                            pix = ((ylo[ph, iy], xlo[pw, ix]), (ylo[ph, iy], xhi[pw, ix]),
                                   (yhi[ph, iy], xlo[pw, ix]), (yhi[ph, iy], xhi[pw, ix]))
                            for wk, (yy, xx) in zip(w, pix):
                                plane[yy, xx] = _f(plane[yy, xx] + _f(_f(d * _f(wk)) / cnt))
    return dx


def unique_elements_touched(rois, shape, pooled, scale, sr):
    """U of SURVEY.md 8(d): number of distinct feature-map elements the forward reads."""
    N, C, H, W = shape
    PH, PW = pooled
    g = numpy_geometry(rois, pooled, scale, sr)
    touched = np.zeros((N, H, W), bool)
    for r in range(len(rois)):
        b = g['batch'][r]
        if not 0 <= b < N:
            continue
        _, ylo, yhi, _, yv = numpy_axis(g['start_h'][r], g['bin_h'][r], PH, g['grid_h'][r], H)
        _, xlo, xhi, _, xv = numpy_axis(g['start_w'][r], g['bin_w'][r], PW, g['grid_w'][r], W)
        ys = np.unique(np.concatenate([ylo[yv], yhi[yv]]))
        xs = np.unique(np.concatenate([xlo[xv], xhi[xv]]))
        if len(ys) and len(xs):
            # a tap (y, x) is read iff its row sample and column sample are both valid;
            # rows/cols combine as an outer product because the grid is separable
            rowmask = np.zeros(H, bool)
            colmask = np.zeros(W, bool)
            rowmask[ys] = True
            colmask[xs] = True
            touched[b] |= rowmask[:, None] & colmask[None, :]
    return int(touched.sum()) * C
