"""Comparison helpers for the parity tests.

`assert_almost_equal` keeps the reference's meaning (/root/reference/mobula/testing.py:33-115):
absolute error against `atol`, then relative error |a-b| / (|b| + eps) against `rtol`.
`ulp_distance` / `assert_within_ulp_or_rel` express the stricter gates of this build:
forward within 2 ULP or 1e-6 relative, backward within 1e-5 relative.
"""
import os

import numpy as np


def to_numpy(data):
    if isinstance(data, np.ndarray):
        return data
    if hasattr(data, 'detach'):
        return data.detach().cpu().numpy()
    if hasattr(data, 'get'):            # cupy
        return data.get()
    if isinstance(data, (list, tuple, int, float)):
        return np.asarray(data)
    raise TypeError('Unsupported Type: %s' % type(data))


def _where(err, a, b):
    idx = np.unravel_index(int(np.argmax(err)), err.shape)
    return 'at %s: %r vs %r' % (idx, a[idx], b[idx])


def assert_almost_equal(a, b, rtol=1e-5, atol=1e-8):
    a, b = to_numpy(a), to_numpy(b)
    if a.shape != b.shape:
        a, b = np.broadcast_arrays(a, b)
    a = np.atleast_1d(a)
    b = np.atleast_1d(b)
    if a.size == 0:
        return
    abs_err = np.abs(a.astype(np.float64) - b.astype(np.float64))
    if atol is not None and abs_err.max() > atol:
        raise AssertionError('Maximum Absolute Error(%g) > atol(%g) %s'
                             % (abs_err.max(), atol, _where(abs_err, a, b)))
    try:
        eps = np.finfo(b.dtype).eps
    except ValueError:
        eps = np.finfo(np.float32).eps
    rel_err = abs_err / (np.abs(b).astype(np.float64) + eps)
    if rel_err.max() > rtol:
        raise AssertionError('Maximum Relative Error(%g) > rtol(%g) %s'
                             % (rel_err.max(), rtol, _where(rel_err, a, b)))


def ulp_distance(a, b):
    """Element-wise distance in units in the last place between two float32 arrays."""
    a = np.ascontiguousarray(to_numpy(a), np.float32)
    b = np.ascontiguousarray(to_numpy(b), np.float32)
    ia = a.view(np.int32).astype(np.int64)
    ib = b.view(np.int32).astype(np.int64)
    ia = np.where(ia < 0, np.int64(-2 ** 31) - ia, ia)     # map to a monotonic integer line
    ib = np.where(ib < 0, np.int64(-2 ** 31) - ib, ib)
    return np.abs(ia - ib)


def assert_within_ulp_or_rel(a, b, ulp=2, rtol=1e-6):
    """Every element within `ulp` units in the last place OR `rtol` relative error."""
    a, b = to_numpy(a), to_numpy(b)
    assert a.shape == b.shape, 'Unmatched Shape: %s vs %s' % (a.shape, b.shape)
    if a.size == 0:
        return
    d = ulp_distance(a, b)
    rel = np.abs(a.astype(np.float64) - b.astype(np.float64)) / np.maximum(np.abs(b), 1e-300)
    bad = (d > ulp) & (rel > rtol)
    if bad.any():
        idx = np.unravel_index(int(np.argmax(np.where(bad, d, 0))), d.shape)
        raise AssertionError('%d elements differ by more than %d ULP and %g relative; worst at %s: '
                             '%r vs %r (%d ULP)' % (bad.sum(), ulp, rtol, idx, a[idx], b[idx], d[idx]))


def assert_bit_exact(a, b):
    a, b = np.ascontiguousarray(to_numpy(a)), np.ascontiguousarray(to_numpy(b))
    assert a.shape == b.shape and a.dtype == b.dtype, (a.shape, b.shape, a.dtype, b.dtype)
    same = a.view(np.uint8).reshape(-1) == b.view(np.uint8).reshape(-1) if a.size else np.ones(0, bool)
    if not same.all():
        diff = a != b
        n = int(diff.sum())
        idx = np.unravel_index(int(np.argmax(diff)), a.shape) if n else None
        raise AssertionError('%d elements differ bitwise; first at %s: %r vs %r'
                             % (n, idx, a[idx] if n else None, b[idx] if n else None))


def assert_scatter_close(a, b, abs_sum, rtol=1e-5):
    """Backward gate for the non-deterministic mode: a scatter-add result may differ from
    the canonical-order result only by re-association of its summands, so the error of
    each element is measured relative to the sum of the ABSOLUTE values of the summands
    that landed on it (`abs_sum` = the same backward applied to |dy|), which is the
    quantity floating-point re-ordering error scales with.  For an element without
    cancellation this is the plain relative error."""
    a, b, s = to_numpy(a), to_numpy(b), to_numpy(abs_sum)
    assert a.shape == b.shape == s.shape, (a.shape, b.shape, s.shape)
    if a.size == 0:
        return
    err = np.abs(a.astype(np.float64) - b.astype(np.float64))
    bound = rtol * s.astype(np.float64)
    bad = err > bound
    if bad.any():
        ratio = np.where(s > 0, err / np.maximum(s, 1e-300), np.where(err > 0, np.inf, 0))
        idx = np.unravel_index(int(np.argmax(ratio)), ratio.shape)
        raise AssertionError('%d elements exceed %g x sum|summands|; worst at %s: %r vs %r '
                             '(error %g, sum|summands| %g)'
                             % (bad.sum(), rtol, idx, a[idx], b[idx], err[idx], s[idx]))


def cuda_ready(timeout_s=45.0):
    """True when a CUDA device is usable.  On a box that has just been started the first driver
    initialisation can fail ("CUDA driver initialization failed") and succeed seconds later; a failed
    first attempt sticks to the process that made it, so the driver is probed in short-lived child
    processes (cuInit through ctypes, ~50 ms each) until it answers or `timeout_s` has passed, and only
    then is torch asked.  Without a driver library (the CPU build container) this returns at once."""
    import ctypes.util
    import subprocess
    import sys
    import time
    import glob
    if not glob.glob('/dev/nvidia[0-9]*') and not os.environ.get('MOBULA_ASSUME_GPU'):
        return False                    # no device node: nothing to wait for
    if ctypes.util.find_library('cuda') is None:
        try:
            ctypes.CDLL('libcuda.so.1')
        except OSError:
            return False
    probe = ("import ctypes, sys\n"
             "lib = ctypes.CDLL('libcuda.so.1')\n"
             "n = ctypes.c_int(0)\n"
             "rc = lib.cuInit(0) or lib.cuDeviceGetCount(ctypes.byref(n))\n"
             "sys.exit(0 if rc == 0 and n.value > 0 else (3 if rc == 0 else 4))\n")
    deadline = time.time() + timeout_s
    while True:
        try:
            rc = subprocess.run([sys.executable, '-c', probe], timeout=20).returncode
        except Exception:
            rc = 4
        if rc == 0:
            break
        if rc == 3 or time.time() >= deadline:      # initialised, but no device / gave up
            return False
        time.sleep(1.0)
    try:
        import torch
        return bool(torch.cuda.is_available())
    except Exception:
        return False
