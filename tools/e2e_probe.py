"""Where the end-to-end step goes: forward and backward through the NumPy operator, and the
same two calls straight through the C-ABI (no Python glue), page-locked host buffers, C2."""
import ctypes
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import numpy as np
import torch

import mobula
from mobulaop_b200 import _native, workloads as WL

lib = _native.load()
mobula.op.load('ROIAlign')
wl = WL.CONFIGS['C2']


def pinned(shape):
    n = int(np.prod(shape))
    ptr = lib.mobula_host_alloc(n * 4)
    buf = (ctypes.c_float * n).from_address(ptr)
    return np.frombuffer(buf, dtype=np.float32).reshape(shape)


rng = np.random.default_rng(1)
data, dy = pinned(wl.data_shape), pinned(wl.out_shape)
data[...] = rng.standard_normal(wl.data_shape, dtype=np.float32)
dy[...] = rng.standard_normal(wl.out_shape, dtype=np.float32)
rois = WL.make_rois(wl)
op = mobula.op.ROIAlign[np.ndarray](pooled_size=wl.pooled, spatial_scale=wl.scale, sampling_ratio=2)


def clock(fn, reps=8):
    fn()
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps * 1e3


print('op forward   %.3f ms' % clock(lambda: op(data, rois)))
op(data, rois)
print('op backward  %.3f ms' % clock(lambda: op.backward(dy)))

out, dx = pinned(wl.out_shape), pinned(wl.data_shape)
vp = lambda a: ctypes.c_void_p(a.ctypes.data)
N, C, H, W = wl.data_shape
PH, PW = wl.pooled
flags = _native.make_flags(host=True)
fwd = lambda: lib.mobula_roialign_forward_f32(0, None, vp(data), vp(rois), vp(out), wl.R, N, C, H, W, PH, PW,
                                              wl.scale, 2, flags)
bwd = lambda: lib.mobula_roialign_backward_f32(0, None, vp(dy), vp(rois), vp(dx), wl.R, N, C, H, W, PH, PW,
                                               wl.scale, 2, _native.BWD_ATOMIC, flags)
print('abi forward  %.3f ms' % clock(fwd))
print('abi backward %.3f ms' % clock(bwd))
gb = data.nbytes / 1e9
print('floor at 55 GB/s one way: %.3f ms per direction-bound call' % (gb / 55 * 1e3))
