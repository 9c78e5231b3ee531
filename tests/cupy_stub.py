"""A stand-in for the `cupy` package (absent from this image) that exposes just
what mobulaop_b200/glue/cupy_glue.py and the NumPy-semantics operator class touch:

    cupy.ndarray        .data.ptr  .device.id  .dtype  .flags.c_contiguous  .shape  .size
                        a[:] = b   a[:] += b   .get()
    cupy.empty / ones / zeros / ascontiguousarray / asarray
    cupy.cuda.get_current_stream().ptr

Arrays are backed by torch CUDA tensors when a GPU is there (the pointers are real device
pointers and the kernels run on them), by host torch tensors otherwise (attribute mapping
only -- nothing may be launched on those).

    import cupy_stub; cupy = cupy_stub.install()      # sys.modules['cupy'] = the stub
"""
import sys
import types

import numpy as np
import torch

_NP_TO_TORCH = {np.dtype('float32'): torch.float32, np.dtype('float64'): torch.float64,
                np.dtype('int32'): torch.int32, np.dtype('int64'): torch.int64}
_TORCH_TO_NP = {v: k for k, v in _NP_TO_TORCH.items()}


def _device():
    return torch.device('cuda', torch.cuda.current_device()) if torch.cuda.is_available() \
        else torch.device('cpu')


class _Mem(object):
    def __init__(self, t):
        self.ptr = int(t.data_ptr())


class _Dev(object):
    def __init__(self, t):
        self.id = t.device.index if t.device.type == 'cuda' else 0


class _Flags(object):
    def __init__(self, t):
        self.c_contiguous = bool(t.is_contiguous())


class ndarray(object):
    def __init__(self, tensor):
        self._t = tensor

    data = property(lambda self: _Mem(self._t))
    device = property(lambda self: _Dev(self._t))
    flags = property(lambda self: _Flags(self._t))
    dtype = property(lambda self: _TORCH_TO_NP[self._t.dtype])
    shape = property(lambda self: tuple(self._t.shape))
    size = property(lambda self: int(self._t.numel()))
    ndim = property(lambda self: self._t.dim())

    def get(self):
        return self._t.detach().cpu().numpy()

    def __getitem__(self, key):
        return ndarray(self._t[key])

    def __setitem__(self, key, value):
        self._t[key] = value._t if isinstance(value, ndarray) else value

    def __iadd__(self, other):
        self._t += other._t if isinstance(other, ndarray) else other
        return self

    def __len__(self):
        return self._t.shape[0]


ndarray.__module__ = 'cupy'          # backend.get_var_type_glue keys on the package name


def _torch_dtype(dtype):
    return _NP_TO_TORCH[np.dtype(dtype)]


def empty(shape, dtype=np.float32):
    return ndarray(torch.empty(tuple(shape), dtype=_torch_dtype(dtype), device=_device()))


def zeros(shape, dtype=np.float32):
    return ndarray(torch.zeros(tuple(shape), dtype=_torch_dtype(dtype), device=_device()))


def ones(shape, dtype=np.float32):
    return ndarray(torch.ones(tuple(shape), dtype=_torch_dtype(dtype), device=_device()))


def asarray(a):
    if isinstance(a, ndarray):
        return a
    return ndarray(torch.from_numpy(np.ascontiguousarray(a)).to(_device()))


def ascontiguousarray(a):
    return ndarray(a._t.contiguous())


class _Stream(object):
    @property
    def ptr(self):
        return int(torch.cuda.current_stream().cuda_stream) if torch.cuda.is_available() else 0


def install():
    mod = types.ModuleType('cupy')
    mod.ndarray = ndarray
    mod.empty, mod.zeros, mod.ones = empty, zeros, ones
    mod.asarray, mod.ascontiguousarray = asarray, ascontiguousarray
    mod.float32, mod.float64 = np.float32, np.float64
    mod.cuda = types.ModuleType('cupy.cuda')
    mod.cuda.get_current_stream = lambda: _Stream()
    mod.__stub__ = True
    sys.modules['cupy'] = mod
    sys.modules['cupy.cuda'] = mod.cuda
    return mod
