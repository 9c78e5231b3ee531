"""NumPy adapter (reference: /root/reference/mobula/glue/numpy_glue.py).

NumPy arrays live in host memory; the launch path hands their pointers to the library
with MOBULA_ROIALIGN_HOST_BUFFERS, which stages them through the current GPU.

Two call forms, as in the reference:
    op = mobula.op.ROIAlign[np.ndarray](pooled_size=..., ...)   # instance
    y = op(data, rois); dx, drois = op.backward(dy)
and the plain forward-only form `mobula.op.ROIAlign(data, rois, pooled_size=...)`, which
raises TypeError in the reference (numpy_glue.py:39-41 builds the op without its
constructor arguments) and works here.
"""
import ctypes

import numpy as np

from .common import (MobulaTensor, TENSOR_VIEWS, assign, ctype_of_numpy_dtype, forward_input_names,
                     shapes_of, split_inputs)

F = np


class NumPyTensor(MobulaTensor):
    F = np

    @property
    def data_ptr(self):
        t = self.tensor
        if not t.flags.c_contiguous:
            c = np.ascontiguousarray(t)
            return ctypes.c_void_p(c.ctypes.data), c
        return ctypes.c_void_p(t.ctypes.data)

    @property
    def ctype(self):
        return ctype_of_numpy_dtype(self.tensor.dtype)

    @property
    def dev_id(self):
        return None


Tensor = NumPyTensor


def make_array_op_class(op, name, xp, suffix):
    """Operator class for array libraries with NumPy semantics (shared with CuPy)."""
    if xp is np:
        from . import pinned
        new_empty = pinned.empty          # results land in page-locked memory
    else:
        def new_empty(shape, dtype):
            return xp.empty(shape, dtype=dtype)

    def forward(self, *args, **kwargs):
        inputs, extra, extra_kw = split_inputs(op, args, kwargs)
        assert not extra and not extra_kw, 'unexpected arguments %s %s' % (extra, extra_kw)
        self.in_data = list(inputs)
        dtype = self.in_data[0].dtype
        self.req = ['write'] * len(self.in_data)
        out_shapes = self.infer_shape(shapes_of(self.in_data))[1]
        odt = self.infer_out_dtype(dtype, xp) if hasattr(self, 'infer_out_dtype') else dtype
        self.out_data = [new_empty(tuple(s), odt) for s in out_shapes]
        result = self._forward(*inputs)
        if result is not None:
            result = result if isinstance(result, (list, tuple)) else [result]
            for dst, value in zip(self.out_data, result):
                self.assign(dst, 'write', value)
        return self.out_data[0] if len(self.out_data) == 1 else self.out_data

    def backward(self, out_grad=None, in_data=None, out_data=None, in_grad=None, req=None):
        if in_data is not None:
            self.in_data = in_data
        if out_data is not None:
            self.out_data = out_data
        dtype = self.in_data[0].dtype
        if in_grad is None:
            in_grad = [new_empty(d.shape, dtype) for d in self.in_data]
        elif not isinstance(in_grad, (list, tuple)):
            in_grad = [in_grad]
        self.in_grad = in_grad
        if out_grad is None:
            out_grad = [xp.ones(d.shape, dtype=d.dtype) for d in self.out_data]
        elif not isinstance(out_grad, (list, tuple)):
            out_grad = [out_grad]
        self.out_grad = out_grad
        if req is None:
            self.req = ['write'] * len(self.in_data)
        else:
            if len(req) != len(self.in_data):
                raise ValueError('len(req) should be %d' % len(self.in_data))
            self.req = list(req)
        result = self._backward(*out_grad)
        if result is not None:
            result = result if isinstance(result, (list, tuple)) else [result]
            for i in range(len(forward_input_names(op))):
                self.assign(in_grad[i], self.req[i], result[i])
        return in_grad[0] if len(in_grad) == 1 else self.in_grad

    members = dict(
        __call__=forward, forward=forward, backward=backward,
        _forward=op.forward, _backward=op.backward, infer_shape=op.infer_shape,
        assign=assign, F=property(lambda self: xp), op=property(lambda self: op),
    )
    members.update(TENSOR_VIEWS)
    return type('_%s_%s_OP' % (name, suffix), (op, object), members)


class OpGen(object):
    array_module = np
    suffix = 'NP'

    def __init__(self, op, name):
        self.op = op
        self.name = name
        self.op_class = None

    def __call__(self, *args, **kwargs):
        if self.op_class is None:
            self.op_class = self.register()
        if '__input_type__' in kwargs:
            kwargs = dict(kwargs)
            kwargs.pop('__input_type__')
            return self.op_class(*args, **kwargs)
        # plain call: tensors + constructor parameters in one go, forward only
        inputs, ctor_args, ctor_kwargs = split_inputs(self.op, args, kwargs)
        return self.op_class(*ctor_args, **ctor_kwargs)(*inputs)

    def register(self):
        return make_array_op_class(self.op, self.name, self.array_module, self.suffix)
