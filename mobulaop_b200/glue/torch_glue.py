"""PyTorch adapter (reference: /root/reference/mobula/glue/torch_glue.py).

Same user-visible forms:
    y = mobula.op.ROIAlign(data, rois, pooled_size=(7, 7), spatial_scale=.25, sampling_ratio=2)
    layer = mobula.op.ROIAlign[torch.Tensor](pooled_size=..., ...)    # an nn.Module
Both are differentiable through torch.autograd.

Differences from the reference:
* every forward call gets its own operator record, so one layer can be applied several
  times before backward (the reference keeps the tensors on the module, torch_glue.py:58-76);
* gradients are written into fresh tensors -- the reference reuses `input.grad` as
  scratch and zero-fills it (torch_glue.py:83-84 with ROIAlign.py:33-34);
* inputs that do not require grad get req='null' and a None gradient;
* kernels run on torch's current CUDA stream, no synchronisation.
"""
import ctypes

import torch
from torch.autograd.function import once_differentiable

from .common import (MobulaTensor, TENSOR_VIEWS, assign, c_bfloat16, c_half, forward_input_names, shapes_of,
                     split_inputs)

F = torch

_CTYPES = {torch.int32: ctypes.c_int, torch.float32: ctypes.c_float, torch.float64: ctypes.c_double,
           torch.int64: ctypes.c_int64, torch.float16: c_half, torch.bfloat16: c_bfloat16}


class TorchTensor(MobulaTensor):
    F = torch

    @property
    def data_ptr(self):
        t = self.tensor
        if not t.is_contiguous():
            c = t.contiguous()
            return ctypes.c_void_p(c.data_ptr()), c
        return ctypes.c_void_p(t.data_ptr())

    @property
    def ctype(self):
        ctype = _CTYPES.get(self.tensor.dtype)
        if ctype is None:
            raise TypeError('Unknown Type: %s' % self.tensor.dtype)
        return ctype

    @property
    def dev_id(self):
        dev = self.tensor.device
        if dev.type == 'cpu':
            return None
        if dev.type != 'cuda':
            raise TypeError('unsupported torch device %s' % dev)
        return dev.index if dev.index is not None else torch.cuda.current_device()

    @property
    def stream(self):
        dev = self.tensor.device
        return torch.cuda.current_stream(dev).cuda_stream if dev.type == 'cuda' else 0


Tensor = TorchTensor


def _make_classes(op, name):
    n_inputs = len(forward_input_names(op))

    record_members = dict(_forward=op.forward, _backward=op.backward, assign=assign,
                          F=property(lambda self: torch))
    record_members.update(TENSOR_VIEWS)
    record_cls = type('_%s_TORCH_OP' % name, (op, object), record_members)

    class _Function(torch.autograd.Function):
        @staticmethod
        def forward(ctx, record, *inputs):
            record.in_data = inputs
            record.req = ['write'] * len(inputs)
            first = inputs[0] if inputs else None
            dtype = first.dtype if first is not None else torch.float32
            device = first.device if first is not None else torch.device('cpu')
            out_shapes = record.infer_shape(shapes_of(inputs))[1]
            # an operator may store its outputs in another element type (ROIAlign: out_dtype)
            odt = record.infer_out_dtype(dtype, torch) if hasattr(record, 'infer_out_dtype') else dtype
            record.out_data = [torch.empty(tuple(s), dtype=odt, device=device) for s in out_shapes]
            result = record._forward(*inputs)
            if result is not None:
                result = result if isinstance(result, (list, tuple)) else [result]
                for dst, value in zip(record.out_data, result):
                    record.assign(dst, 'write', value)
            outs = record.out_data
            # the inputs go through autograd's own storage (version-counter check, no
            # output -> grad_fn -> record -> output cycle); the record keeps scalars only
            ctx.save_for_backward(*inputs)
            ctx.record = record
            record.in_data = record.out_data = None
            return outs[0] if len(outs) == 1 else tuple(outs)

        @staticmethod
        @once_differentiable
        def backward(ctx, *out_grads):
            record = ctx.record
            inputs = ctx.saved_tensors
            needs = ctx.needs_input_grad[1:]
            record.in_data = inputs
            record.req = ['write' if n else 'null' for n in needs]
            record.in_grad = [torch.empty_like(d) if n else None
                              for d, n in zip(inputs, needs)]
            record.out_grad = [g.contiguous() if g is not None else None for g in out_grads]
            result = record._backward(*record.out_grad)
            if result is not None:
                result = result if isinstance(result, (list, tuple)) else [result]
                for i in range(n_inputs):
                    if record.in_grad[i] is not None:
                        record.assign(record.in_grad[i], record.req[i], result[i])
            grads = tuple(record.in_grad)
            record.in_data = record.in_grad = record.out_grad = None
            return (None,) + grads

    _Function.__name__ = '_%s_TORCH_FUNC' % name

    class _Module(torch.nn.Module):
        def __init__(self, *args, **kwargs):
            torch.nn.Module.__init__(self)
            self._ctor = (args, kwargs)
            record_cls(*args, **kwargs)       # validate the parameters now

        def forward(self, *inputs):
            args, kwargs = self._ctor
            return _Function.apply(record_cls(*args, **kwargs), *inputs)

        def extra_repr(self):
            args, kwargs = self._ctor
            parts = [repr(a) for a in args] + ['%s=%r' % kv for kv in kwargs.items()]
            return ', '.join(parts)

    _Module.__name__ = '_%s_TORCH_NN_MODULE' % name
    return record_cls, _Function, _Module


class OpGen(object):
    def __init__(self, op, name):
        self.op = op
        self.name = name
        self.classes = None

    def register(self):
        if self.classes is None:
            self.classes = _make_classes(self.op, self.name)
        return self.classes[2]

    def __call__(self, *args, **kwargs):
        module_cls = self.register()
        if '__input_type__' in kwargs:
            kwargs = dict(kwargs)
            kwargs.pop('__input_type__')
            return module_cls(*args, **kwargs)
        inputs, ctor_args, ctor_kwargs = split_inputs(self.op, args, kwargs)
        record_cls, function_cls, _ = self.classes
        return function_cls.apply(record_cls(*ctor_args, **ctor_kwargs), *inputs)
