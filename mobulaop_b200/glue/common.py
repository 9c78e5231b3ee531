"""Pieces shared by the framework adapters.

Mirrors the contract of /root/reference/mobula/glue/common.py: `register` turns an
operator class into a `MobulaOperator` published as `mobula.op.<Name>`
(common.py:205-226), `MobulaOperator.__call__` picks the adapter from the argument
types (:177-182) and `__getitem__` from an explicit type (:184-194);
operator instances see their tensors through X/Y/dX/dY/x/y/dx/dy (:229-238).
"""
import ctypes
import functools
import inspect
import warnings

CUSTOM_OP_LIST = {}
OP_MODULE_GLOBALS = None     # set by mobulaop_b200.op to its module dict
backend = None               # set by glue/__init__.py


_FORWARD_SPECS = {}


def _forward_spec(op):
    """(input names, number of trailing inputs with a default) of `op.forward`, looked up once per
    operator class: inspect.getfullargspec costs tens of microseconds, every call came through it."""
    spec = _FORWARD_SPECS.get(op)
    if spec is None:
        full = inspect.getfullargspec(op.forward)
        spec = _FORWARD_SPECS[op] = (full.args[1:], len(full.defaults or ()))
    return spec


def forward_input_names(op):
    """Names of the tensor inputs = parameters of `op.forward` after self."""
    return _forward_spec(op)[0]


def split_inputs(op, args, kwargs):
    """Separates the tensor inputs of `op.forward` from constructor parameters.

    Returns (inputs, ctor_args, ctor_kwargs).  Inputs may be given positionally or by
    name; trailing inputs with defaults may be omitted (reference: common.py:68-105).
    """
    names, n_default = _forward_spec(op)
    kwargs = dict(kwargs)
    if len(args) >= len(names):
        return list(args[:len(names)]), list(args[len(names):]), kwargs
    inputs = list(args)
    for pos in range(len(args), len(names)):
        name = names[pos]
        if name in kwargs:
            inputs.append(kwargs.pop(name))
        elif pos >= len(names) - n_default:
            break
        else:
            raise AssertionError('Variable %s not found' % name)
    for name in names[:len(args)]:
        assert name not in kwargs, 'input %s given twice' % name
    return inputs, [], kwargs


def shapes_of(tensors):
    return [tuple(t.shape) for t in tensors]


def assign(_self, dst, request, src):
    """Honours a write request when an op returns its result instead of writing it."""
    if request == 'null':
        return
    if request in ('write', 'inplace'):
        dst[:] = src
    elif request == 'add':
        dst[:] += src
    else:
        raise ValueError('unknown req %r' % (request,))


class MobulaTensor(object):
    """Adapter view of one framework tensor: raw pointer, element ctype, device."""
    F = None

    def __init__(self, tensor):
        self.tensor = tensor

    @property
    def data_ptr(self):
        """c_void_p, or (c_void_p, contiguous_copy) when a copy had to be made."""
        raise NotImplementedError

    @property
    def ctype(self):
        raise NotImplementedError

    @property
    def dev_id(self):
        """GPU ordinal, or None for host memory."""
        raise NotImplementedError

    @property
    def stream(self):
        """Raw CUDA stream the framework is currently issuing work on (0 = default)."""
        return 0

    @property
    def shape(self):
        return tuple(self.tensor.shape)


class MobulaOperator(object):
    def __init__(self, op, name, **attrs):
        self.op = op
        self.name = name
        self.attrs = attrs

    def _generator(self, glue_mod):
        return backend.op_gen(glue_mod, op=self.op, name=self.name)

    def __call__(self, *args, **kwargs):
        glue_mod = backend.get_args_glue(*args, **kwargs)
        if glue_mod is None:
            raise ValueError('No explicit backend: pass tensors, or use %s[tensor_type](...)'
                             % self.name)
        merged = dict(kwargs)
        merged.update(self.attrs)
        return self._generator(glue_mod)(*args, **merged)

    def __getitem__(self, input_type):
        glue_mod = backend.get_var_type_glue(input_type)
        if glue_mod is None:
            raise ValueError('The backend of %s is not found' % (input_type,))
        gen = self._generator(glue_mod)

        def make(*args, **kwargs):
            merged = dict(kwargs)
            merged.update(self.attrs)
            merged['__input_type__'] = input_type
            return gen(*args, **merged)
        return make


def register(op_name=None, **attrs):
    """@register / @register('Name') / @register(attr=...) on an operator class."""
    def decorate(name, op):
        name = name or op.__name__
        inst = MobulaOperator(op=op, name=name, **attrs)
        if name in CUSTOM_OP_LIST:
            warnings.warn('Duplicate operator name %s, please rename it' % name)
        CUSTOM_OP_LIST[name] = inst
        if OP_MODULE_GLOBALS is not None:
            OP_MODULE_GLOBALS[name] = inst
        return inst
    if op_name is not None and not isinstance(op_name, str):
        return decorate(None, op_name)
    return functools.partial(decorate, op_name)


def _first(seq):
    return seq[0]


# properties every generated operator class carries
TENSOR_VIEWS = dict(
    X=property(lambda self: self.in_data), Y=property(lambda self: self.out_data),
    dX=property(lambda self: self.in_grad), dY=property(lambda self: self.out_grad),
    x=property(lambda self: _first(self.in_data)), y=property(lambda self: _first(self.out_data)),
    dx=property(lambda self: _first(self.in_grad)), dy=property(lambda self: _first(self.out_grad)),
)

class c_half(ctypes.c_uint16):
    """IEEE binary16 storage (no ctypes equivalent); only the pooled tensor of ROIAlign may have it."""


class c_bfloat16(ctypes.c_uint16):
    """bfloat16 storage."""


_NP_NAME_TO_CTYPE = {
    'float16': c_half,
    'int8': ctypes.c_int8, 'int16': ctypes.c_int16, 'int32': ctypes.c_int32,
    'int64': ctypes.c_int64, 'float32': ctypes.c_float, 'float64': ctypes.c_double,
}


def ctype_of_numpy_dtype(dtype):
    import numpy as np
    ctype = _NP_NAME_TO_CTYPE.get(np.dtype(dtype).name)
    if ctype is None:
        raise TypeError('Unknown Type: %s' % (dtype,))
    return ctype
