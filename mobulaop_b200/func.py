"""`mobula.func.<kernel>` launch path for the prebuilt kernels (ROIAlign; im2col / col2im).

Keeps the calling convention of the reference's dispatcher
(/root/reference/mobula/func.py:190-265 `MobulaFunc.__call__`, :138-164
`CFuncDef.__call__`): positional/keyword arguments in the order of the C++ kernel
signature (/root/reference/opzoo/ROIAlign/ROIAlign.cpp:9-16, 78-83), template type T
bound from the first tensor and enforced on the others, every tensor on one device,
non-contiguous tensors copied and (if written) copied back.

What changed: there is no JIT, no per-context loader and no CPU/HIP/TVM branch -- both
kernels resolve to libmobula_roialign_sm100.so (include/mobula_roialign.h).  Device
tensors are launched on the framework's current CUDA stream without a synchronise;
host tensors (NumPy, CPU torch) are staged through the GPU by the library and the call
returns when the result is in host memory.

Extra keyword-only arguments (absent from the reference): `mode` ('atomic' |
'deterministic', backward only), `accumulate` (default True for backward = the
reference semantics "adds into a caller-zeroed buffer", False for forward), `algo`,
`aligned` (half-pixel box coordinates = torchvision.ops.roi_align(aligned=True)), `layout`
('RCHW' = the reference's [R, C, PH, PW], 'RHWC' = channel-last [R, PH, PW, C]); the pooled tensor
(`top_data` / `top_diff`) may be float16 or bfloat16 -- the arithmetic stays fp32.
"""
import ctypes
import sys

from . import _native
from . import glue
from .config import config

__all__ = ['MobulaFunc', 'bind']

_INT, _SCALAR_T, _IN_T, _OUT_T = 'int', 'T', 'const T*', 'T*'

_COMMON = [('spatial_scale', _SCALAR_T), ('channels', _INT), ('height', _INT), ('width', _INT),
           ('pooled_height', _INT), ('pooled_width', _INT), ('sampling_ratio', _INT)]

# name -> (parameters in reference order, launcher)
_SIGNATURES = {
    'roi_align_forward': [('nthreads', _INT), ('bottom_data', _IN_T)] + _COMMON +
                         [('bottom_rois', _IN_T), ('top_data', _OUT_T)],
    'roi_align_backward': [('nthreads', _INT), ('top_diff', _IN_T)] + _COMMON +
                          [('bottom_diff', _OUT_T), ('bottom_rois', _IN_T)],
}

# im2col_kernel / col2im_kernel (/root/reference/opzoo/Convolution/Convolution.cpp:14-24, 48-56)
_CONV_GEOM = [(n, _INT) for n in ('height', 'width', 'kernel_h', 'kernel_w', 'pad_h', 'pad_w', 'stride_h',
                                  'stride_w', 'dilation_h', 'dilation_w', 'height_col', 'width_col')]
_SIGNATURES['im2col'] = [('n', _INT), ('data_im', _IN_T)] + _CONV_GEOM + [('data_col', _OUT_T)]
_SIGNATURES['col2im'] = [('n', _INT), ('data_col', _IN_T)] + _CONV_GEOM + [('data_im', _OUT_T)]

_SUPPORTED_T = (ctypes.c_float,)
# the pooled tensor of ROIAlign may be stored in half precision (include/mobula_roialign.h, `_ex` entry points)
_TOP_NAMES = ('top_data', 'top_diff')
_TOP_DTYPES = {ctypes.c_float: _native.DTYPE_F32, glue.common.c_half: _native.DTYPE_F16,
               glue.common.c_bfloat16: _native.DTYPE_BF16}


class _TensorArg(object):
    __slots__ = ('var', 'view', 'ptr', 'keep', 'writeback')

    def __init__(self, var, view):
        self.var = var
        self.view = view
        self.ptr = None
        self.keep = None
        self.writeback = False

    def resolve(self, mutable):
        p = self.view.data_ptr
        if isinstance(p, (tuple, list)):
            p, self.keep = p
            self.writeback = mutable
        self.ptr = p
        return p

    def finish(self):
        if self.writeback:
            self.var[:] = self.keep


class MobulaFunc(object):
    """One native kernel, callable with framework tensors."""

    def __init__(self, name, params):
        self.name = name
        self.params = list(params)
        self.arg_names = [p[0] for p in self.params]
        self._conv = name in ('im2col', 'col2im')
        self._launcher = None if self._conv else (_launch_forward if name == 'roi_align_forward' else _launch_backward)
        self._fast = None if self._conv else (_fast_torch_forward if name == 'roi_align_forward' else _fast_torch_backward)

    def __repr__(self):
        return '<mobula.func.%s(%s)>' % (self.name, ', '.join('%s %s' % (k, n) for n, k in self.params))

    def _collect(self, args, kwargs):
        args = list(args)
        if len(args) > len(self.params):
            raise TypeError('%s takes %d arguments, %d given' % (self.name, len(self.params), len(args)))
        for name in self.arg_names[len(args):]:
            if name not in kwargs:
                raise TypeError('%s: missing argument `%s`' % (self.name, name))
            args.append(kwargs.pop(name))
        return args

    def __call__(self, *args, **kwargs):
        # (this is the per-call dispatch path: plain loops and locals, no helper objects)
        if kwargs:
            mode = kwargs.pop('mode', None)
            accumulate = kwargs.pop('accumulate', None)
            algo = kwargs.pop('algo', None)
            aligned = bool(kwargs.pop('aligned', False))
            layout = kwargs.pop('layout', None)
        else:
            mode = accumulate = algo = layout = None
            aligned = False
        if kwargs or len(args) != len(self.params):
            args = self._collect(args, kwargs)
            if kwargs:
                raise TypeError('%s: unexpected arguments %s' % (self.name, sorted(kwargs)))
        if self._fast is not None and self._fast(args, mode, accumulate, algo, aligned, layout):
            return None

        values = {}
        tensors = {}
        bound_t = None
        dev_id = None
        dev_view = None
        n_host = 0
        type_glue = glue.backend.get_var_type_glue
        for (name, kind), var in zip(self.params, args):
            if kind is _INT:
                values[name] = int(var)
            elif kind is _SCALAR_T:
                values[name] = float(var.value) if hasattr(var, '_type_') else float(var)
            else:
                glue_mod = type_glue(type(var))
                if glue_mod is None:
                    raise TypeError('Unmatched parameters list of the function `%s`: `%s` must be a '
                                    'tensor, got %s' % (self.name, name, type(var)))
                view = glue_mod.Tensor(var)
                ctype = view.ctype
                if ctype is not bound_t:
                    if name in _TOP_NAMES and ctype in _TOP_DTYPES:
                        values['__top_dtype__'] = _TOP_DTYPES[ctype]      # fp32 arithmetic, typed storage
                    elif bound_t is None:
                        bound_t = ctype
                    else:
                        raise TypeError('Expected Type %s instead of %s for `%s`' % (bound_t, ctype, name))
                tdev = view.dev_id
                if tdev is None:
                    n_host += 1
                elif dev_id is None:
                    dev_id = tdev
                    dev_view = view
                elif tdev != dev_id:
                    raise ValueError("Don't use multiple devices in a call :-(")
                tensors[name] = _TensorArg(var, view)
        if bound_t not in _SUPPORTED_T:
            raise TypeError('%s is instantiated for float only (no %s kernel in '
                            'libmobula_roialign_sm100)' % (self.name, bound_t))
        host = dev_view is None
        if not host and n_host:
            raise ValueError('%s: host and device tensors mixed in one call' % self.name)
        stream = 0 if host else dev_view.stream
        if layout is not None:
            if layout not in _native.LAYOUTS:
                raise ValueError("layout must be 'RCHW' ([R, C, PH, PW]) or 'RHWC' ([R, PH, PW, C])")
            values['__top_layout__'] = _native.LAYOUTS[layout]
        if self._conv:
            if mode is not None or accumulate is not None or algo is not None or aligned or layout is not None:
                raise TypeError('%s takes no mode / accumulate / algo / aligned / layout' % self.name)
            _launch_conv(self.name, values, tensors, -1 if host else dev_id, stream, host)
        else:
            self._launcher(values, tensors, -1 if host else dev_id, stream, host, mode, accumulate,
                           algo or config.ROIALIGN_ALGO, aligned)
        for t in tensors.values():
            if t.writeback:
                t.finish()
        return None

    def build(self, ctx=None, template_types=None):
        """The reference pre-compiles a template instance here (func.py:351-405); the
        library is prebuilt, so this only checks that it loads."""
        _native.load()


def _rows(values, per_roi_name='nthreads'):
    per_roi = values['channels'] * values['pooled_height'] * values['pooled_width']
    return values[per_roi_name] // per_roi if per_roi > 0 else 0


def _vp(p):
    return p if isinstance(p, ctypes.c_void_p) else ctypes.c_void_p(p)


def _launch_forward(v, t, dev_id, stream, host, mode, accumulate, algo, aligned=False):
    if mode is not None:
        raise TypeError('roi_align_forward has no `mode`')
    lib = _native.load()
    data, rois, out = t['bottom_data'], t['bottom_rois'], t['top_data']
    num_rois = _rows(v)
    shape = data.view.shape
    num_images = shape[0] if len(shape) == 4 else -1
    flags = _native.make_flags(accumulate=bool(accumulate), host=host, algo=algo, aligned=aligned)
    _native.check(lib.mobula_roialign_forward_ex(
        dev_id, _vp(stream), _vp(data.resolve(False)), _vp(rois.resolve(False)),
        _vp(out.resolve(True)), v.get('__top_dtype__', 0), v.get('__top_layout__', 0), num_rois, num_images,
        v['channels'], v['height'], v['width'], v['pooled_height'], v['pooled_width'], v['spatial_scale'],
        v['sampling_ratio'], flags))


def _launch_backward(v, t, dev_id, stream, host, mode, accumulate, algo, aligned=False):
    lib = _native.load()
    dy, rois, dx = t['top_diff'], t['bottom_rois'], t['bottom_diff']
    if mode is None:
        mode = 'deterministic' if config.ROIALIGN_DETERMINISTIC else 'atomic'
    if mode not in ('atomic', 'deterministic'):
        raise ValueError("mode must be 'atomic' or 'deterministic'")
    if accumulate is None:
        accumulate = True       # reference kernel semantics (ROIAlign.cpp:148-157)
    num_rois = _rows(v)
    shape = dx.view.shape
    num_images = shape[0] if len(shape) == 4 else -1
    flags = _native.make_flags(accumulate=bool(accumulate), host=host, algo=algo, aligned=aligned)
    _native.check(lib.mobula_roialign_backward_ex(
        dev_id, _vp(stream), _vp(dy.resolve(False)), v.get('__top_dtype__', 0), v.get('__top_layout__', 0),
        _vp(rois.resolve(False)), _vp(dx.resolve(True)), num_rois, num_images, v['channels'], v['height'], v['width'],
        v['pooled_height'], v['pooled_width'], v['spatial_scale'], v['sampling_ratio'],
        _native.BWD_DETERMINISTIC if mode == 'deterministic' else _native.BWD_ATOMIC, flags))


# ---- the common case without adapter objects: three contiguous torch tensors of the expected
# types on one device.  Anything else (other frameworks, copies to make, errors to report) returns
# False and takes the general path above, which owns the error messages.
_INT_POSITIONS = (0, 3, 4, 5, 6, 7, 8)      # nthreads, channels .. sampling_ratio in both signatures
_TORCH_TOP = None          # torch dtype -> DTYPE_* of the pooled tensor, filled on first use


def _torch_fast_setup(torch):
    global _TORCH_TOP, _raw_stream
    _TORCH_TOP = {torch.float32: _native.DTYPE_F32, torch.float16: _native.DTYPE_F16,
                  torch.bfloat16: _native.DTYPE_BF16}
    raw = getattr(torch._C, '_cuda_getCurrentRawStream', None)
    if raw is None:
        raw = lambda index: torch.cuda.current_stream(index).cuda_stream
    _raw_stream = raw


def _fast_torch_common(args, ia, ib, ic, top, layout):
    """(lib, dev_id, stream, host, top_dtype, top_layout, num_rois, num_images) or None"""
    torch = sys.modules.get('torch')
    if torch is None:
        return None
    a, b, c = args[ia], args[ib], args[ic]
    T = torch.Tensor
    if type(a) is not T or type(b) is not T or type(c) is not T:
        return None
    if _TORCH_TOP is None:
        _torch_fast_setup(torch)
    f32 = torch.float32
    t = args[top]
    top_dtype = _TORCH_TOP.get(t.dtype)
    if top_dtype is None:
        return None
    for x in (a, b, c):
        if x is not t and x.dtype is not f32:
            return None
    dev = a.device
    if b.device != dev or c.device != dev:
        return None
    if not (a.is_contiguous() and b.is_contiguous() and c.is_contiguous()):
        return None
    scale = args[2]
    if type(scale) is not float and type(scale) is not int:
        return None
    for k in _INT_POSITIONS:
        if type(args[k]) is not int:
            return None
    if layout is None:
        top_layout = 0
    else:
        top_layout = _native.LAYOUTS.get(layout)
        if top_layout is None:
            return None
    if dev.type == 'cuda':
        dev_id = dev.index
        if dev_id is None:
            dev_id = torch.cuda.current_device()
        stream = _raw_stream(dev_id)
        host = False
    elif dev.type == 'cpu':
        dev_id, stream, host = -1, 0, True
    else:
        return None
    per_roi = args[3] * args[6] * args[7]
    num_rois = args[0] // per_roi if per_roi > 0 else 0
    return _native.load(), dev_id, stream, host, top_dtype, top_layout, num_rois


def _fast_torch_forward(args, mode, accumulate, algo, aligned, layout):
    if mode is not None:
        return False
    r = _fast_torch_common(args, 1, 9, 10, 10, layout)
    if r is None:
        return False
    lib, dev_id, stream, host, top_dtype, top_layout, num_rois = r
    data = args[1]
    flags = _native.make_flags(bool(accumulate), host, algo or config.ROIALIGN_ALGO, False, aligned)
    _native.check(lib.mobula_roialign_forward_ex(
        dev_id, stream, data.data_ptr(), args[9].data_ptr(), args[10].data_ptr(), top_dtype, top_layout, num_rois,
        data.shape[0] if data.dim() == 4 else -1, args[3], args[4], args[5], args[6], args[7], float(args[2]),
        args[8], flags))
    return True


def _fast_torch_backward(args, mode, accumulate, algo, aligned, layout):
    if mode is None:
        mode = 'deterministic' if config.ROIALIGN_DETERMINISTIC else 'atomic'
    elif mode != 'atomic' and mode != 'deterministic':
        return False
    r = _fast_torch_common(args, 1, 9, 10, 1, layout)
    if r is None:
        return False
    lib, dev_id, stream, host, top_dtype, top_layout, num_rois = r
    dx = args[9]
    flags = _native.make_flags(True if accumulate is None else bool(accumulate), host, algo or config.ROIALIGN_ALGO,
                               False, aligned)
    _native.check(lib.mobula_roialign_backward_ex(
        dev_id, stream, args[1].data_ptr(), top_dtype, top_layout, args[10].data_ptr(), dx.data_ptr(), num_rois,
        dx.shape[0] if dx.dim() == 4 else -1, args[3], args[4], args[5], args[6], args[7], float(args[2]), args[8],
        _native.BWD_DETERMINISTIC if mode == 'deterministic' else _native.BWD_ATOMIC, flags))
    return True


def _launch_conv(name, v, t, dev_id, stream, host):
    """im2col / col2im: `n` counts the elements of the output, the channel count is implied
    (Convolution.cpp:97-104, 117-124)."""
    lib = _native.load()
    to_col = name == 'im2col'
    src, dst = (t['data_im'], t['data_col']) if to_col else (t['data_col'], t['data_im'])
    per_channel = (v['kernel_h'] * v['kernel_w'] * v['height_col'] * v['width_col'] if to_col
                   else v['height'] * v['width'])
    if per_channel <= 0 or v['n'] % per_channel:
        raise ValueError('%s: n = %d is not a whole number of channels' % (name, v['n']))
    entry = lib.mobula_im2col_f32 if to_col else lib.mobula_col2im_f32
    _native.check(entry(
        dev_id, _vp(stream), _vp(src.resolve(False)), v['n'] // per_channel, v['height'], v['width'],
        v['kernel_h'], v['kernel_w'], v['pad_h'], v['pad_w'], v['stride_h'], v['stride_w'],
        v['dilation_h'], v['dilation_w'], v['height_col'], v['width_col'], _vp(dst.resolve(True)),
        _native.make_flags(host=host)))


_bound = {}


def bind(names=None):
    """Publishes kernels as module attributes, like the reference's `bind`
    (func.py:411-425) does after parsing an operator's .cpp file."""
    for name in (names or _SIGNATURES):
        if name not in _SIGNATURES:
            raise KeyError('no prebuilt kernel named %s' % name)
        if name not in _bound:
            _bound[name] = MobulaFunc(name, _SIGNATURES[name])
        globals()[name] = _bound[name]
    return _bound


def __getattr__(name):
    # kernels become visible after mobula.op.load('ROIAlign'), as in the reference
    raise AttributeError("module 'mobula.func' has no attribute %r (did you call "
                         "mobula.op.load('ROIAlign')?)" % name)
