"""Maps tensor types to adapter modules, importing each adapter on first use
(reference: /root/reference/mobula/glue/backend.py:60-139).

Differences from the reference, all defect fixes:
* an unknown type from a known package (e.g. numpy.int64 scalars) resolves to None
  instead of raising KeyError (reference backend.py:96-101);
* CuPy is registered under its current class path `cupy.ndarray`;
* MXNet is not supported.
"""
import importlib
from itertools import chain

from .common import MobulaTensor

_TYPE_TO_GLUE = {}
_NAME_TO_GLUE = {}
# top-level package -> (adapter name, dotted class paths)
_KNOWN = {
    'numpy': ('numpy', ['numpy.ndarray']),
    'torch': ('torch', ['torch.Tensor']),
    'cupy': ('cupy', ['cupy.ndarray']),
}


def register_glue(glue_name, type_names):
    _NOT_TENSORS.clear()
    assert type_names, ValueError('type_names should be not empty')
    pkgs = set(name.split('.')[0] for name in type_names)
    assert len(pkgs) == 1, TypeError('types of one adapter must share a package: %s' % type_names)
    _KNOWN[pkgs.pop()] = (glue_name, list(type_names))


def _resolve(dotted):
    parts = dotted.split('.')
    obj = importlib.import_module(parts[0])
    for part in parts[1:]:
        obj = getattr(obj, part)
    return obj


def _activate(glue_name, type_names):
    if glue_name in _NAME_TO_GLUE:
        return _NAME_TO_GLUE[glue_name]
    try:
        mod = importlib.import_module('.%s_glue' % glue_name, __package__)
    except ImportError:
        return None
    assert isinstance(mod.Tensor, type) and issubclass(mod.Tensor, MobulaTensor)
    mod.gen_cache = {}
    for dotted in type_names:
        try:
            _TYPE_TO_GLUE[_resolve(dotted)] = mod
        except (ImportError, AttributeError):
            pass
    _NAME_TO_GLUE[glue_name] = mod
    return mod


_NOT_TENSORS = set()       # types already found to belong to no adapter (int, float, tuple, str, ...)


def get_var_type_glue(vtype):
    mod = _TYPE_TO_GLUE.get(vtype)
    if mod is not None:
        return mod
    if vtype in _NOT_TENSORS:
        return None
    pkg = getattr(vtype, '__module__', '').split('.')[0]
    if pkg not in _KNOWN:
        try:
            _NOT_TENSORS.add(vtype)
        except TypeError:      # unhashable type object: look it up again next time
            pass
        return None
    _activate(*_KNOWN[pkg])
    for known, mod in _TYPE_TO_GLUE.items():
        if vtype is known:
            return mod
    # subclasses of a registered tensor type (torch.nn.Parameter, numpy.matrix, ...)
    for known, mod in list(_TYPE_TO_GLUE.items()):
        if isinstance(vtype, type) and issubclass(vtype, known):
            _TYPE_TO_GLUE[vtype] = mod
            return mod
    return None


def get_var_glue(var):
    return get_var_type_glue(type(var))


def get_args_glue(*args, **kwargs):
    mods = [m for m in map(get_var_glue, chain(args, kwargs.values())) if m is not None]
    if not mods:
        return None
    if any(m is not mods[0] for m in mods):
        raise TypeError('Support only 1 backend in a call, now: %s' % mods)
    return mods[0]


def op_gen(glue_mod, op, name):
    if name not in glue_mod.gen_cache:
        glue_mod.gen_cache[name] = glue_mod.OpGen(op=op, name=name)
    return glue_mod.gen_cache[name]


def get_glue_modules():
    return list(_NAME_TO_GLUE.values())
