"""`mobula.op.load(name)` for prebuilt operators.

The reference's loader parses `<Op>.cpp` for MOBULA_KERNEL signatures, binds them as
`mobula.func.*` and later JIT-compiles them per argument signature
(/root/reference/mobula/op/loader.py:623-659, 385-535).  Here the kernels are already
inside libmobula_roialign_sm100.so, so loading an operator means: check the library
loads (no fallback exists), publish its kernels on `mobula.func`, and execute the
operator's Python class file so that `@register` publishes `mobula.op.<Name>`.
"""
import importlib.util
import os
import sys

from .. import _native
from .. import func

OPZOO = os.path.normpath(os.path.join(os.path.dirname(__file__), '..', 'opzoo'))

# operator -> kernels of the native library it launches
PREBUILT = {'ROIAlign': ('roi_align_forward', 'roi_align_backward'),
            'Convolution': ('im2col', 'col2im')}

_loaded = {}


def load(module_name, path=''):
    """Load an operator module by name (searched in the package's opzoo/ first)."""
    op_name = os.path.basename(module_name.rstrip('/'))
    if op_name in _loaded:
        return _loaded[op_name]
    base = path
    if not base and os.path.isdir(os.path.join(OPZOO, op_name)):
        base = OPZOO
    op_dir = os.path.join(base, module_name)
    py_file = os.path.join(op_dir, op_name + '.py')
    if not os.path.exists(py_file):
        py_file = os.path.join(op_dir, '__init__.py')
    if not os.path.exists(py_file):
        raise IOError('%s.py or __init__.py not found in the path %s' % (op_name, op_dir))
    if op_name in PREBUILT:
        _native.load()                      # raises when the sm_100a library is missing
        func.bind(PREBUILT[op_name])
    elif os.path.exists(os.path.join(op_dir, op_name + '.cpp')):
        raise NotImplementedError(
            'operator %s ships C++ kernels; this build has no JIT and only carries the '
            'prebuilt kernels of %s' % (op_name, sorted(PREBUILT)))
    spec = importlib.util.spec_from_file_location('mobula_opzoo_' + op_name, py_file)
    module = importlib.util.module_from_spec(spec)
    sys.modules[spec.name] = module
    spec.loader.exec_module(module)
    _loaded[op_name] = module
    return module
