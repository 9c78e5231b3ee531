"""`import mobula` -> the B200 ROIAlign build (mobulaop_b200).

MobulaOP users import the package as `mobula`
(/root/reference/mobula/__init__.py); this alias makes `mobula`, `mobula.op`,
`mobula.func`, `mobula.glue`, `mobula.config` ... the very same module objects as
their mobulaop_b200 counterparts, so state (registered operators, bound kernels) is
shared whichever name is used.
"""
import importlib
import importlib.abc
import importlib.util
import sys

import mobulaop_b200 as _impl

_SRC, _DST = 'mobulaop_b200', __name__


class _AliasLoader(importlib.abc.Loader):
    def __init__(self, target):
        self.target = target

    def create_module(self, spec):
        return importlib.import_module(self.target)

    def exec_module(self, module):
        pass


class _AliasFinder(importlib.abc.MetaPathFinder):
    def find_spec(self, fullname, path=None, target=None):
        if not fullname.startswith(_DST + '.'):
            return None
        real = _SRC + fullname[len(_DST):]
        if importlib.util.find_spec(real) is None:
            return None
        return importlib.util.spec_from_loader(fullname, _AliasLoader(real))


sys.meta_path.insert(0, _AliasFinder())
for _name, _mod in list(sys.modules.items()):
    if _name == _SRC or _name.startswith(_SRC + '.'):
        sys.modules[_DST + _name[len(_SRC):]] = _mod
sys.modules[_DST] = _impl
