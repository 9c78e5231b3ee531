"""mobulaop_b200 -- the ROIAlign hot path of MobulaOP on one prebuilt sm_100a library.

Importable as `mobulaop_b200` or, through the alias package `mobula/` at the repository
root, as `mobula` -- the name users of the reference import
(/root/reference/mobula/__init__.py:1-8):

    import mobula
    mobula.op.load('ROIAlign')
    y = mobula.op.ROIAlign(data, rois, pooled_size=(7, 7), spatial_scale=0.25, sampling_ratio=2)
"""
from .version import __version__
from . import const
from .config import config
from . import glue
from . import func
from . import op
from . import testing

function = func
operator = op

__all__ = ['__version__', 'config', 'const', 'func', 'function', 'glue', 'op', 'operator',
           'testing']
