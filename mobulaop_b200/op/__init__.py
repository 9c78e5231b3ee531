"""`mobula.op`: operator registry and loader (reference: /root/reference/mobula/op/)."""
from .. import glue

glue.common.OP_MODULE_GLOBALS = globals()

from .register import register    # noqa: E402
from .loader import load          # noqa: E402

__all__ = ['register', 'load']
