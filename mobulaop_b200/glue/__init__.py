"""Framework adapters (PyTorch, NumPy, CuPy) for operators registered with
`mobula.op.register` -- the counterpart of /root/reference/mobula/glue/."""
from . import common
from . import backend
from .common import register, CUSTOM_OP_LIST, MobulaOperator, MobulaTensor

common.backend = backend
