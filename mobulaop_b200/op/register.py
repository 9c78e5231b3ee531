"""`@mobula.op.register` (reference: /root/reference/mobula/op/register.py:5-13)."""
from .. import glue


def register(*args, **kwargs):
    return glue.register(*args, **kwargs)
