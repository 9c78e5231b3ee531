"""Page-locked host arrays for the results of the NumPy adapter.

The host-buffer launch path (MOBULA_ROIALIGN_HOST_BUFFERS) copies the result device -> host;
into pageable memory that copy runs at a fraction of the PCIe rate.  Results of 1 MiB and more
are therefore NumPy arrays over buffers from mobula_host_alloc (cudaMallocHost).  Page-locking
is slow, so a released buffer is kept for the next array of the same size (a few per size).
"""
import ctypes
import weakref

import numpy as np

_MIN_BYTES = 1 << 20
_KEEP_PER_SIZE = 4
_free = {}          # nbytes -> [pointer]


def _release(nbytes, ptr):
    try:
        from .. import _native
        pool = _free.setdefault(nbytes, [])
        if len(pool) < _KEEP_PER_SIZE:
            pool.append(ptr)
        else:
            _native.load().mobula_host_free(ctypes.c_void_p(ptr))
    except Exception:       # interpreter shutdown
        pass


def empty(shape, dtype):
    """np.empty(shape, dtype), page-locked when it is large and the library is available."""
    dtype = np.dtype(dtype)
    shape = tuple(int(s) for s in shape)
    nbytes = int(np.prod(shape, dtype=np.int64)) * dtype.itemsize
    if nbytes < _MIN_BYTES:
        return np.empty(shape, dtype)
    try:
        from .. import _native
        lib = _native.load()
    except Exception:
        return np.empty(shape, dtype)
    pool = _free.get(nbytes)
    ptr = pool.pop() if pool else lib.mobula_host_alloc(nbytes)
    if not ptr:
        return np.empty(shape, dtype)
    buf = (ctypes.c_char * nbytes).from_address(ptr)
    weakref.finalize(buf, _release, nbytes, ptr)     # buf lives as long as any view of the array
    return np.frombuffer(buf, dtype=dtype).reshape(shape)


def drain():
    """Frees the cached buffers."""
    from .. import _native
    lib = _native.load()
    for pool in _free.values():
        while pool:
            lib.mobula_host_free(ctypes.c_void_p(pool.pop()))
