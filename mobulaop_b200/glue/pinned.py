"""Page-locked host arrays for the results of the NumPy adapter.

The host-buffer launch path (MOBULA_ROIALIGN_HOST_BUFFERS) copies the result device -> host;
into pageable memory that copy runs at a fraction of the PCIe rate.  Results of 1 MiB and more
are therefore NumPy arrays over buffers from mobula_host_alloc (cudaMallocHost).  Page-locking
is slow, so released buffers are kept for the next array of the same size -- in an LRU pool
bounded by TOTAL bytes (MOBULA_PINNED_POOL_MB; default a quarter of the physical memory,
within 1..16 GiB): a detection workload changes R every step, so a per-size cache without a
global cap would pin host memory without limit.
The pool is emptied by `drain()` and at interpreter exit.
"""
import atexit
import collections
import ctypes
import os
import threading
import weakref

import numpy as np

_MIN_BYTES = 1 << 20


def _default_cap():
    env = os.environ.get('MOBULA_PINNED_POOL_MB')
    if env:
        return int(env) << 20
    try:
        phys = os.sysconf('SC_PHYS_PAGES') * os.sysconf('SC_PAGE_SIZE')
    except (ValueError, OSError, AttributeError):
        phys = 4 << 30
    return int(min(max(phys // 4, 1 << 30), 16 << 30))


_MAX_POOL_BYTES = _default_cap()
_pool = collections.OrderedDict()      # pointer -> nbytes, least recently released first
_pool_bytes = 0
_lock = threading.Lock()


def pooled_bytes():
    return _pool_bytes


def _pool_put(nbytes, ptr, free, cap=None):
    """Keeps a released buffer, evicting the least recently released ones beyond `cap`."""
    global _pool_bytes
    cap = _MAX_POOL_BYTES if cap is None else cap
    evict = []
    with _lock:
        if nbytes > cap:
            evict.append(ptr)
        else:
            _pool[ptr] = nbytes
            _pool_bytes += nbytes
            while _pool_bytes > cap:
                old, size = _pool.popitem(last=False)
                _pool_bytes -= size
                evict.append(old)
    for p in evict:
        free(p)


def _pool_take(nbytes):
    """Most recently released buffer of exactly `nbytes`, or None."""
    global _pool_bytes
    with _lock:
        for ptr in reversed(_pool):
            if _pool[ptr] == nbytes:
                del _pool[ptr]
                _pool_bytes -= nbytes
                return ptr
    return None


def _native_free(ptr):
    from .. import _native
    _native.load().mobula_host_free(ctypes.c_void_p(ptr))


def _release(nbytes, ptr):
    try:
        _pool_put(nbytes, ptr, _native_free)
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
    ptr = _pool_take(nbytes) or lib.mobula_host_alloc(nbytes)
    if not ptr:
        return np.empty(shape, dtype)
    buf = (ctypes.c_char * nbytes).from_address(ptr)
    weakref.finalize(buf, _release, nbytes, ptr)     # buf lives as long as any view of the array
    return np.frombuffer(buf, dtype=dtype).reshape(shape)


def drain():
    """Frees the cached buffers."""
    global _pool_bytes
    with _lock:
        ptrs = list(_pool)
        _pool.clear()
        _pool_bytes = 0
    if not ptrs:
        return
    try:
        for p in ptrs:
            _native_free(p)
    except Exception:       # library gone at shutdown
        pass


atexit.register(drain)
