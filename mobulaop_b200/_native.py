"""ctypes binding of libmobula_roialign_sm100.so (ABI: include/mobula_roialign.h).

There is no CPU, HIP or TVM fallback: if the library is missing or cannot be loaded the
import of the launch path fails loudly.  (The reference picks a context per call and
JIT-builds it, /root/reference/mobula/func.py:138-164.)
"""
import ctypes
import os

from . import build as _build

OK = 0
ERR_INVALID_ARGUMENT = 1
ERR_CUDA = 2
ERR_UNSUPPORTED = 3

BWD_ATOMIC = 0
BWD_DETERMINISTIC = 1

# element type / layout of the pooled tensor (mobula_roialign_{forward,backward}_ex)
DTYPE_F32, DTYPE_F16, DTYPE_BF16 = 0, 1, 2
LAYOUT_RCHW, LAYOUT_RHWC = 0, 1
LAYOUTS = {'RCHW': LAYOUT_RCHW, 'RHWC': LAYOUT_RHWC}

FLAG_ACCUMULATE = 1
FLAG_HOST_BUFFERS = 2
FLAG_ALIGNED = 4
ALGO_SHIFT = 8
ALGO_AUTO, ALGO_GATHER, ALGO_PLANE, ALGO_RESIDENT, ALGO_SWEEP, ALGO_PULL = 0, 1, 2, 3, 4, 5
ALGOS = {'auto': ALGO_AUTO, 'gather': ALGO_GATHER, 'plane': ALGO_PLANE, 'resident': ALGO_RESIDENT,
         'sweep': ALGO_SWEEP, 'pull': ALGO_PULL}

# every symbol include/mobula_roialign.h declares
EXPORTED_SYMBOLS = (
    'roi_align_forward_ab78de3c', 'roi_align_backward_f318a24e', 'set_device',
    'mobula_roialign_abi_version', 'mobula_last_error', 'mobula_roialign_forward_f32',
    'mobula_roialign_backward_f32', 'mobula_roialign_bookkeeping_f32', 'mobula_host_alloc',
    'mobula_host_free', 'mobula_roialign_launch_count', 'mobula_roialign_last_algo',
    'mobula_roialign_release',
    'im2col_341af5d3', 'col2im_341af5d3', 'mobula_im2col_f32', 'mobula_col2im_f32',
    'mobula_roialign_forward_ex', 'mobula_roialign_backward_ex',
    'mobula_roialign_prepare_f32', 'mobula_roialign_forward_prepared_f32',
    'mobula_roialign_backward_prepared_f32', 'mobula_roialign_plan_destroy',
)


class NativeLibraryError(RuntimeError):
    pass


_lib = None


def _declare(lib):
    i32, f32, vp, u32 = ctypes.c_int, ctypes.c_float, ctypes.c_void_p, ctypes.c_uint
    lib.mobula_roialign_abi_version.restype = i32
    lib.mobula_last_error.restype = ctypes.c_char_p
    lib.mobula_roialign_forward_f32.restype = i32
    lib.mobula_roialign_forward_f32.argtypes = [i32, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32,
                                                i32, f32, i32, u32]
    lib.mobula_roialign_backward_f32.restype = i32
    lib.mobula_roialign_backward_f32.argtypes = [i32, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32,
                                                 i32, f32, i32, i32, u32]
    lib.mobula_roialign_forward_ex.restype = i32
    lib.mobula_roialign_forward_ex.argtypes = [i32, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, i32,
                                               i32, f32, i32, u32]
    lib.mobula_roialign_backward_ex.restype = i32
    lib.mobula_roialign_backward_ex.argtypes = [i32, vp, vp, i32, i32, vp, vp, i32, i32, i32, i32, i32, i32,
                                                i32, f32, i32, i32, u32]
    lib.mobula_roialign_prepare_f32.restype = i32
    lib.mobula_roialign_prepare_f32.argtypes = [i32, vp, vp, i32, i32, i32, i32, i32, i32, f32, i32, u32,
                                                ctypes.POINTER(vp)]
    lib.mobula_roialign_forward_prepared_f32.restype = i32
    lib.mobula_roialign_forward_prepared_f32.argtypes = [vp, vp, vp, vp, i32, u32]
    lib.mobula_roialign_backward_prepared_f32.restype = i32
    lib.mobula_roialign_backward_prepared_f32.argtypes = [vp, vp, vp, vp, i32, i32, u32]
    lib.mobula_roialign_plan_destroy.restype = None
    lib.mobula_roialign_plan_destroy.argtypes = [vp]
    lib.mobula_roialign_bookkeeping_f32.restype = i32
    lib.mobula_roialign_bookkeeping_f32.argtypes = [i32, vp, vp, i32, i32, i32, i32, i32, f32, i32,
                                                    vp, vp, i32, vp, vp]
    lib.mobula_host_alloc.restype = vp
    lib.mobula_host_alloc.argtypes = [ctypes.c_size_t]
    lib.mobula_host_free.restype = None
    lib.mobula_host_free.argtypes = [vp]
    lib.mobula_roialign_launch_count.restype = ctypes.c_uint64
    lib.mobula_roialign_last_algo.restype = ctypes.c_char_p
    lib.mobula_roialign_last_algo.argtypes = [i32]
    lib.mobula_roialign_release.restype = None
    lib.set_device.restype = None
    lib.set_device.argtypes = [i32]
    # the two drop-in symbols keep the reference convention: no argtypes, callers pass
    # explicit ctypes instances (mobula/func.py:248-255, 335-337)
    lib.roi_align_forward_ab78de3c.restype = None
    lib.roi_align_backward_f318a24e.restype = None
    lib.im2col_341af5d3.restype = None
    lib.col2im_341af5d3.restype = None
    conv = [i32, vp, vp] + [i32] * 13 + [vp, u32]
    lib.mobula_im2col_f32.restype = i32
    lib.mobula_im2col_f32.argtypes = conv
    lib.mobula_col2im_f32.restype = i32
    lib.mobula_col2im_f32.argtypes = conv


def library_path():
    return _build.LIB_PATH


def load():
    """Returns the loaded library; raises NativeLibraryError when it is not usable."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise NativeLibraryError(
            '%s is not built: run `python -m mobulaop_b200.build` (needs nvcc). '
            'There is no CPU fallback for the ROIAlign path.' % path)
    try:
        lib = ctypes.CDLL(path)
    except OSError as err:
        raise NativeLibraryError('cannot load %s: %s' % (path, err))
    missing = [s for s in EXPORTED_SYMBOLS if not hasattr(lib, s)]
    if missing:
        raise NativeLibraryError('%s lacks symbols %s' % (path, missing))
    _declare(lib)
    _lib = lib
    return lib


def check(status):
    if status != OK:
        msg = load().mobula_last_error().decode('utf-8', 'replace')
        if status == ERR_INVALID_ARGUMENT:
            raise ValueError(msg)
        if status == ERR_UNSUPPORTED:
            raise NotImplementedError(msg)
        raise RuntimeError('CUDA failure in libmobula_roialign_sm100: ' + msg)


FLAG_DEBUG_NO_TABLES = 0x10000


def make_flags(accumulate=False, host=False, algo='auto', no_tables=False, aligned=False):
    flags = (FLAG_ACCUMULATE if accumulate else 0) | (FLAG_HOST_BUFFERS if host else 0)
    if aligned:
        flags |= FLAG_ALIGNED
    if no_tables:
        flags |= FLAG_DEBUG_NO_TABLES
    return flags | (ALGOS[algo] << ALGO_SHIFT)


def launch_count():
    return int(load().mobula_roialign_launch_count())


def last_algo(backward=False):
    return load().mobula_roialign_last_algo(1 if backward else 0).decode()


class RoiPlan(object):
    """Prepared ROI table (mobula_roialign_prepare_f32): validated, copied and grouped by image
    once, then shared by forward and backward calls.  `rois`: device pointer of the [R, 5] table."""

    def __init__(self, device_id, stream, rois_ptr, num_rois, num_images, height, width, pooled, scale,
                 sampling_ratio, aligned=False):
        lib = load()
        self._lib = lib
        self.handle = ctypes.c_void_p()
        check(lib.mobula_roialign_prepare_f32(device_id, stream, rois_ptr, num_rois, num_images, height, width,
                                              pooled[0], pooled[1], scale, sampling_ratio,
                                              FLAG_ALIGNED if aligned else 0, ctypes.byref(self.handle)))

    def forward(self, stream, data_ptr, out_ptr, channels, flags=0):
        check(self._lib.mobula_roialign_forward_prepared_f32(self.handle, stream, data_ptr, out_ptr, channels, flags))

    def backward(self, stream, dy_ptr, dx_ptr, channels, mode=BWD_ATOMIC, flags=0):
        check(self._lib.mobula_roialign_backward_prepared_f32(self.handle, stream, dy_ptr, dx_ptr, channels, mode, flags))

    def close(self):
        if self.handle:
            self._lib.mobula_roialign_plan_destroy(self.handle)
            self.handle = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
