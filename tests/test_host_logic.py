"""Host-side mirror of the reference interface, exercised without a GPU: registry,
adapters' argument handling, the launch path's type checks, config, sharding."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT


@pytest.fixture(scope='module')
def mobula(native):
    import mobula
    mobula.op.load('ROIAlign')
    return mobula


def test_alias_package_is_the_implementation(mobula):
    import mobulaop_b200
    import mobula.func as f1
    import mobula.glue.backend as b1
    assert mobula is mobulaop_b200
    assert f1 is mobulaop_b200.func and b1 is mobulaop_b200.glue.backend
    assert mobula.operator is mobula.op and mobula.function is mobula.func


def test_load_publishes_op_and_kernels(mobula):
    assert isinstance(mobula.op.ROIAlign, mobula.glue.MobulaOperator)
    assert mobula.func.roi_align_forward.arg_names[:2] == ['nthreads', 'bottom_data']
    # argument order of backward differs from forward (ROIAlign.cpp:82-83 vs :15-16)
    assert mobula.func.roi_align_backward.arg_names[-2:] == ['bottom_diff', 'bottom_rois']
    assert mobula.func.roi_align_forward.arg_names[-2:] == ['bottom_rois', 'top_data']
    with pytest.raises(IOError):
        mobula.op.load('NoSuchOperator')
    with pytest.raises(AttributeError):
        mobula.func.softmax_forward


def test_infer_shape_and_ctor(mobula):
    op = mobula.op.ROIAlign[np.ndarray](pooled_size=(3, 4), spatial_scale=0.5, sampling_ratio=0)
    assert op.infer_shape([(5, 3, 16, 16), (7, 5)])[1] == [[7, 3, 3, 4]]
    with pytest.raises(AssertionError):
        op.infer_shape([(5, 3, 16, 16), (7, 4)])
    with pytest.raises(AssertionError):
        op.infer_shape([(3, 16, 16), (7, 5)])
    assert mobula.op.ROIAlign[np.ndarray](pooled_size=7, spatial_scale=1., sampling_ratio=2).pooled_size == (7, 7)


def test_split_inputs():
    from mobulaop_b200.glue.common import split_inputs

    class Op:
        def forward(self, data, rois):
            pass
    a, b = object(), object()
    assert split_inputs(Op, (a, b, 7), dict(k=1)) == ([a, b], [7], dict(k=1))
    assert split_inputs(Op, (a,), dict(rois=b, k=1)) == ([a, b], [], dict(k=1))
    assert split_inputs(Op, (), dict(data=a, rois=b)) == ([a, b], [], {})
    with pytest.raises(AssertionError):
        split_inputs(Op, (a,), dict(k=1))


def test_backend_resolution(mobula):
    import torch
    be = mobula.glue.backend
    assert be.get_var_glue(np.zeros(1)).__name__.endswith('numpy_glue')
    assert be.get_var_glue(torch.zeros(1)).__name__.endswith('torch_glue')
    assert be.get_var_glue(torch.nn.Parameter(torch.zeros(1))).__name__.endswith('torch_glue')
    # the reference raises KeyError for numpy scalars; here: not a tensor
    assert be.get_var_glue(np.int64(3)) is None
    assert be.get_var_glue(3.5) is None
    with pytest.raises(TypeError):
        be.get_args_glue(np.zeros(1), torch.zeros(1))
    with pytest.raises(ValueError):
        mobula.op.ROIAlign(pooled_size=(7, 7), spatial_scale=1.0, sampling_ratio=2)


def test_func_type_checks_happen_before_launch(mobula):
    f = mobula.func.roi_align_forward
    data = np.zeros((1, 2, 4, 4), np.float32)
    rois = np.zeros((1, 5), np.float32)
    out = np.zeros((1, 2, 2, 2), np.float32)
    with pytest.raises(TypeError):          # template T bound to float, rois is double
        f(out.size, data, 1.0, 2, 4, 4, 2, 2, 1, rois.astype(np.float64), out)
    with pytest.raises(TypeError):          # double is not instantiated
        f(out.size, data.astype(np.float64), 1.0, 2, 4, 4, 2, 2, 1, rois.astype(np.float64),
          out.astype(np.float64))
    with pytest.raises(TypeError):          # a scalar where a tensor is expected
        f(out.size, 3, 1.0, 2, 4, 4, 2, 2, 1, rois, out)
    with pytest.raises(TypeError):
        f(out.size, data, 1.0, 2, 4, 4, 2, 2, 1, rois)
    with pytest.raises(TypeError):
        f(out.size, data, 1.0, 2, 4, 4, 2, 2, 1, rois, out, mode='atomic')
    with pytest.raises(ValueError):
        mobula.func.roi_align_backward(out.size, out, 1.0, 2, 4, 4, 2, 2, 1, data, rois, mode='fast')


def test_config_contract(mobula):
    cfg = mobula.config
    assert cfg.GPU_BACKEND == 'cuda' and cfg.ROIALIGN_DETERMINISTIC is False
    with pytest.raises(AttributeError):
        cfg.NO_SUCH_FLAG = 1
    with pytest.raises(TypeError):
        cfg.ROIALIGN_DETERMINISTIC = 1
    with pytest.raises(ValueError):
        cfg.GPU_BACKEND = 'hip'
    cfg.HOST_NUM_THREADS = 1            # reference build flag: accepted, inert
    cfg.HOST_NUM_THREADS = 0
    with cfg.TempConfig(ROIALIGN_DETERMINISTIC=True):
        assert cfg.ROIALIGN_DETERMINISTIC is True
    assert cfg.ROIALIGN_DETERMINISTIC is False
    with pytest.raises(AttributeError):
        with cfg.TempConfig(NOPE=1):
            pass


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from mobulaop_b200 import _native, build
    monkeypatch.setattr(_native, '_lib', None)
    monkeypatch.setattr(build, 'LIB_PATH', str(tmp_path / 'absent.so'))
    with pytest.raises(_native.NativeLibraryError):
        _native.load()


def test_testing_helpers():
    from mobulaop_b200 import testing as T
    a = np.array([1.0, -2.0, 0.0], np.float32)
    assert T.ulp_distance(a, a).max() == 0
    assert T.ulp_distance(a, np.nextafter(a, np.float32(9))).tolist() == [1, 1, 1]
    assert T.ulp_distance(np.float32([0.0]), np.float32([-0.0])).tolist() == [0]
    T.assert_within_ulp_or_rel(a, np.nextafter(np.nextafter(a, np.float32(9)), np.float32(9)))
    with pytest.raises(AssertionError):
        T.assert_within_ulp_or_rel(np.float32([1e-3]), np.float32([1.1e-3]))
    with pytest.raises(AssertionError):
        T.assert_bit_exact(np.float32([0.0]), np.float32([-0.0]))
    T.assert_almost_equal(a, a, rtol=0, atol=0)
    T.assert_almost_equal(a[:2], a[:2] * (1 + 1e-7), atol=1e-6)
    with pytest.raises(AssertionError):
        T.assert_almost_equal(a, a + 1e-3)


# ---- sharding ----------------------------------------------------------------------
def _random_rois(rng, R, N):
    rois = rng.uniform(0, 100, (R, 5)).astype(np.float32)
    rois[:, 0] = rng.integers(-1, N + 1, R)          # some out of range on both sides
    return rois


@pytest.mark.parametrize('N,G', [(1, 1), (2, 2), (5, 2), (8, 4), (64, 8), (3, 8)])
def test_shards_partition_rois_exactly(N, G):
    from mobulaop_b200 import sharding as S
    rng = np.random.default_rng(N * 31 + G)
    rois = _random_rois(rng, 200, N)
    shards = S.make_shards(rois, N, G)
    assert [s.image_begin for s in shards] == [(g * N) // G for g in range(G)]
    assert shards[-1].image_end == N and shards[0].image_begin == 0
    seen = np.concatenate([s.roi_index for s in shards])
    valid = np.nonzero((rois[:, 0] >= 0) & (rois[:, 0] < N))[0]
    assert sorted(seen.tolist()) == valid.tolist()             # every valid ROI exactly once
    for s in shards:
        assert (np.diff(s.roi_index) > 0).all()                 # stable: original order kept
        np.testing.assert_array_equal(s.rois[:, 1:], rois[s.roi_index, 1:])
        np.testing.assert_array_equal(s.rois[:, 0] + s.image_begin, rois[s.roi_index, 0])
        assert ((s.rois[:, 0] >= 0) & (s.rois[:, 0] < max(s.num_images, 1))).all()
        own = S.owner_of_images(S.batch_indices(rois[s.roi_index]), G, N)
        assert (own == s.rank).all()
    rows = rng.standard_normal((200, 3)).astype(np.float32)
    back = S.assemble_rows(shards, [S.scatter_rows(s, rows) for s in shards], 200)
    np.testing.assert_array_equal(back[valid], rows[valid])
    assert not back[np.setdiff1d(np.arange(200), valid)].any()


def test_sharded_result_equals_unsharded_bitwise(oracle):
    """Per-image ROI ownership must not change a single bit of y or dX."""
    from mobulaop_b200 import sharding as S
    rng = np.random.default_rng(5)
    N, C, H, W, R = 5, 3, 12, 15, 64
    data = rng.standard_normal((N, C, H, W)).astype(np.float32)
    rois = _random_rois(rng, R, N)
    rois[:, 1:] *= 0.6
    dy = rng.standard_normal((R, C, 3, 3)).astype(np.float32)
    keep = (rois[:, 0] >= 0) & (rois[:, 0] < N)
    rois[~keep, 1:] = 1e6                   # make unowned ROIs contribute nothing anywhere
    full_rois = rois.copy()
    full_rois[~keep, 0] = 0
    y = oracle.forward(data, full_rois, (3, 3), 0.25, 2)
    dx = oracle.backward(dy, full_rois, data.shape, 0.25, 2)
    for G in (2, 3):
        shards = S.make_shards(rois, N, G)
        ys, dxs = [], []
        for s in shards:
            local = np.ascontiguousarray(data[s.image_begin:s.image_end])
            ys.append(oracle.forward(local, s.rois, (3, 3), 0.25, 2))
            dxs.append(oracle.backward(S.scatter_rows(s, dy), s.rois, local.shape, 0.25, 2))
        np.testing.assert_array_equal(S.assemble_rows(shards, ys, R), np.where(keep[:, None, None, None], y, 0))
        np.testing.assert_array_equal(S.assemble_images(shards, dxs), dx)


def test_two_rank_gloo_driver(tmp_path):
    """world_size-2 run of the sharded driver on CPU (gloo): each rank computes its shard
    with the oracle standing in for the kernels, results are gathered and compared."""
    script = os.path.join(ROOT, 'tests', 'gloo_shard_worker.py')
    env = dict(os.environ, PYTHONPATH=ROOT, MASTER_ADDR='127.0.0.1')
    out = subprocess.run(
        [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=2',
         '--master-addr', '127.0.0.1', '--master-port', '29611', script, str(tmp_path)],
        capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert 'SHARD-OK' in out.stdout


def test_result_arrays_fall_back_to_pageable_memory_without_a_gpu():
    """mobulaop_b200.glue.pinned: small results are plain NumPy arrays; large ones are
    page-locked when cudaMallocHost works and plain arrays otherwise (this container)."""
    from mobulaop_b200.glue import pinned
    small = pinned.empty((3, 5), np.float32)
    assert small.shape == (3, 5) and small.dtype == np.float32
    big = pinned.empty((1 << 19,), np.float32)           # 2 MiB
    assert big.shape == (1 << 19,) and big.dtype == np.float32
    big[:] = 1.0
    assert float(big.sum()) == float(1 << 19)
    view = big[::2]
    del big                                              # the buffer outlives its first owner
    assert view[0] == 1.0
    pinned.drain()


def test_cupy_adapter_attribute_mapping(mobula):
    """CuPy is absent from the image: the adapter is exercised against a stub
    module exposing .data.ptr / .device.id / .dtype / .flags (tests/cupy_stub.py).  No launch
    here -- the GPU leg is tests/test_parity_gpu.py::test_cupy_operator_through_stub."""
    import ctypes
    try:
        import cupy
    except ImportError:
        import cupy_stub
        cupy = cupy_stub.install()
    if not getattr(cupy, '__stub__', False):
        pytest.skip('a real CuPy is installed; the stub test does not apply')
    be = mobula.glue.backend
    a = cupy.zeros((2, 3, 4, 5), np.float32)
    glue_mod = be.get_var_glue(a)
    assert glue_mod is not None and glue_mod.__name__.endswith('cupy_glue')
    assert be.get_var_type_glue(cupy.ndarray) is glue_mod
    view = glue_mod.Tensor(a)
    assert view.ctype is ctypes.c_float
    assert view.dev_id == a.device.id
    assert view.shape == (2, 3, 4, 5)
    ptr = view.data_ptr
    assert isinstance(ptr, ctypes.c_void_p) and ptr.value == a.data.ptr
    strided = a[:, :, ::2]
    assert not strided.flags.c_contiguous
    ptr2, keep = glue_mod.Tensor(strided).data_ptr        # copy made, kept alive for the call
    assert keep.flags.c_contiguous and ptr2.value == keep.data.ptr
    assert isinstance(view.stream, int)
    # the operator class is generated for the CuPy array module
    op = mobula.op.ROIAlign[cupy.ndarray](pooled_size=(2, 2), spatial_scale=1.0, sampling_ratio=1)
    assert op.F is cupy
    assert op.infer_shape([(1, 3, 8, 8), (4, 5)])[1] == [[4, 3, 2, 2]]
    with pytest.raises(TypeError):
        be.get_args_glue(np.zeros(1), a)


def test_pinned_pool_is_bounded():
    """The result cache of the NumPy adapter keeps at most `_MAX_POOL_BYTES` of page-locked
    memory, whatever sizes come and go (detection workloads change R every step)."""
    from mobulaop_b200.glue import pinned
    pinned.drain()
    freed = []
    pinned._pool_put(100, 0xA0, free=freed.append, cap=250)
    pinned._pool_put(100, 0xB0, free=freed.append, cap=250)
    assert pinned.pooled_bytes() == 200 and not freed
    pinned._pool_put(100, 0xC0, free=freed.append, cap=250)     # evicts the oldest
    assert freed == [0xA0] and pinned.pooled_bytes() == 200
    pinned._pool_put(300, 0xD0, free=freed.append, cap=250)     # larger than the cap: not kept
    assert freed == [0xA0, 0xD0] and pinned.pooled_bytes() == 200
    assert pinned._pool_take(100) == 0xC0                        # most recently released first
    assert pinned._pool_take(100) == 0xB0
    assert pinned._pool_take(100) is None and pinned.pooled_bytes() == 0


def test_fpn_level_assignment_matches_the_paper_and_torchvision():
    """k = floor(k0 + log2(sqrt(wh) / 224)), clamped: 224^2 boxes map to level 4, 112^2 to 3,
    448^2 to 5, anything smaller than 112 to 2; same answers as torchvision's LevelMapper."""
    import torch
    from mobulaop_b200 import fpn
    sizes = np.array([20, 111.9, 112, 223.9, 224, 447.9, 448, 2000], np.float32)
    rois = np.stack([np.zeros_like(sizes), np.zeros_like(sizes), np.zeros_like(sizes), sizes, sizes], 1)
    want = [2, 2, 3, 3, 4, 4, 5, 5]
    assert fpn.assign_levels(rois, 2, 5).tolist() == want
    assert fpn.assign_levels(torch.from_numpy(rois), 2, 5).tolist() == want
    try:
        from torchvision.ops.poolers import LevelMapper
    except Exception:
        LevelMapper = None
    if LevelMapper is not None:
        rng = np.random.default_rng(0)
        xy = rng.uniform(0, 600, (500, 2)).astype(np.float32)
        wh = np.exp(rng.uniform(np.log(4), np.log(900), (500, 2))).astype(np.float32)
        boxes = np.concatenate([xy, xy + wh], 1)
        tv = LevelMapper(2, 5)([torch.from_numpy(boxes)]) + 2
        mine = fpn.assign_levels(np.concatenate([np.zeros((500, 1), np.float32), boxes], 1), 2, 5)
        assert (tv.numpy() == mine).mean() > 0.995          # ties at exact powers of two aside
    parts = fpn.split_by_level(np.array([2, 5, 3, 2, 5]), 2, 5)
    assert [p.tolist() for p in parts] == [[0, 3], [2], [], [1, 4]]


class _RecordingLib(object):
    """Stands in for the native library: records the arguments of every entry point called."""

    def __init__(self):
        self.calls = []

    def __getattr__(self, name):
        def entry(*args):
            self.calls.append((name, tuple((a.value or 0) if hasattr(a, 'value') else a for a in args)))   # c_void_p(0).value is None
            return 0
        return entry


def test_torch_fast_dispatch_passes_the_same_arguments_as_the_general_path(mobula, monkeypatch):
    """`mobula.func.roi_align_*` on three contiguous torch tensors skips the adapter objects; the
    native call it makes must be the one the general path makes, and everything the fast path does
    not cover (copies to make, other scalar types, errors) must still reach the general path."""
    import torch
    from mobulaop_b200 import _native, func as F
    mobula.op.load('ROIAlign')
    lib = _RecordingLib()
    monkeypatch.setattr(_native, '_lib', lib)
    monkeypatch.setattr(_native, 'load', lambda: lib)
    N, C, H, W, R, PH, PW = 2, 3, 9, 11, 5, 2, 4
    x = torch.zeros(N, C, H, W)
    rois = torch.zeros(R, 5)
    y = torch.zeros(R, C, PH, PW)
    dx = torch.zeros_like(x)
    n = y.numel()

    def both(call):
        """the call through the fast path and with the fast path switched off"""
        lib.calls.clear()
        call()
        fast = list(lib.calls)
        lib.calls.clear()
        saved = {}
        for fn in (F.roi_align_forward, F.roi_align_backward):
            saved[fn] = fn._fast
            fn._fast = None
        try:
            call()
        finally:
            for fn, f in saved.items():
                fn._fast = f
        general = list(lib.calls)
        lib.calls.clear()
        return fast, general

    for kw in ({}, {'accumulate': True}, {'aligned': True}, {'layout': 'RHWC'}, {'algo': 'sweep'}):
        fast, general = both(lambda: F.roi_align_forward(n, x, 0.25, C, H, W, PH, PW, 2, rois, y, **kw))
        assert len(fast) == 1 and fast == general, (kw, fast, general)
    for kw in ({}, {'accumulate': False}, {'mode': 'deterministic'}, {'aligned': True}, {'layout': 'RHWC'}):
        fast, general = both(lambda: F.roi_align_backward(n, y, 0.25, C, H, W, PH, PW, 2, dx, rois, **kw))
        assert len(fast) == 1 and fast == general, (kw, fast, general)
    # a half-precision pooled tensor is a typed top, not a type error
    fast, general = both(lambda: F.roi_align_forward(n, x, 0.25, C, H, W, PH, PW, 2, rois, y.half()))
    assert fast == general and fast[0][1][5] == _native.DTYPE_F16
    # the fast path declines: a non-contiguous tensor (the general path copies and writes back) ...
    calls = []
    monkeypatch.setattr(F.roi_align_forward, '_fast',
                        lambda *a: calls.append(F._fast_torch_forward(*a)) or calls[-1])
    yt = torch.zeros(R, C, PW, PH).transpose(2, 3)
    F.roi_align_forward(n, x, 0.25, C, H, W, PH, PW, 2, rois, yt)
    assert calls == [False] and len(lib.calls) == 1
    # ... and the errors are still the general path's
    with pytest.raises(TypeError):
        F.roi_align_forward(n, x.double(), 0.25, C, H, W, PH, PW, 2, rois, y)
    with pytest.raises(ValueError):
        F.roi_align_backward(n, y, 0.25, C, H, W, PH, PW, 2, dx, rois, mode='fast')
    with pytest.raises(ValueError):
        F.roi_align_forward(n, x, 0.25, C, H, W, PH, PW, 2, rois, y, layout='NHWC')
