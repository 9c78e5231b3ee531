"""Parity of the CUDA path against the oracle -- runs on the B200 box (`-m gpu`).

All compute goes through the C-ABI of libmobula_roialign_sm100.so: directly with ctypes
(v2 entry points and the reference's drop-in symbols) and through the mobula API.
Gates (BASELINE.json): forward within 2 ULP or 1e-6 relative (the kernels are in fact
bit-exact, which is asserted where it holds by construction), backward within 1e-5
relative to the summand magnitude in atomic mode and bit-exact in deterministic mode,
ROI bookkeeping bit-exact.
"""
import ctypes

import numpy as np
import pytest

from conftest import golden_names, load_golden, has_cuda
from mobulaop_b200 import workloads as WL
from mobulaop_b200.testing import (assert_bit_exact, assert_scatter_close,
                                   assert_within_ulp_or_rel)

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not has_cuda(), reason='needs a CUDA device')]

ALGOS = ['gather', 'plane', 'resident', 'sweep', 'pull', 'auto']


def _skip_unless_supported(algo, pooled, sr, backward=False, width=None):
    """The resident kernels exist for a fixed sampling ratio (forward 1..4, backward 2) and
    bounded pooled shapes, the sweep kernels for pooled widths up to 16; asking for them
    explicitly elsewhere is an error by design."""
    if algo == 'sweep':
        if pooled[1] > 16:
            pytest.skip('sweep kernels: pooled width <= 16 only')
        return
    if algo == 'pull':
        if not backward:
            pytest.skip('pull is a backward formulation (as a forward algorithm it means auto)')
        if max(pooled) * max(sr, 1) > 64:
            pytest.skip('pull kernels: at most 64 samples per ROI axis at a fixed sampling ratio, pooled sizes <= 64')
        return
    if algo != 'resident':
        return
    if backward:
        ok = sr == 2 and pooled[0] <= 16 and pooled[1] <= 16
    else:
        ok = 1 <= sr <= 4 and (sr == 2 or (pooled[0] + pooled[1]) * sr <= 64)
    if not ok:
        pytest.skip('resident kernels: fixed sampling ratio only')


@pytest.fixture(scope='module')
def torch():
    import torch
    return torch


@pytest.fixture(scope='module')
def api(native):
    import mobula
    mobula.op.load('ROIAlign')
    return mobula


class Direct(object):
    """v2 C-ABI with device pointers held by torch tensors."""

    def __init__(self, lib, torch, device=0):
        self.lib, self.torch = lib, torch
        self.device = device
        self.dev = torch.device('cuda', device)

    def _stream(self):
        return ctypes.c_void_p(self.torch.cuda.current_stream(self.dev).cuda_stream)

    @staticmethod
    def _p(t):
        return ctypes.c_void_p(t.data_ptr())

    def forward(self, data, rois, pooled, scale, sr, algo='auto', accumulate_into=None, no_tables=False):
        from mobulaop_b200 import _native
        t = self.torch
        d = t.from_numpy(data).to(self.dev)
        r = t.from_numpy(rois).to(self.dev)
        N, C, H, W = data.shape
        R = rois.shape[0]
        if accumulate_into is None:
            out = t.full((R, C, pooled[0], pooled[1]), float('nan'), device=self.dev)
        else:
            out = t.from_numpy(accumulate_into).to(self.dev)
        flags = _native.make_flags(accumulate=accumulate_into is not None, algo=algo, no_tables=no_tables)
        _native.check(self.lib.mobula_roialign_forward_f32(
            self.device, self._stream(), self._p(d), self._p(r), self._p(out), R, N, C, H, W,
            pooled[0], pooled[1], scale, sr, flags))
        t.cuda.synchronize(self.dev)
        return out.cpu().numpy()

    def backward(self, dy, rois, shape, scale, sr, mode='atomic', algo='auto', accumulate_into=None,
                 no_tables=False):
        from mobulaop_b200 import _native
        t = self.torch
        g = t.from_numpy(dy).to(self.dev)
        r = t.from_numpy(rois).to(self.dev)
        N, C, H, W = shape
        R, _, PH, PW = dy.shape
        if accumulate_into is None:
            dx = t.full(shape, float('nan'), device=self.dev)    # the zero-fill is the op's job
        else:
            dx = t.from_numpy(accumulate_into).to(self.dev)
        flags = _native.make_flags(accumulate=accumulate_into is not None, algo=algo, no_tables=no_tables)
        m = _native.BWD_DETERMINISTIC if mode == 'deterministic' else _native.BWD_ATOMIC
        _native.check(self.lib.mobula_roialign_backward_f32(
            self.device, self._stream(), self._p(g), self._p(r), self._p(dx), R, N, C, H, W, PH, PW,
            scale, sr, m, flags))
        t.cuda.synchronize(self.dev)
        return dx.cpu().numpy()


@pytest.fixture(scope='module')
def direct(native, torch):
    return Direct(native, torch)


def _abs_sum(oracle, g):
    return oracle.backward(np.abs(g['dy']), g['rois'], g['data'].shape, g['scale'], g['sr'])


# ---- golden vectors produced by the unmodified reference -------------------------------
@pytest.mark.parametrize('algo', ALGOS)
@pytest.mark.parametrize('name', golden_names())
def test_forward_golden(direct, name, algo):
    g = load_golden(name)
    _skip_unless_supported(algo, g['pooled'], g['sr'], width=g['data'].shape[3])
    y = direct.forward(g['data'], g['rois'], g['pooled'], g['scale'], g['sr'], algo=algo)
    assert_within_ulp_or_rel(y, g['y'], ulp=2, rtol=1e-6)
    assert_bit_exact(y, g['y'])         # same operations in the same order as the reference


@pytest.mark.parametrize('algo', ALGOS)
@pytest.mark.parametrize('name', golden_names())
def test_backward_atomic_golden(direct, oracle, name, algo):
    g = load_golden(name)
    _skip_unless_supported(algo, g['pooled'], g['sr'], backward=True, width=g['data'].shape[3])
    dx = direct.backward(g['dy'], g['rois'], g['data'].shape, g['scale'], g['sr'], algo=algo)
    assert_scatter_close(dx, g['dx'], _abs_sum(oracle, g), rtol=1e-5)


@pytest.mark.parametrize('name', golden_names())
def test_backward_deterministic_golden_bit_exact(direct, name):
    g = load_golden(name)
    dx = direct.backward(g['dy'], g['rois'], g['data'].shape, g['scale'], g['sr'], mode='deterministic')
    assert_bit_exact(dx, g['dx'])
    again = direct.backward(g['dy'], g['rois'], g['data'].shape, g['scale'], g['sr'], mode='deterministic')
    assert_bit_exact(dx, again)
    # sampling_ratio 2 takes the resident kernel's canonical-order variant, anything else the
    # whole-plane canonical kernel; both must give the single-thread reference's bits
    from mobulaop_b200 import _native
    # sampling_ratio 2 takes the resident kernel's canonical-order variant, anything else the
    # row-tile kernel's; both must give the single-thread reference's bits (as must the old
    # whole-plane canonical kernel, below)
    # fixed sampling ratios take the pull kernel's canonical-order variant, adaptive grids the
    # row-tile kernel's; the resident / whole-plane canonical kernels must give the same bits
    N, _, H, W = g['data'].shape
    if N * H * W >= 32768:
        expect = 'pull-canonical'
    else:
        expect = 'resident-canonical' if g['sr'] == 2 and max(g['pooled']) <= 16 else 'tile-canonical'
    assert _native.last_algo(True) == expect
    if max(g['pooled']) * max(g['sr'], 1) <= 64:
        pull = direct.backward(g['dy'], g['rois'], g['data'].shape, g['scale'], g['sr'], mode='deterministic',
                               algo='pull')
        assert _native.last_algo(True) == 'pull-canonical'
        assert_bit_exact(pull, g['dx'])
    tile = direct.backward(g['dy'], g['rois'], g['data'].shape, g['scale'], g['sr'], mode='deterministic',
                           algo='sweep')
    assert _native.last_algo(True) == 'tile-canonical'
    assert_bit_exact(tile, g['dx'])
    plane = direct.backward(g['dy'], g['rois'], g['data'].shape, g['scale'], g['sr'], mode='deterministic',
                            algo='plane')
    assert _native.last_algo(True) == 'plane-canonical'
    assert_bit_exact(plane, g['dx'])


# ---- seeded workloads shaped like the BASELINE configs, oracle-sized -------------------
def _sample(key, **kw):
    return WL.CONFIGS[key].scaled(**kw)


SAMPLES = {
    'C2': lambda: _sample('C2', C=6, rois_per_image=96),
    'C3': lambda: _sample('C3', N=3, C=4, rois_per_image=40),
    'C5': lambda: _sample('C5', C=3, rois_per_image=150),
    'C1': lambda: _sample('C1', C=24),
}


@pytest.mark.parametrize('algo', ALGOS)
@pytest.mark.parametrize('key', sorted(SAMPLES))
def test_seeded_workload_parity(direct, oracle, key, algo):
    wl = SAMPLES[key]()
    _skip_unless_supported(algo, wl.pooled, wl.sampling_ratio, width=wl.W)
    data, rois, dy = WL.make_inputs(wl, seed=17)
    y_ref = oracle.forward(data, rois, wl.pooled, wl.scale, wl.sampling_ratio, threads=0)
    y = direct.forward(data, rois, wl.pooled, wl.scale, wl.sampling_ratio, algo=algo)
    assert_bit_exact(y, y_ref)
    dx_ref = oracle.backward(dy, rois, data.shape, wl.scale, wl.sampling_ratio)
    dx_abs = oracle.backward(np.abs(dy), rois, data.shape, wl.scale, wl.sampling_ratio)
    if not (algo == 'resident' and wl.sampling_ratio != 2):
        dx = direct.backward(dy, rois, data.shape, wl.scale, wl.sampling_ratio, algo=algo)
        assert_scatter_close(dx, dx_ref, dx_abs, rtol=1e-5)
    dxd = direct.backward(dy, rois, data.shape, wl.scale, wl.sampling_ratio, mode='deterministic')
    assert_bit_exact(dxd, dx_ref)


@pytest.mark.parametrize('name', ['edges_sr0', 'edges_sr2', 'edges_sr3_p5x3', 'ref_value_test'])
def test_untabled_rois_take_the_on_the_fly_path(direct, name):
    """ROIs that do not fit the table arena are decoded inside the kernels; forced here."""
    g = load_golden(name)
    y = direct.forward(g['data'], g['rois'], g['pooled'], g['scale'], g['sr'], algo='plane', no_tables=True)
    assert_bit_exact(y, g['y'])
    dx = direct.backward(g['dy'], g['rois'], g['data'].shape, g['scale'], g['sr'], mode='deterministic',
                         no_tables=True)
    assert_bit_exact(dx, g['dx'])


# ---- odd shapes: everything the kernels special-case ---------------------------------------------
ODD_SHAPES = [
    # N, C, H, W, rois/img, pooled, scale, sr
    (3, 5, 21, 30, 17, (7, 7), 0.25, 2),       # odd C (half a channel pair), W % 4 != 0
    (2, 3, 40, 52, 23, (16, 16), 0.125, 2),    # widest pooled shape of the sr = 2 kernels
    (2, 4, 33, 36, 19, (1, 1), 0.5, 2),        # a single bin
    (2, 2, 18, 20, 9, (9, 3), 0.5, 2),         # PH > 8, PW small: mixed table sizes
    (1, 7, 150, 200, 40, (7, 7), 0.25, 2),     # tall plane: many row bands per plane
    (1, 3, 37, 271, 30, (7, 7), 0.25, 2),      # wide plane, odd width: two column tiles, the second ragged
    (1, 2, 30, 300, 25, (14, 14), 0.25, 2),    # two column tiles with two chunks of column records
    (1, 2, 150, 200, 30, (5, 12), 0.25, 2),    # odd pooled height on a single-plane kernel: two ROIs per warp,
    (1, 2, 150, 200, 30, (3, 16), 0.25, 2),    # ... two chunks of sample columns
    (1, 3, 150, 200, 30, (1, 3), 0.25, 2),     # ... one bin row
    (2, 3, 24, 28, 11, (7, 7), 0.25, 1),       # generic resident forward, sr 1
    (2, 3, 24, 28, 11, (5, 6), 0.25, 4),       # ... sr 4
]


@pytest.mark.parametrize('shape', ODD_SHAPES, ids=lambda s: 'x'.join(str(v) for v in s[:5]) + '-p%dx%d-sr%d' % (s[5] + (s[7],)))
def test_odd_shapes_against_the_oracle(direct, oracle, shape):
    N, C, H, W, k, pooled, scale, sr = shape
    wl = WL.Workload('odd', N, C, H, W, k, pooled, scale, sr, min_box=2.0, max_box=4.0 * max(H, W) / scale)
    data, rois, dy = WL.make_inputs(wl, seed=23)
    rng = np.random.default_rng(5)
    perm = rng.permutation(len(rois))                  # ROIs not grouped by image
    rois = np.ascontiguousarray(rois[perm])
    dy = np.ascontiguousarray(dy[perm])
    rois[rois[:, 0] == N - 1, 0] = 0 if N > 1 else 0   # the last image has no ROI at all
    y_ref = oracle.forward(data, rois, pooled, scale, sr, threads=0)
    dx_ref = oracle.backward(dy, rois, data.shape, scale, sr)
    dx_abs = oracle.backward(np.abs(dy), rois, data.shape, scale, sr)
    for algo in ('auto', 'gather'):
        assert_bit_exact(direct.forward(data, rois, pooled, scale, sr, algo=algo), y_ref)
        assert_scatter_close(direct.backward(dy, rois, data.shape, scale, sr, algo=algo), dx_ref, dx_abs,
                             rtol=1e-5)
    assert_bit_exact(direct.backward(dy, rois, data.shape, scale, sr, mode='deterministic'), dx_ref)
    base = rng.standard_normal(data.shape).astype(np.float32)
    acc = direct.backward(dy, rois, data.shape, scale, sr, accumulate_into=base.copy())
    np.testing.assert_allclose(acc, base + dx_ref, rtol=1e-5, atol=1e-5)


# ---- sweep kernels (lane = channel, TMA row ring): shapes that exercise their special cases --------
SWEEP_SHAPES = [
    # N, C, H, W, rois/img, pooled, scale, sr
    (3, 5, 21, 32, 17, (7, 7), 0.25, 2),        # one ragged channel group
    (2, 40, 40, 52, 23, (16, 16), 0.125, 2),    # two channel groups (32 + 8), widest pooled shape
    (2, 4, 33, 36, 19, (1, 1), 0.5, 2),         # a single bin
    (2, 2, 18, 20, 9, (9, 3), 0.5, 2),          # PH > PW
    (1, 7, 150, 200, 40, (7, 7), 0.25, 2),      # tall plane: several row segments per image, ring wraps often
    (1, 3, 37, 272, 30, (7, 7), 0.25, 2),       # wider than one TMA box: two boxes per channel group
    (1, 2, 30, 300, 25, (14, 14), 0.25, 2),     # ... with the mask-head pooling
    (1, 33, 150, 200, 30, (5, 12), 0.25, 0),    # adaptive grid, compacted column records, two channel groups
    (2, 3, 24, 28, 11, (7, 7), 0.25, 1),        # sampling ratio 1 / 4 / 3: compacted records, fixed grid
    (2, 3, 24, 28, 11, (5, 6), 0.25, 4),
    (2, 3, 64, 48, 31, (6, 5), 0.25, 3),
    (2, 64, 200, 272, 48, (7, 7), 0.25, 2),     # the box-head plane: six-row ring, chained items
    (1, 3, 60, 400, 30, (7, 7), 0.25, 0),       # 400 wide: four-row ring
    (2, 3, 50, 68, 600, (7, 7), 0.25, 0),       # many ROIs per image, adaptive
    (3, 5, 21, 30, 17, (7, 7), 0.25, 2),        # width % 4 != 0: rows not 16-byte aligned, 4-byte fill
    (1, 3, 37, 271, 30, (7, 7), 0.25, 0),       # ... odd width, adaptive
    (2, 8, 14, 14, 64, (7, 7), 1.0, 2),         # the reference benchmark's plane
]


@pytest.mark.parametrize('shape', SWEEP_SHAPES, ids=lambda s: 'x'.join(str(v) for v in s[:5]) + '-p%dx%d-sr%d' % (s[5] + (s[7],)))
def test_sweep_family_against_the_oracle(direct, oracle, shape):
    from mobulaop_b200 import _native
    N, C, H, W, k, pooled, scale, sr = shape
    wl = WL.Workload('sweep', N, C, H, W, k, pooled, scale, sr, min_box=2.0, max_box=4.0 * max(H, W) / scale,
                     stress=sr == 0)
    data, rois, _ = WL.make_inputs(wl, seed=31, with_dy=False)
    rng = np.random.default_rng(6)
    perm = rng.permutation(len(rois))                  # ROIs not grouped by image
    rois = np.ascontiguousarray(rois[perm])
    if N > 2:
        rois[rois[:, 0] == N - 1, 0] = 0               # the last image has no ROI at all
    rois[1, 0] = N + 2                                 # points at no image: zero rows
    y_ref = oracle.forward(data, np.delete(rois, 1, axis=0), pooled, scale, sr, threads=0)
    y = direct.forward(data, rois, pooled, scale, sr, algo='sweep')
    assert _native.last_algo(False) == 'sweep'
    assert not y[1].any()
    assert_bit_exact(np.delete(y, 1, axis=0), y_ref)
    base = rng.standard_normal(y.shape).astype(np.float32)
    acc = direct.forward(data, rois, pooled, scale, sr, algo='sweep', accumulate_into=base.copy())
    expect = base.copy()
    expect[np.arange(len(rois)) != 1] += y_ref
    assert_bit_exact(acc, expect)
    # the backward of the family: dX row tiles, strip-owner warps
    dy = rng.standard_normal(y.shape).astype(np.float32)
    dyg = np.ascontiguousarray(np.delete(dy, 1, axis=0))
    good = np.delete(rois, 1, axis=0)
    dx_ref = oracle.backward(dyg, good, data.shape, scale, sr)
    dx_abs = oracle.backward(np.abs(dyg), good, data.shape, scale, sr)
    dx = direct.backward(dy, rois, data.shape, scale, sr, algo='sweep')
    assert _native.last_algo(True) == 'tile'
    assert_scatter_close(dx, dx_ref, dx_abs, rtol=1e-5)
    dxd = direct.backward(dy, rois, data.shape, scale, sr, algo='sweep', mode='deterministic')
    assert _native.last_algo(True) == 'tile-canonical'
    assert_bit_exact(dxd, dx_ref)
    dbase = rng.standard_normal(data.shape).astype(np.float32)
    dacc = direct.backward(dy, rois, data.shape, scale, sr, algo='sweep', mode='deterministic',
                           accumulate_into=dbase.copy())
    assert_bit_exact(dacc, (dbase + dx_ref).astype(np.float32))


PULL_SHAPES = [
    # N, C, H, W, rois/img, pooled, scale, sr
    (3, 5, 21, 32, 17, (7, 7), 0.25, 2),        # one channel per lane, ragged group
    (2, 40, 40, 52, 23, (16, 16), 0.125, 2),    # 32 samples per axis
    (2, 4, 33, 36, 19, (1, 1), 0.5, 2),         # a single bin
    (2, 2, 18, 20, 9, (9, 3), 0.5, 2),          # PH > PW
    (1, 70, 150, 200, 40, (7, 7), 0.25, 2),     # two channels per lane (72 padded), ragged
    (1, 3, 37, 271, 30, (7, 7), 0.25, 2),       # odd width: ragged last 32-pixel segment, unaligned rows
    (1, 2, 31, 300, 25, (14, 14), 0.25, 2),     # odd height: the last row band is half empty
    (2, 3, 24, 28, 11, (7, 7), 0.25, 1),        # sampling ratio 1 / 4 / 3 (count 9: true division)
    (2, 3, 24, 28, 11, (5, 6), 0.25, 4),
    (2, 3, 64, 48, 31, (6, 5), 0.25, 3),
    (2, 130, 50, 68, 48, (7, 7), 0.25, 2),      # four channels per lane, two groups, the second ragged (2 of 128)
    (2, 3, 50, 68, 600, (7, 7), 0.25, 2),       # many ROIs per image: long lists, several mask words
    (2, 8, 14, 14, 64, (7, 7), 1.0, 2),         # the reference benchmark's plane: every sample clamps or overlaps
    (5, 3, 9, 11, 3, (2, 2), 1.0, 2),           # tiny
    (1, 33, 150, 200, 30, (5, 12), 0.25, 0),    # adaptive grids: per-ROI sample counts, degenerate and outside boxes
    (2, 3, 50, 68, 600, (7, 7), 0.25, 0),       # ... many ROIs per image
    (1, 3, 37, 271, 30, (7, 7), 0.25, 0),       # ... odd width
    (2, 130, 40, 52, 40, (7, 7), 0.25, 0),      # ... four channels per lane
]


@pytest.mark.parametrize('shape', PULL_SHAPES, ids=lambda s: 'x'.join(str(v) for v in s[:5]) + '-p%dx%d-sr%d' % (s[5] + (s[7],)))
def test_pull_backward_against_the_oracle(direct, oracle, shape):
    """The per-pixel-list backward: merged taps within 1e-5, canonical variant bit-exact, both
    reproducible; unsorted ROIs, an image without ROIs, a ROI that points at no image."""
    from mobulaop_b200 import _native
    N, C, H, W, k, pooled, scale, sr = shape
    wl = WL.Workload('pull', N, C, H, W, k, pooled, scale, sr, min_box=2.0, max_box=4.0 * max(H, W) / scale,
                     stress=sr == 0)
    data, rois, dy = WL.make_inputs(wl, seed=37)
    rng = np.random.default_rng(8)
    perm = rng.permutation(len(rois))
    rois = np.ascontiguousarray(rois[perm])
    if N > 2:
        rois[rois[:, 0] == N - 1, 0] = 0
    rois[1, 0] = N + 2
    dyg = np.ascontiguousarray(np.delete(dy, 1, axis=0))
    good = np.delete(rois, 1, axis=0)
    dx_ref = oracle.backward(dyg, good, data.shape, scale, sr)
    dx_abs = oracle.backward(np.abs(dyg), good, data.shape, scale, sr)
    dx = direct.backward(dy, rois, data.shape, scale, sr, algo='pull')
    assert _native.last_algo(True) == 'pull'
    assert_scatter_close(dx, dx_ref, dx_abs, rtol=1e-5)
    assert_bit_exact(direct.backward(dy, rois, data.shape, scale, sr, algo='pull'), dx)
    dxd = direct.backward(dy, rois, data.shape, scale, sr, algo='pull', mode='deterministic')
    assert _native.last_algo(True) == 'pull-canonical'
    assert_bit_exact(dxd, dx_ref)
    dbase = rng.standard_normal(data.shape).astype(np.float32)
    dacc = direct.backward(dy, rois, data.shape, scale, sr, algo='pull', mode='deterministic',
                           accumulate_into=dbase.copy())
    assert_bit_exact(dacc, (dbase + dx_ref).astype(np.float32))


def test_pull_rejects_what_it_cannot_do(direct):
    rois = np.array([[0, 1, 1, 9, 9]], np.float32)
    with pytest.raises(NotImplementedError):            # 70 samples per ROI axis at a fixed ratio
        direct.backward(np.zeros((1, 2, 35, 2), np.float32), rois, (1, 2, 8, 20), 1.0, 2, algo='pull')


def test_sweep_rejects_what_it_cannot_do(direct):
    rois = np.array([[0, 1, 1, 9, 9]], np.float32)
    with pytest.raises(NotImplementedError):            # pooled width beyond the register tile
        direct.forward(np.zeros((1, 2, 8, 20), np.float32), rois, (2, 17), 1.0, 2, algo='sweep')


# ---- ROI bookkeeping --------------------------------------------------------------------
@pytest.mark.parametrize('name', ['edges_sr0', 'edges_sr2', 'ref_value_test', 'fpn_14x14_sr2'])
def test_bookkeeping_bit_exact(native, torch, oracle, name):
    from mobulaop_b200 import _native
    g = load_golden(name)
    H, W = g['data'].shape[2:]
    max_grid = 16
    gi, gf, ti, tf = oracle.bookkeeping(g['rois'], (H, W), g['pooled'], g['scale'], g['sr'], max_grid)
    dev = torch.device('cuda', 0)
    R = len(g['rois'])
    d_gi = torch.zeros(gi.shape, dtype=torch.int32, device=dev)
    d_gf = torch.zeros(gf.shape, dtype=torch.float32, device=dev)
    d_ti = torch.full(ti.shape, -7, dtype=torch.int32, device=dev)
    d_tf = torch.full(tf.shape, float('nan'), dtype=torch.float32, device=dev)
    r = torch.from_numpy(g['rois']).to(dev)
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    _native.check(native.mobula_roialign_bookkeeping_f32(
        0, ctypes.c_void_p(0), p(r), R, H, W, g['pooled'][0], g['pooled'][1], g['scale'], g['sr'],
        p(d_gi), p(d_gf), max_grid, p(d_ti), p(d_tf)))
    torch.cuda.synchronize()
    np.testing.assert_array_equal(d_gi.cpu().numpy(), gi)
    assert_bit_exact(d_gf.cpu().numpy(), gf)
    np.testing.assert_array_equal(d_ti.cpu().numpy(), ti)
    assert_bit_exact(d_tf.cpu().numpy(), tf)


# ---- edge cases ---------------------------------------------------------------------------
def test_empty_and_tiny_inputs(direct, oracle):
    data = np.random.default_rng(0).standard_normal((2, 3, 9, 12)).astype(np.float32)
    none = np.zeros((0, 5), np.float32)
    assert direct.forward(data, none, (7, 7), 0.5, 2).shape == (0, 3, 7, 7)
    dx = direct.backward(np.zeros((0, 3, 7, 7), np.float32), none, data.shape, 0.5, 2)
    assert dx.shape == data.shape and not dx.any()          # still zero-fills dX
    dxd = direct.backward(np.zeros((0, 3, 7, 7), np.float32), none, data.shape, 0.5, 2,
                          mode='deterministic')
    assert not dxd.any()
    one = np.array([[1, 2.5, 1.5, 14.0, 12.0]], np.float32)
    for algo in ALGOS:
        y = direct.forward(data[:, :1], one, (1, 1), 0.5, 1, algo=algo)
        assert_bit_exact(y, oracle.forward(np.ascontiguousarray(data[:, :1]), one, (1, 1), 0.5, 1))


@pytest.mark.parametrize('algo', ALGOS)
def test_batch_index_out_of_range_is_ignored(direct, oracle, algo):
    """With N known, ROIs pointing outside [0, N) read nothing and scatter nothing
    (the reference would read out of bounds, ROIAlign.cpp:43-44)."""
    rng = np.random.default_rng(1)
    data = rng.standard_normal((2, 2, 10, 12)).astype(np.float32)
    rois = np.array([[0, 1, 1, 8, 8], [2, 1, 1, 8, 8], [-1, 0, 0, 5, 5], [1, 0, 0, 9, 9]], np.float32)
    dy = rng.standard_normal((4, 2, 2, 2)).astype(np.float32)
    y = direct.forward(data, rois, (2, 2), 1.0, 2, algo=algo)
    good = np.array([0, 3])
    assert_bit_exact(y[good], oracle.forward(data, rois[good], (2, 2), 1.0, 2))
    assert not y[[1, 2]].any()
    for mode in ('atomic', 'deterministic'):
        if mode == 'deterministic' and algo == 'gather':
            continue                     # global atomics cannot be deterministic
        dx = direct.backward(dy, rois, data.shape, 1.0, 2, mode=mode, algo=algo)
        ref = oracle.backward(np.ascontiguousarray(dy[good]), rois[good], data.shape, 1.0, 2)
        np.testing.assert_allclose(dx, ref, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize('algo', ALGOS)
def test_accumulate_flags(direct, oracle, algo):
    g = load_golden('edges_w20_sr2' if algo == 'sweep' else 'edges_sr2')
    base = np.random.default_rng(2).standard_normal(g['y'].shape).astype(np.float32)
    y = direct.forward(g['data'], g['rois'], g['pooled'], g['scale'], g['sr'], algo=algo,
                       accumulate_into=base)
    assert_bit_exact(y, (base + g['y']).astype(np.float32))
    dbase = np.random.default_rng(3).standard_normal(g['dx'].shape).astype(np.float32)
    dx = direct.backward(g['dy'], g['rois'], g['data'].shape, g['scale'], g['sr'], algo=algo,
                         accumulate_into=dbase)
    np.testing.assert_allclose(dx, dbase + g['dx'], rtol=1e-5, atol=1e-5)


# ---- the reference's drop-in symbols ---------------------------------------------------------
def test_dropin_symbols_device_pointers(native, torch, oracle):
    g = load_golden('fpn_14x14_sr2')
    dev = torch.device('cuda', 0)
    N, C, H, W = g['data'].shape
    R = len(g['rois'])
    PH, PW = g['pooled']
    d = torch.from_numpy(g['data']).to(dev)
    r = torch.from_numpy(g['rois']).to(dev)
    out = torch.empty(g['y'].shape, device=dev)
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    i32, f32 = ctypes.c_int, ctypes.c_float
    torch.cuda.synchronize()
    # exactly how mobula/func.py:155-156 calls it: explicit ctypes instances, device_id first
    native.roi_align_forward_ab78de3c(i32(0), i32(out.numel()), p(d), f32(g['scale']), i32(C), i32(H),
                                      i32(W), i32(PH), i32(PW), i32(g['sr']), p(r), p(out))
    assert_bit_exact(out.cpu().numpy(), g['y'])       # synchronous: no explicit sync needed
    dy = torch.from_numpy(g['dy']).to(dev)
    dx = torch.zeros(g['dx'].shape, device=dev)       # caller zero-fills (ROIAlign.py:33-34)
    torch.cuda.synchronize()
    native.roi_align_backward_f318a24e(i32(0), i32(dy.numel()), p(dy), f32(g['scale']), i32(C), i32(H),
                                       i32(W), i32(PH), i32(PW), i32(g['sr']), p(dx), p(r))
    assert_scatter_close(dx.cpu().numpy(), g['dx'], _abs_sum(oracle, g), rtol=1e-5)
    native.set_device(i32(0))


def test_dropin_symbols_host_pointers(native, oracle):
    """device_id = -1 is the reference's CPU context (func.py:139-141): host pointers."""
    g = load_golden('edges_sr0')
    N, C, H, W = g['data'].shape
    PH, PW = g['pooled']
    out = np.full(g['y'].shape, np.nan, np.float32)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    i32, f32 = ctypes.c_int, ctypes.c_float
    native.roi_align_forward_ab78de3c(i32(-1), i32(out.size), p(g['data']), f32(g['scale']), i32(C),
                                      i32(H), i32(W), i32(PH), i32(PW), i32(g['sr']), p(g['rois']), p(out))
    assert_bit_exact(out, g['y'])
    dx = np.zeros(g['dx'].shape, np.float32)
    native.roi_align_backward_f318a24e(i32(-1), i32(g['dy'].size), p(g['dy']), f32(g['scale']), i32(C),
                                       i32(H), i32(W), i32(PH), i32(PW), i32(g['sr']), p(dx), p(g['rois']))
    assert_scatter_close(dx, g['dx'], _abs_sum(oracle, g), rtol=1e-5)


# ---- through the mobula API -----------------------------------------------------------------
@pytest.mark.parametrize('deterministic', [False, True])
def test_torch_operator_forward_backward(api, torch, oracle, deterministic):
    g = load_golden('fpn_14x14_sr2')
    dev = torch.device('cuda', 0)
    x = torch.from_numpy(g['data']).to(dev).requires_grad_(True)
    rois = torch.from_numpy(g['rois']).to(dev).requires_grad_(True)
    y = api.op.ROIAlign(x, rois, pooled_size=g['pooled'], spatial_scale=g['scale'],
                        sampling_ratio=g['sr'], deterministic=deterministic)
    assert_bit_exact(y.detach().cpu().numpy(), g['y'])
    y.backward(torch.from_numpy(g['dy']).to(dev))
    if deterministic:
        assert_bit_exact(x.grad.cpu().numpy(), g['dx'])
    else:
        assert_scatter_close(x.grad.cpu().numpy(), g['dx'], _abs_sum(oracle, g), rtol=1e-5)
    assert rois.grad is not None and not rois.grad.any().item()     # ROIAlign.py:41-42
    # the reference's own tolerance (test_roialign.py:278-282)
    api.testing.assert_almost_equal(y, g['y'], rtol=1e-3, atol=1e-3)


def test_torch_module_form_noncontiguous_and_accumulating_grads(api, torch):
    g = load_golden('edges_sr2')
    dev = torch.device('cuda', 0)
    layer = api.op.ROIAlign[torch.Tensor](pooled_size=g['pooled'], spatial_scale=g['scale'],
                                          sampling_ratio=g['sr'])
    assert isinstance(layer, torch.nn.Module)
    nhwc = torch.from_numpy(g['data']).to(dev).permute(0, 2, 3, 1).contiguous()
    x = nhwc.permute(0, 3, 1, 2).requires_grad_(True)              # NCHW view, not contiguous
    assert not x.is_contiguous()
    rois = torch.from_numpy(g['rois']).to(dev)
    y1 = layer(x, rois)
    y2 = layer(x, rois)                                            # same layer twice before backward
    assert_bit_exact(y1.detach().cpu().numpy(), g['y'])
    (y1.sum() + y2.sum()).backward()
    ones = np.ones_like(g['dy'])
    from oracle.roialign import COracle
    ref = COracle().backward(ones, g['rois'], g['data'].shape, g['scale'], g['sr'])
    np.testing.assert_allclose(x.grad.cpu().numpy(), 2 * ref, rtol=1e-5, atol=1e-5)


def test_numpy_operator_both_call_forms(api, oracle):
    g = load_golden('ref_value_test')
    op = api.op.ROIAlign[np.ndarray](pooled_size=g['pooled'], spatial_scale=g['scale'],
                                     sampling_ratio=g['sr'])
    y = op(g['data'], g['rois'])
    assert isinstance(y, np.ndarray)
    assert_bit_exact(y, g['y'])
    dx, drois = op.backward(g['dy'])
    assert_scatter_close(dx, g['dx'], _abs_sum(oracle, g), rtol=1e-5)
    assert not drois.any()
    # req='add' accumulates into the given gradient buffers
    acc = [np.ones_like(g['data']), np.ones_like(g['rois'])]
    op.backward(g['dy'], in_grad=acc, req=['add', 'null'])
    np.testing.assert_allclose(acc[0], 1 + g['dx'], rtol=1e-5, atol=1e-5)
    assert (acc[1] == 1).all()
    # plain call form (broken in the reference, numpy_glue.py:39-41)
    y2 = api.op.ROIAlign(g['data'], g['rois'], pooled_size=g['pooled'], spatial_scale=g['scale'],
                         sampling_ratio=g['sr'])
    assert_bit_exact(y2, g['y'])
    # func-level call with a non-contiguous output is copied back (test_func.py:9-22)
    wide = np.zeros(g['y'].shape[:3] + (2 * g['y'].shape[3],), np.float32)
    view = wide[..., ::2]
    api.func.roi_align_forward(view.size, g['data'], g['scale'], g['data'].shape[1], g['data'].shape[2],
                               g['data'].shape[3], g['pooled'][0], g['pooled'][1], g['sr'], g['rois'], view)
    assert_bit_exact(np.ascontiguousarray(view), g['y'])
    assert not wide[..., 1::2].any()


def test_host_buffer_path_is_pipelined_over_channel_chunks(api, native, oracle):
    """Host tensors of 16 MiB and more are staged in channel chunks (copy-in, compute and
    copy-out overlap); the result must not depend on that."""
    from mobulaop_b200 import _native
    wl = WL.Workload('host-chunks', 2, 50, 200, 272, 40, (7, 7), 0.25, 2)     # 21.8 MB of data
    data, rois, dy = WL.make_inputs(wl, seed=5)
    rois[3, 0] = 7                                  # a batch index outside [0, N)
    good = np.array([r for r in range(len(rois)) if r != 3])
    op = api.op.ROIAlign[np.ndarray](pooled_size=wl.pooled, spatial_scale=wl.scale,
                                     sampling_ratio=wl.sampling_ratio)
    y = op(data, rois)
    y_ref = oracle.forward(data, rois[good], wl.pooled, wl.scale, wl.sampling_ratio, threads=0)
    assert_bit_exact(y[good], y_ref)
    assert not y[3].any()
    dx, _ = op.backward(dy)
    dyg = np.ascontiguousarray(dy[good])
    dx_ref = oracle.backward(dyg, rois[good], data.shape, wl.scale, wl.sampling_ratio)
    dx_abs = oracle.backward(np.abs(dyg), rois[good], data.shape, wl.scale, wl.sampling_ratio)
    assert_scatter_close(dx, dx_ref, dx_abs, rtol=1e-5)
    assert _native.last_algo(False) == 'resident' and _native.last_algo(True) == 'pull'      # 2 x 200 x 272 pixels
    # accumulating variants (req = 'add') through the v2 entry points with host pointers
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    N, C, H, W = data.shape
    R = len(rois)
    flags = _native.make_flags(accumulate=True, host=True)
    base = np.random.default_rng(6).standard_normal(y.shape).astype(np.float32)
    acc = base.copy()
    _native.check(native.mobula_roialign_forward_f32(0, None, p(data), p(rois), p(acc), R, N, C, H, W, 7, 7,
                                                     wl.scale, 2, flags))
    assert_bit_exact(acc, (base + y).astype(np.float32))
    dbase = np.random.default_rng(7).standard_normal(data.shape).astype(np.float32)
    dacc = dbase.copy()
    _native.check(native.mobula_roialign_backward_f32(0, None, p(dy), p(rois), p(dacc), R, N, C, H, W, 7, 7,
                                                      wl.scale, 2, _native.BWD_ATOMIC, flags))
    np.testing.assert_allclose(dacc, dbase + dx, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize('deterministic', [False, True])
def test_launches_can_be_captured_in_a_cuda_graph(api, torch, oracle, deterministic):
    """Device-pointer calls are stream-ordered: no synchronise, no allocation and no host
    read once the workspace exists, so forward + backward replay from a CUDA graph on new
    contents of the same buffers."""
    wl = _sample('C2', C=5, rois_per_image=64)
    dev = torch.device('cuda', 0)
    PH, PW = wl.pooled
    first, second = WL.make_inputs(wl, seed=5), WL.make_inputs(wl, seed=6)
    x = torch.from_numpy(first[0]).to(dev)
    rois = torch.from_numpy(first[1]).to(dev)
    dy = torch.from_numpy(first[2]).to(dev)
    y = torch.empty(wl.out_shape, device=dev)
    dx = torch.empty(wl.data_shape, device=dev)
    nthreads = int(y.numel())
    mode = 'deterministic' if deterministic else 'atomic'

    def step():
        api.func.roi_align_forward(nthreads, x, wl.scale, wl.C, wl.H, wl.W, PH, PW, wl.sampling_ratio,
                                   rois, y)
        api.func.roi_align_backward(nthreads, dy, wl.scale, wl.C, wl.H, wl.W, PH, PW, wl.sampling_ratio,
                                    dx, rois, mode=mode, accumulate=False)

    side = torch.cuda.Stream(dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        step()                      # sizes the workspace outside the capture
    torch.cuda.current_stream(dev).wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        step()
    x.copy_(torch.from_numpy(second[0]))
    rois.copy_(torch.from_numpy(second[1]))
    dy.copy_(torch.from_numpy(second[2]))
    y.fill_(float('nan'))
    dx.fill_(float('nan'))
    graph.replay()
    torch.cuda.synchronize(dev)
    data2, rois2, dy2 = second
    assert_bit_exact(y.cpu().numpy(), oracle.forward(data2, rois2, wl.pooled, wl.scale, wl.sampling_ratio))
    dx_ref = oracle.backward(dy2, rois2, data2.shape, wl.scale, wl.sampling_ratio)
    if deterministic:
        assert_bit_exact(dx.cpu().numpy(), dx_ref)
    else:
        dx_abs = oracle.backward(np.abs(dy2), rois2, data2.shape, wl.scale, wl.sampling_ratio)
        assert_scatter_close(dx.cpu().numpy(), dx_ref, dx_abs, rtol=1e-5)


def test_mixed_devices_rejected(api, torch):
    dev = torch.device('cuda', 0)
    data = torch.zeros((1, 1, 4, 4), device=dev)
    rois = torch.zeros((1, 5))                                     # host tensor
    with pytest.raises(ValueError):
        api.op.ROIAlign(data, rois, pooled_size=(2, 2), spatial_scale=1.0, sampling_ratio=1)


def test_same_result_on_every_gpu(direct, native, torch):
    """tests/gpu/test_gpu.py:19-43 of the reference: the op on two devices agrees."""
    if torch.cuda.device_count() < 2:
        pytest.skip('one GPU')
    g = load_golden('fpn_14x14_sr2')
    other = Direct(native, torch, device=1)
    assert_bit_exact(direct.forward(g['data'], g['rois'], g['pooled'], g['scale'], g['sr']),
                     other.forward(g['data'], g['rois'], g['pooled'], g['scale'], g['sr']))
    assert_bit_exact(
        direct.backward(g['dy'], g['rois'], g['data'].shape, g['scale'], g['sr'], mode='deterministic'),
        other.backward(g['dy'], g['rois'], g['data'].shape, g['scale'], g['sr'], mode='deterministic'))


# ---- full BASELINE sizes: size-independent properties ----------------------------------------
def _bits_equal(torch, a, b):
    return bool(torch.equal(a.view(torch.int32), b.view(torch.int32)))


@pytest.mark.parametrize('key', ['C2', 'C3', 'C4', 'C5'])
def test_full_size_properties(native, oracle, torch, key):
    """Every BASELINE configuration at its FULL size.  The tensors stay on the device (C4 is
    3.6 GB of feature maps); only channel slices travel to the host for the oracle."""
    from mobulaop_b200 import _native
    wl = WL.CONFIGS[key]
    dev = torch.device('cuda', 0)
    data_h, rois_h, dy_h = WL.make_inputs(wl, seed=0)
    if key == 'C5':
        # the stress set must contain what it is for: grids of 16+ samples per bin axis,
        # degenerate and fully-outside boxes
        gi = oracle.bookkeeping(rois_h, (wl.H, wl.W), wl.pooled, wl.scale, wl.sampling_ratio, 1)[0]
        assert gi[:, 1].max() >= 16 and gi[:, 2].max() >= 16 and gi[:, 1].min() == 1
    data = torch.from_numpy(data_h).to(dev)
    rois = torch.from_numpy(rois_h).to(dev)
    dy = torch.from_numpy(dy_h).to(dev)
    PH, PW = wl.pooled
    sp = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    p = lambda t: ctypes.c_void_p(t.data_ptr())

    def forward():
        out = torch.full(wl.out_shape, float('nan'), device=dev)
        _native.check(native.mobula_roialign_forward_f32(
            0, sp, p(data), p(rois), p(out), wl.R, wl.N, wl.C, wl.H, wl.W, PH, PW, wl.scale,
            wl.sampling_ratio, _native.make_flags()))
        return out

    def backward(g, mode):
        dx = torch.full(wl.data_shape, float('nan'), device=dev)
        _native.check(native.mobula_roialign_backward_f32(
            0, sp, p(g), p(rois), p(dx), wl.R, wl.N, wl.C, wl.H, wl.W, PH, PW, wl.scale,
            wl.sampling_ratio, mode, _native.make_flags()))
        return dx

    y = forward()
    dx = backward(dy, _native.BWD_ATOMIC)
    torch.cuda.synchronize(dev)
    # (1) a channel slice of the full-size result equals the oracle on that slice, all ROIs
    cs = [0, wl.C // 2, wl.C - 1]
    sub = np.ascontiguousarray(data_h[:, cs])
    assert_bit_exact(y[:, cs].contiguous().cpu().numpy(),
                     oracle.forward(sub, rois_h, wl.pooled, wl.scale, wl.sampling_ratio, threads=0))
    dys = np.ascontiguousarray(dy_h[:, cs])
    ref = oracle.backward(dys, rois_h, sub.shape, wl.scale, wl.sampling_ratio)
    ref_abs = oracle.backward(np.abs(dys), rois_h, sub.shape, wl.scale, wl.sampling_ratio)
    assert_scatter_close(dx[:, cs].contiguous().cpu().numpy(), ref, ref_abs, rtol=1e-5)
    # (2) adjoint identity <fwd(x), dy> == <x, bwd(dy)> ties the two kernels together
    lhs = float(torch.dot(y.double().reshape(-1), dy.double().reshape(-1)))
    rhs = float(torch.dot(data.double().reshape(-1), dx.double().reshape(-1)))
    scale = float(torch.linalg.vector_norm(y.double()) * torch.linalg.vector_norm(dy.double()))
    assert abs(lhs - rhs) <= 1e-5 * scale, (lhs, rhs, scale)
    del y, dx
    # (3) deterministic mode: bit-reproducible and equal to the canonical order
    d1 = backward(dy, _native.BWD_DETERMINISTIC)
    d2 = backward(dy, _native.BWD_DETERMINISTIC)
    assert _bits_equal(torch, d1, d2)
    assert_bit_exact(d1[:, cs].contiguous().cpu().numpy(), ref)
    del d2
    # (4) linearity of the scatter in dy (exact for power-of-two factors)
    d4 = backward(dy * 4.0, _native.BWD_DETERMINISTIC)
    assert _bits_equal(torch, d4, d1 * 4.0)


# ---- the sharded path on the devices (BASELINE config 4) ---------------------------
def test_sharded_cuda_path_matches_single_device(native, torch, oracle):
    """A C4-shaped mini problem split by image with sharding.make_shard, every shard run through
    the CUDA path on its own device (all visible GPUs; four shards in turn on one GPU when only
    one is there), reassembled, and compared bit for bit with the single-device result and the
    oracle: forward and deterministic backward; the atomic backward within its tolerance."""
    from mobulaop_b200 import sharding as S
    G = torch.cuda.device_count()
    world = G if G >= 2 else 4
    wl = WL.CONFIGS['C4'].scaled(N=2 * world + 1, C=4, rois_per_image=48)     # uneven image split
    data, rois, dy = WL.make_inputs(wl, seed=29)
    perm = np.random.default_rng(3).permutation(len(rois))                    # rows not grouped by image
    rois, dy = np.ascontiguousarray(rois[perm]), np.ascontiguousarray(dy[perm])
    rois[5, 0] = wl.N + 3                                                     # owned by nobody
    rois[9, 0] = -2
    shards = S.make_shards(rois, wl.N, world)
    ys, dxs, dxa = [], [], []
    for s in shards:
        d = Direct(native, torch, device=s.rank % G)
        local = np.ascontiguousarray(data[s.image_begin:s.image_end])
        ys.append(d.forward(local, s.rois, wl.pooled, wl.scale, wl.sampling_ratio))
        g = S.scatter_rows(s, dy)
        dxs.append(d.backward(g, s.rois, local.shape, wl.scale, wl.sampling_ratio, mode='deterministic'))
        dxa.append(d.backward(g, s.rois, local.shape, wl.scale, wl.sampling_ratio))
    y = S.assemble_rows(shards, ys, wl.R)
    dx = S.assemble_images(shards, dxs)
    single = Direct(native, torch, device=0)
    assert_bit_exact(y, single.forward(data, rois, wl.pooled, wl.scale, wl.sampling_ratio))
    assert_bit_exact(dx, single.backward(dy, rois, data.shape, wl.scale, wl.sampling_ratio,
                                         mode='deterministic'))
    good = np.array([r for r in range(wl.R) if r not in (5, 9)])
    assert not y[[5, 9]].any()
    assert_bit_exact(y[good], oracle.forward(data, rois[good], wl.pooled, wl.scale, wl.sampling_ratio, threads=0))
    dyg = np.ascontiguousarray(dy[good])
    ref = oracle.backward(dyg, rois[good], data.shape, wl.scale, wl.sampling_ratio)
    assert_bit_exact(dx, ref)
    ref_abs = oracle.backward(np.abs(dyg), rois[good], data.shape, wl.scale, wl.sampling_ratio)
    assert_scatter_close(S.assemble_images(shards, dxa), ref, ref_abs, rtol=1e-5)


# ---- CuPy adapter (through a stub module: CuPy is not in the image) ---------------------------
def test_cupy_operator_through_stub(api, oracle):
    try:
        import cupy
    except ImportError:
        import cupy_stub
        cupy = cupy_stub.install()
    g = load_golden('fpn_14x14_sr2')
    data, rois, dy = cupy.asarray(g['data']), cupy.asarray(g['rois']), cupy.asarray(g['dy'])
    op = api.op.ROIAlign[cupy.ndarray](pooled_size=g['pooled'], spatial_scale=g['scale'],
                                       sampling_ratio=g['sr'], deterministic=True)
    y = op(data, rois)
    assert isinstance(y, cupy.ndarray)
    assert_bit_exact(y.get(), g['y'])
    dx, drois = op.backward(dy)
    assert_bit_exact(dx.get(), g['dx'])
    assert not drois.get().any()
    # plain call form picks the adapter from the argument types
    y2 = api.op.ROIAlign(data, rois, pooled_size=g['pooled'], spatial_scale=g['scale'],
                         sampling_ratio=g['sr'])
    assert_bit_exact(y2.get(), g['y'])
    # a non-contiguous view is copied for the call
    wide = cupy.asarray(np.repeat(g['data'], 2, axis=3))
    y3 = op(wide[:, :, :, ::2], rois)
    assert_bit_exact(y3.get(), g['y'])


# ---- hostile ROIs -------------------------------------------------------------------------------
@pytest.mark.parametrize('sr', [0, 2])
def test_nonfinite_and_gigantic_rois_do_not_hang(direct, oracle, sr):
    """A non-finite coordinate, or an adaptive grid of more than 1024 samples per bin axis, would
    overflow the reference's int grid sizes ((int)inf, ROIAlign.cpp:47-51) and spin a GPU thread
    for minutes: such ROIs produce zero rows and no gradient; their neighbours are untouched."""
    rng = np.random.default_rng(4)
    data = rng.standard_normal((1, 3, 20, 24)).astype(np.float32)
    rois = np.array([[0, 2, 2, 30, 40],
                     [0, 1, 1, np.inf, 9],
                     [0, np.nan, 1, 5, 9],
                     [0, 0, 0, 3.0e7, 3.0e7],
                     [0, 5, 6, 60, 70]], np.float32)
    dy = rng.standard_normal((5, 3, 3, 3)).astype(np.float32)
    bad = [1, 2] + ([3] if sr == 0 else [])
    good = [r for r in range(5) if r not in bad]
    y_ref = oracle.forward(data, rois[good], (3, 3), 0.5, sr, threads=0)
    dyg = np.ascontiguousarray(dy[good])
    dx_ref = oracle.backward(dyg, rois[good], data.shape, 0.5, sr)
    dx_abs = oracle.backward(np.abs(dyg), rois[good], data.shape, 0.5, sr)
    for algo in ('auto', 'gather'):
        y = direct.forward(data, rois, (3, 3), 0.5, sr, algo=algo)
        assert_bit_exact(y[good], y_ref)
        assert not y[bad].any()
        dx = direct.backward(dy, rois, data.shape, 0.5, sr, algo=algo)
        assert_scatter_close(dx, dx_ref, dx_abs, rtol=1e-5)
    assert_bit_exact(direct.backward(dy, rois, data.shape, 0.5, sr, mode='deterministic'), dx_ref)


def test_graph_replay_survives_workspace_growth(api, torch, oracle):
    """A captured graph keeps the workspace it was captured with: a later, larger eager call
    allocates its own and must not free or overwrite the tables the replays read."""
    wl = _sample('C2', C=3, rois_per_image=40)
    big = _sample('C2', C=3, rois_per_image=900)
    dev = torch.device('cuda', 0)
    PH, PW = wl.pooled
    data, rois_h, _ = WL.make_inputs(wl, seed=8)
    x = torch.from_numpy(data).to(dev)
    rois = torch.from_numpy(rois_h).to(dev)
    y = torch.empty(wl.out_shape, device=dev)

    def step():
        api.func.roi_align_forward(int(y.numel()), x, wl.scale, wl.C, wl.H, wl.W, PH, PW,
                                   wl.sampling_ratio, rois, y)

    side = torch.cuda.Stream(dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        step()
    torch.cuda.current_stream(dev).wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        step()
    # an eager call with 22x the ROIs: needs a larger workspace than the captured one
    bdata, brois, _ = WL.make_inputs(big, seed=9)
    by = torch.empty(big.out_shape, device=dev)
    api.func.roi_align_forward(int(by.numel()), torch.from_numpy(bdata).to(dev), big.scale, big.C, big.H,
                               big.W, PH, PW, big.sampling_ratio, torch.from_numpy(brois).to(dev), by)
    y.fill_(float('nan'))
    graph.replay()
    torch.cuda.synchronize(dev)
    assert_bit_exact(y.cpu().numpy(), oracle.forward(data, rois_h, wl.pooled, wl.scale, wl.sampling_ratio))
    assert_bit_exact(by.cpu().numpy(), oracle.forward(bdata, brois, big.pooled, big.scale, big.sampling_ratio))


# ---- aligned=True (half-pixel) variant: torchvision on the CPU is the stated oracle for it ----------
@pytest.mark.parametrize('sr', [2, 0])
def test_aligned_variant_matches_torchvision(api, torch, sr):
    tv = pytest.importorskip('torchvision')
    from torchvision.ops import roi_align as tv_roi_align
    wl = WL.Workload('aligned', 2, 37, 50, 68, 60, (7, 7), 0.25, sr, min_box=2.0, max_box=400.0)
    data, rois, dy = WL.make_inputs(wl, seed=41)
    rois[3, 3] = rois[3, 1] + 0.3                      # narrower than a pixel: NOT widened when aligned
    x_cpu = torch.from_numpy(data).requires_grad_(True)
    y_ref = tv_roi_align(x_cpu, torch.from_numpy(rois), wl.pooled, wl.scale, sr, aligned=True)
    y_ref.backward(torch.from_numpy(dy))
    dev = torch.device('cuda', 0)
    x = torch.from_numpy(data).to(dev).requires_grad_(True)
    y = api.op.ROIAlign(x, torch.from_numpy(rois).to(dev), pooled_size=wl.pooled, spatial_scale=wl.scale,
                        sampling_ratio=sr, aligned=True, deterministic=True)
    y.backward(torch.from_numpy(dy).to(dev))
    from mobulaop_b200 import _native
    assert _native.last_algo(False) == 'sweep' and _native.last_algo(True) == 'tile-canonical'
    # same operations in the same order as torchvision's CPU kernel (no FMA in either): bit for bit
    assert_within_ulp_or_rel(y.detach().cpu().numpy(), y_ref.detach().numpy(), ulp=2, rtol=1e-6)
    assert_bit_exact(y.detach().cpu().numpy(), y_ref.detach().numpy())
    np.testing.assert_allclose(x.grad.cpu().numpy(), x_cpu.grad.numpy(), rtol=1e-5, atol=1e-5)
    # and it differs from the reference convention (aligned=False), as it must
    y0 = api.op.ROIAlign(x, torch.from_numpy(rois).to(dev), pooled_size=wl.pooled, spatial_scale=wl.scale,
                         sampling_ratio=sr)
    assert not np.array_equal(y0.detach().cpu().numpy(), y.detach().cpu().numpy())


def test_fpn_pooler_rows_equal_single_level_calls(api, torch, oracle):
    from mobulaop_b200 import fpn
    rng = np.random.default_rng(12)
    N, C = 2, 5
    shapes = [(64, 88), (32, 44), (16, 22), (8, 11)]
    scales = [0.25, 0.125, 0.0625, 0.03125]
    feats = [rng.standard_normal((N, C) + s).astype(np.float32) for s in shapes]
    xy = rng.uniform(0, 250, (90, 2)).astype(np.float32)
    wh = np.exp(rng.uniform(np.log(8), np.log(600), (90, 2))).astype(np.float32)
    rois = np.concatenate([rng.integers(0, N, (90, 1)).astype(np.float32), xy, xy + wh], 1).astype(np.float32)
    dev = torch.device('cuda', 0)
    y = fpn.roi_align_fpn([torch.from_numpy(f).to(dev) for f in feats], torch.from_numpy(rois).to(dev), 7, scales)
    levels = fpn.assign_levels(rois, 2, 5)
    assert len(set(levels.tolist())) >= 3
    y = y.cpu().numpy()
    for k, (f, s) in enumerate(zip(feats, scales)):
        idx = np.nonzero(levels == 2 + k)[0]
        if len(idx):
            assert_bit_exact(y[idx], oracle.forward(f, np.ascontiguousarray(rois[idx]), (7, 7), s, 2))


@pytest.mark.gpu
@pytest.mark.parametrize('shape', PULL_SHAPES, ids=lambda s: 'x'.join(str(v) for v in s[:5]) + '-p%dx%d-sr%d' % (s[5] + (s[7],)))
def test_pull_grouped_lists_equal_per_pixel_lists(direct, oracle, shape, monkeypatch):
    """Merged-tap pull backward on grouped lists (one entry per bin and group of 8 pixels, the
    gradient line loaded once for the group) against the per-pixel lists: the same fused
    multiply-adds in the same order per pixel, hence bit-identical; and within 1e-5 of the oracle."""
    from mobulaop_b200 import _native
    N, C, H, W, k, pooled, scale, sr = shape
    wl = WL.Workload('pull', N, C, H, W, k, pooled, scale, sr, min_box=2.0, max_box=4.0 * max(H, W) / scale,
                     stress=sr == 0)
    data, rois, dy = WL.make_inputs(wl, seed=41)
    rng = np.random.default_rng(9)
    rois = np.ascontiguousarray(rois[rng.permutation(len(rois))])
    dx_ref = oracle.backward(dy, rois, data.shape, scale, sr)
    dx_abs = oracle.backward(np.abs(dy), rois, data.shape, scale, sr)
    dbase = rng.standard_normal(data.shape).astype(np.float32)
    got = {}
    for grouped in ('0', '1'):
        monkeypatch.setenv('MOBULA_PULL_GROUP', grouped)
        dx = direct.backward(dy, rois, data.shape, scale, sr, algo='pull')
        assert _native.last_algo(True) == 'pull'
        assert_scatter_close(dx, dx_ref, dx_abs, rtol=1e-5)
        acc = direct.backward(dy, rois, data.shape, scale, sr, algo='pull', accumulate_into=dbase.copy())
        got[grouped] = (dx, acc)
    assert_bit_exact(got['1'][0], got['0'][0])
    assert_bit_exact(got['1'][1], got['0'][1])


DENSE_SHAPES = [
    # N, C, H, W, rois/img, scale: pooled 7 x 7, sampling ratio 2, rows of W % 4 == 0 floats
    (2, 3, 50, 68, 40, 0.25),
    (1, 5, 24, 28, 64, 0.5),
    (3, 2, 33, 36, 19, 1.0),         # boxes far larger than the plane: rejected samples, clamped rows / columns
    (1, 2, 200, 272, 48, 0.25),      # the box-head plane
    (2, 4, 14, 16, 32, 1.0),
]


@pytest.mark.gpu
@pytest.mark.parametrize('shape', DENSE_SHAPES, ids=lambda s: 'x'.join(str(v) for v in s[:5]))
def test_dense_plane_forward_equals_padded_plane_forward(direct, oracle, shape, monkeypatch):
    """The 7 x 7 forward on a dense shared-memory plane (pitch == width, brought in by bulk copies; the
    column records carry the clamped / rejected flags the padding columns stood for) against the padded
    plane and the oracle: bit-identical, including boxes that hang over every edge of the plane."""
    from mobulaop_b200 import _native
    N, C, H, W, k, scale = shape
    wl = WL.Workload('dense', N, C, H, W, k, (7, 7), scale, 2, min_box=1.0, max_box=3.0 * max(H, W) / scale, stress=True)
    data, rois, _ = WL.make_inputs(wl, seed=53)
    rng = np.random.default_rng(11)
    # boxes that start left of / above the plane and boxes that end beyond it
    rois[::5, 1:3] -= rng.uniform(0, 0.5 * W / scale, (len(rois[::5]), 2)).astype(np.float32)
    rois[1::5, 3:5] += rng.uniform(0, 0.5 * W / scale, (len(rois[1::5]), 2)).astype(np.float32)
    y_ref = oracle.forward(data, rois, (7, 7), scale, 2)
    got = {}
    for dense in ('0', '1'):
        monkeypatch.setenv('MOBULA_FWD_DENSE', dense)
        got[dense] = direct.forward(data, rois, (7, 7), scale, 2, algo='resident')
        assert _native.last_algo(False) == 'resident'
        assert_bit_exact(got[dense], y_ref)
    assert_bit_exact(got['1'], got['0'])


@pytest.mark.gpu
def test_dense_plane_forward_with_nonfinite_feature_map(direct, oracle, monkeypatch):
    """A rejected sample column must contribute a literal 0 (bilinear.h:15-18) whatever the feature map
    holds: with Inf / NaN in the border columns and rows, the dense-plane kernel (rejected lanes load
    nothing) equals the padded-plane kernel (they read zero columns) and the oracle bit for bit."""
    N, C, H, W, scale = 2, 3, 24, 28, 0.5
    wl = WL.Workload('dense-nan', N, C, H, W, 48, (7, 7), scale, 2, min_box=1.0, max_box=3.0 * W / scale, stress=True)
    data, rois, _ = WL.make_inputs(wl, seed=59)
    rng = np.random.default_rng(13)
    rois[::3, 1:3] -= rng.uniform(0, 0.6 * W / scale, (len(rois[::3]), 2)).astype(np.float32)
    rois[1::3, 3:5] += rng.uniform(0, 0.6 * W / scale, (len(rois[1::3]), 2)).astype(np.float32)
    data[:, :, :, 0] = np.inf
    data[:, :, :, -1] = -np.inf
    data[:, :, 0, 1:-1] = np.nan
    data[:, :, -1, 5] = np.inf
    with np.errstate(invalid='ignore'):
        y_ref = oracle.forward(data, rois, (7, 7), scale, 2)
    assert np.isfinite(y_ref).any() and not np.isfinite(y_ref).all()
    got = {}
    for dense in ('0', '1'):
        monkeypatch.setenv('MOBULA_FWD_DENSE', dense)
        got[dense] = direct.forward(data, rois, (7, 7), scale, 2, algo='resident')
    # NaN payloads are not pinned by IEEE 754: compare where the values are numbers, and the NaN masks
    for a in (got['0'], got['1']):
        assert np.array_equal(np.isnan(a), np.isnan(y_ref))
        ok = ~np.isnan(y_ref)
        assert_bit_exact(a[ok], y_ref[ok])


@pytest.mark.gpu
@pytest.mark.parametrize('mode', ['atomic', 'deterministic'])
def test_pull_backward_does_not_depend_on_the_block_order_of_the_list_grid(direct, mode, monkeypatch):
    """The list / copy grid of the pull backward hands out list blocks first, spread out, or in
    between (MOBULA_TABLES_Q): where a pixel's list lives in the entry array changes with it, what
    the list holds and the order of its entries do not -- dX is bit-identical."""
    wl = WL.Workload('order', 3, 40, 50, 68, 90, (7, 7), 0.25, 2, min_box=2.0, max_box=300.0)
    data, rois, dy = WL.make_inputs(wl, seed=61)
    got = []
    for q in ('0', '1', '3'):
        monkeypatch.setenv('MOBULA_TABLES_Q', q)
        got.append(direct.backward(dy, rois, data.shape, wl.scale, wl.sampling_ratio, algo='pull', mode=mode))
    assert_bit_exact(got[1], got[0])
    assert_bit_exact(got[2], got[0])
