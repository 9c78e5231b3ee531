"""fp16 / bf16 and channel-last pooled tensors -- `-m gpu`.

The arithmetic is the fp32 path (bit-identical to the reference); the output is that result rounded
once to nearest-even at the store, so the expected value is round(oracle) EXACTLY; a typed dy is
widened exactly, so the expected dX is the oracle's backward of the widened dy (bit-exact in
deterministic mode, within 1e-5 of the summand magnitude otherwise).
"""
import ctypes

import numpy as np
import pytest

from conftest import load_golden, has_cuda
from mobulaop_b200.testing import assert_bit_exact, assert_scatter_close

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not has_cuda(), reason='needs a CUDA device')]

CASES = ['fpn_14x14_sr2', 'wide_w280_sr2', 'edges_sr2', 'edges_sr0', 'edges_sr3_p5x3', 'bench_literal_7x7_sr2']
DTYPES = {'float16': 1, 'bfloat16': 2, 'float32': 0}


def _round_to(y, dtype_name):
    import torch
    return torch.from_numpy(np.ascontiguousarray(y)).to(getattr(torch, dtype_name))


def _bits(t):
    import torch
    t = t.detach().cpu().contiguous()
    if t.dtype == torch.float32:
        return t.numpy().view(np.uint32)
    return t.view(torch.int16).numpy()


def _layout(y, layout):
    return np.ascontiguousarray(y.transpose(0, 2, 3, 1)) if layout == 'RHWC' else y


@pytest.mark.parametrize('name', CASES)
@pytest.mark.parametrize('dtype_name', ['float16', 'bfloat16', 'float32'])
@pytest.mark.parametrize('layout', ['RCHW', 'RHWC'])
@pytest.mark.parametrize('algo', ['auto', 'sweep'])
def test_forward_is_the_oracle_rounded_once(native, name, dtype_name, layout, algo):
    import torch
    from mobulaop_b200 import _native
    g = load_golden(name)
    if g['pooled'][1] > 16:
        pytest.skip('typed outputs outside the resident kernels need the sweep kernels: pooled width <= 16')
    dev = torch.device('cuda', 0)
    data, rois = torch.from_numpy(g['data']).to(dev), torch.from_numpy(g['rois']).to(dev)
    expect = _round_to(_layout(g['y'], layout), dtype_name)
    out = torch.full(expect.shape, float('nan'), dtype=expect.dtype, device=dev)
    N, C, H, W = g['data'].shape
    vp = lambda t: ctypes.c_void_p(t.data_ptr())
    _native.check(native.mobula_roialign_forward_ex(
        0, ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream), vp(data), vp(rois), vp(out),
        DTYPES[dtype_name], _native.LAYOUTS[layout], g['rois'].shape[0], N, C, H, W, g['pooled'][0], g['pooled'][1],
        g['scale'], g['sr'], _native.make_flags(algo=algo)))
    torch.cuda.synchronize()
    assert np.array_equal(_bits(out), _bits(expect))
    # the type-only case of a sampling_ratio 2 shape stays on the resident kernels
    if algo == 'auto' and layout == 'RCHW' and g['sr'] == 2 and max(g['pooled']) <= 16:
        assert _native.last_algo(False) == 'resident'


@pytest.mark.parametrize('name', CASES)
@pytest.mark.parametrize('dtype_name', ['float16', 'bfloat16', 'float32'])
@pytest.mark.parametrize('layout', ['RCHW', 'RHWC'])
@pytest.mark.parametrize('deterministic', [False, True])
def test_backward_widens_dy_exactly(native, oracle, name, dtype_name, layout, deterministic):
    import torch
    from mobulaop_b200 import _native
    g = load_golden(name)
    dev = torch.device('cuda', 0)
    dy_t = _round_to(g['dy'], dtype_name)                         # what a half-precision consumer hands back
    dy_wide = dy_t.float().numpy()
    ref = oracle.backward(dy_wide, g['rois'], g['data'].shape, g['scale'], g['sr'])
    ref_abs = oracle.backward(np.abs(dy_wide), g['rois'], g['data'].shape, g['scale'], g['sr'])
    dy_dev = _round_to(_layout(dy_wide, layout), dtype_name).to(dev)
    rois = torch.from_numpy(g['rois']).to(dev)
    dx = torch.full(g['data'].shape, float('nan'), device=dev)
    N, C, H, W = g['data'].shape
    vp = lambda t: ctypes.c_void_p(t.data_ptr())
    _native.check(native.mobula_roialign_backward_ex(
        0, ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream), vp(dy_dev), DTYPES[dtype_name],
        _native.LAYOUTS[layout], vp(rois), vp(dx), g['rois'].shape[0], N, C, H, W, g['pooled'][0], g['pooled'][1],
        g['scale'], g['sr'], _native.BWD_DETERMINISTIC if deterministic else _native.BWD_ATOMIC, 0))
    torch.cuda.synchronize()
    if deterministic:
        assert_bit_exact(dx.cpu().numpy(), ref)
    else:
        assert_scatter_close(dx.cpu().numpy(), ref, ref_abs, rtol=1e-5)
    if dtype_name != 'float32' or layout != 'RCHW':
        assert _native.last_algo(True).startswith('pull')


@pytest.mark.parametrize('out_dtype,layout', [('bfloat16', 'RHWC'), ('float16', 'RCHW'), (None, 'RHWC')])
def test_operator_through_torch_autograd(native, oracle, out_dtype, layout):
    import torch
    import mobula
    mobula.op.load('ROIAlign')
    g = load_golden('fpn_14x14_sr2')
    dev = torch.device('cuda', 0)
    x = torch.from_numpy(g['data']).to(dev).requires_grad_(True)
    rois = torch.from_numpy(g['rois']).to(dev)
    y = mobula.op.ROIAlign(x, rois, pooled_size=g['pooled'], spatial_scale=g['scale'], sampling_ratio=g['sr'],
                           out_dtype=out_dtype, layout=layout, deterministic=True)
    name = out_dtype or 'float32'
    expect = _round_to(_layout(g['y'], layout), name)
    assert y.dtype == expect.dtype and tuple(y.shape) == tuple(expect.shape)
    assert np.array_equal(_bits(y), _bits(expect))
    gy = _round_to(_layout(g['dy'], layout), name).to(dev)
    y.backward(gy)
    wide = gy.float().cpu().numpy()
    if layout == 'RHWC':
        wide = np.ascontiguousarray(wide.transpose(0, 3, 1, 2))
    assert_bit_exact(x.grad.cpu().numpy(), oracle.backward(wide, g['rois'], g['data'].shape, g['scale'], g['sr']))


def test_host_buffers_with_a_half_precision_output(native):
    # NumPy float16 output through the staged path: half the device -> host bytes
    import mobula
    mobula.op.load('ROIAlign')
    g = load_golden('wide_w280_sr2')
    op = mobula.op.ROIAlign[np.ndarray](pooled_size=g['pooled'], spatial_scale=g['scale'], sampling_ratio=g['sr'],
                                        out_dtype='float16', deterministic=True)
    y = op(g['data'], g['rois'])
    assert y.dtype == np.float16
    assert np.array_equal(y.view(np.uint16), g['y'].astype(np.float16).view(np.uint16))
    dx = op.backward(g['dy'].astype(np.float16))[0]
    from oracle.roialign import COracle
    ref = COracle().backward(g['dy'].astype(np.float16).astype(np.float32), g['rois'], g['data'].shape, g['scale'], g['sr'])
    assert_bit_exact(dx, ref)


def test_accumulate_needs_the_reference_format(native):
    import torch
    from mobulaop_b200 import _native
    g = load_golden('edges_sr2')
    dev = torch.device('cuda', 0)
    data, rois = torch.from_numpy(g['data']).to(dev), torch.from_numpy(g['rois']).to(dev)
    out = torch.zeros(g['y'].shape, dtype=torch.float16, device=dev)
    N, C, H, W = g['data'].shape
    vp = lambda t: ctypes.c_void_p(t.data_ptr())
    rc = native.mobula_roialign_forward_ex(0, None, vp(data), vp(rois), vp(out), 1, 0, g['rois'].shape[0], N, C, H, W,
                                           g['pooled'][0], g['pooled'][1], g['scale'], g['sr'],
                                           _native.make_flags(accumulate=True))
    assert rc == _native.ERR_UNSUPPORTED
    assert native.mobula_roialign_forward_ex(0, None, vp(data), vp(rois), vp(out), 7, 0, g['rois'].shape[0], N, C, H, W,
                                             g['pooled'][0], g['pooled'][1], g['scale'], g['sr'], 0) == _native.ERR_INVALID_ARGUMENT
