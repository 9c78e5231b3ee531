"""im2col / col2im: the oracle against the reference's golden vectors (CPU), the
CUDA kernels against both through the C-ABI and through mobula.func / mobula.op.Conv2D (`-m gpu`).

Integer/copy work and a fixed summation order: everything is asserted BIT-EXACT.
"""
import ctypes

import numpy as np
import pytest

from conftest import conv_golden_names, load_conv_golden, has_cuda
from mobulaop_b200.testing import assert_bit_exact


# ---- CPU: the checker itself -------------------------------------------------------------------
@pytest.mark.parametrize('name', conv_golden_names())
def test_numpy_restatement_matches_the_reference_vectors(name):
    from oracle import conv as O
    g = load_conv_golden(name)
    assert_bit_exact(O.numpy_im2col(g['im'], g['kernel'], g['pad'], g['stride'], g['dilation']), g['col'])
    assert_bit_exact(O.numpy_col2im(g['colg'], g['im'].shape, g['kernel'], g['pad'], g['stride'], g['dilation']),
                     g['im_grad'])


@pytest.mark.parametrize('name', conv_golden_names())
def test_compiled_reference_reproduces_its_vectors(name):
    from oracle import conv as O
    if not O.RefConv.available():
        pytest.skip('oracle/_ref/libmobula_ref_conv.so is built only where /root/reference exists')
    ref, g = O.RefConv(), load_conv_golden(name)
    assert_bit_exact(ref.im2col(g['im'], g['kernel'], g['pad'], g['stride'], g['dilation']), g['col'])
    assert_bit_exact(ref.col2im(g['colg'], g['im'].shape, g['kernel'], g['pad'], g['stride'], g['dilation']),
                     g['im_grad'])


def test_known_answer_by_hand():
    # 1 channel, 3x3 image, 2x2 kernel, no padding: four patches
    from oracle import conv as O
    im = np.arange(9, dtype=np.float32).reshape(1, 3, 3)
    col = O.numpy_im2col(im, (2, 2), (0, 0), (1, 1), (1, 1))
    assert col.shape == (4, 2, 2)
    assert col[:, 0, 0].tolist() == [0, 1, 3, 4] and col[:, 1, 1].tolist() == [4, 5, 7, 8]
    back = O.numpy_col2im(np.ones_like(col), (1, 3, 3), (2, 2), (0, 0), (1, 1), (1, 1))
    assert back[0].tolist() == [[1, 2, 1], [2, 4, 2], [1, 2, 1]]      # how many patches cover a pixel


def test_symbol_names_follow_the_reference_hash():
    # mobula/func.py:13-50; im2col_kernel and col2im_kernel share their C signature
    import hashlib
    types = 'const int,const float*,' + 'const int,' * 12 + 'float*'
    assert hashlib.md5(types.encode()).hexdigest()[:8] == '341af5d3'


def test_loading_the_operator_publishes_kernels_and_class(native):
    import mobula
    mobula.op.load('Convolution')
    assert callable(mobula.func.im2col) and callable(mobula.func.col2im)
    assert mobula.func.im2col.arg_names[:3] == ['n', 'data_im', 'height']
    assert mobula.func.col2im.arg_names[-1] == 'data_im'
    assert hasattr(mobula.op, 'Conv2D')


# ---- GPU ---------------------------------------------------------------------------------------
def _geom_args(g, OH, OW):
    C, H, W = g['im'].shape
    return [H, W, g['kernel'][0], g['kernel'][1], g['pad'][0], g['pad'][1], g['stride'][0], g['stride'][1],
            g['dilation'][0], g['dilation'][1], OH, OW]


@pytest.mark.gpu
@pytest.mark.skipif(not has_cuda(), reason='needs a CUDA device')
@pytest.mark.parametrize('name', conv_golden_names())
@pytest.mark.parametrize('where', ['device', 'host'])
def test_cuda_kernels_bit_exact_through_the_c_abi(native, name, where):
    import torch
    from mobulaop_b200 import _native
    g = load_conv_golden(name)
    C = g['im'].shape[0]
    OH, OW = g['col'].shape[1:]
    geom = _geom_args(g, OH, OW)
    if where == 'device':
        dev = torch.device('cuda', 0)
        im, colg = torch.from_numpy(g['im']).to(dev), torch.from_numpy(g['colg']).to(dev)
        col, im_grad = torch.full(g['col'].shape, np.nan, device=dev), torch.full(g['im'].shape, np.nan, device=dev)
        vp = lambda t: ctypes.c_void_p(t.data_ptr())
        stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        _native.check(native.mobula_im2col_f32(0, stream, vp(im), C, *geom, vp(col), 0))
        _native.check(native.mobula_col2im_f32(0, stream, vp(colg), C, *geom, vp(im_grad), 0))
        torch.cuda.synchronize()
        col, im_grad = col.cpu().numpy(), im_grad.cpu().numpy()
    else:
        col, im_grad = np.full(g['col'].shape, np.nan, np.float32), np.full(g['im'].shape, np.nan, np.float32)
        vp = lambda a: ctypes.c_void_p(a.ctypes.data)
        flags = _native.make_flags(host=True)
        _native.check(native.mobula_im2col_f32(-1, None, vp(g['im']), C, *geom, vp(col), flags))
        _native.check(native.mobula_col2im_f32(-1, None, vp(g['colg']), C, *geom, vp(im_grad), flags))
    assert_bit_exact(col, g['col'])
    assert_bit_exact(im_grad, g['im_grad'])


@pytest.mark.gpu
@pytest.mark.skipif(not has_cuda(), reason='needs a CUDA device')
@pytest.mark.parametrize('name', ['conv_ref_test', 'conv_dilated'])
def test_dropin_symbols_called_like_the_reference(native, name):
    # exactly the call of mobula/func.py:155-156: explicit ctypes instances, device_id first
    import torch
    g = load_conv_golden(name)
    OH, OW = g['col'].shape[1:]
    ints = [ctypes.c_int(v) for v in _geom_args(g, OH, OW)]
    dev = torch.device('cuda', 0)
    im, colg = torch.from_numpy(g['im']).to(dev), torch.from_numpy(g['colg']).to(dev)
    col, im_grad = torch.empty(g['col'].shape, device=dev), torch.empty(g['im'].shape, device=dev)
    native.im2col_341af5d3(ctypes.c_int(0), ctypes.c_int(col.numel()), ctypes.c_void_p(im.data_ptr()), *ints,
                           ctypes.c_void_p(col.data_ptr()))
    native.col2im_341af5d3(ctypes.c_int(0), ctypes.c_int(im_grad.numel()), ctypes.c_void_p(colg.data_ptr()), *ints,
                           ctypes.c_void_p(im_grad.data_ptr()))
    assert_bit_exact(col.cpu().numpy(), g['col'])        # synchronous like the reference: no sync needed
    assert_bit_exact(im_grad.cpu().numpy(), g['im_grad'])
    # host context (device_id -1): NumPy pointers
    col_h = np.empty(g['col'].shape, np.float32)
    native.im2col_341af5d3(ctypes.c_int(-1), ctypes.c_int(col_h.size), ctypes.c_void_p(g['im'].ctypes.data), *ints,
                           ctypes.c_void_p(col_h.ctypes.data))
    assert_bit_exact(col_h, g['col'])


@pytest.mark.gpu
@pytest.mark.skipif(not has_cuda(), reason='needs a CUDA device')
def test_large_random_shapes_against_the_numpy_restatement(native):
    # property at size: col2im(im2col(x)) = x * (number of patches covering each pixel), exactly
    # for integer-valued x; im2col bit-exact against the restatement
    import torch
    import mobula
    from oracle import conv as O
    mobula.op.load('Convolution')
    rng = np.random.default_rng(5)
    dev = torch.device('cuda', 0)
    for (C, H, W, k, p, s, d) in [(64, 56, 72, (3, 3), (1, 1), (1, 1), (1, 1)), (19, 37, 41, (5, 3), (2, 0), (2, 3), (1, 2))]:
        OH, OW = O.out_size(H, k[0], p[0], s[0], d[0]), O.out_size(W, k[1], p[1], s[1], d[1])
        x = rng.integers(-8, 9, (C, H, W)).astype(np.float32)
        geom = [H, W, k[0], k[1], p[0], p[1], s[0], s[1], d[0], d[1], OH, OW]
        xd = torch.from_numpy(x).to(dev)
        col = torch.empty((C * k[0] * k[1], OH, OW), device=dev)
        mobula.func.im2col(col.numel(), xd, *geom, col)
        assert_bit_exact(col.cpu().numpy(), O.numpy_im2col(x, k, p, s, d))
        back = torch.empty_like(xd)
        mobula.func.col2im(back.numel(), col, *geom, back)
        ones = torch.ones_like(col)
        cover = torch.empty_like(xd)
        mobula.func.col2im(cover.numel(), ones, *geom, cover)
        cover = cover.cpu().numpy()
        # (a pixel no patch covers is +0, not x * 0 = -0 for negative x)
        assert_bit_exact(back.cpu().numpy(), np.where(cover == 0, np.float32(0), x * cover))


@pytest.mark.gpu
@pytest.mark.skipif(not has_cuda(), reason='needs a CUDA device')
@pytest.mark.parametrize('with_bias', [True, False])
def test_conv2d_operator_matches_torch_conv2d(native, with_bias):
    # the reference's own test compares with a framework convolution (test_conv.py:9-52, atol 1e-6
    # on its 2x2x3x4 case); same shape plus a larger one, against torch.nn.functional.conv2d in fp64
    import torch
    import mobula
    mobula.op.load('Convolution')
    dev = torch.device('cuda', 0)
    gen = torch.Generator(device='cpu').manual_seed(11)
    for (N, C, H, W, D, k, s, p, dil) in [(2, 2, 3, 4, 3, (2, 3), (1, 2), (0, 1), (1, 1)),
                                          (3, 8, 20, 24, 16, (3, 3), (2, 1), (1, 2), (2, 1))]:
        x = torch.rand((N, C, H, W), generator=gen).to(dev).requires_grad_(True)
        w = torch.rand((D, C) + k, generator=gen).to(dev).requires_grad_(True)
        b = torch.rand((D,), generator=gen).to(dev).requires_grad_(True) if with_bias else None
        args = (x, w, b) if with_bias else (x, w)
        y = mobula.op.Conv2D(*args, channels=D, kernel_size=k, strides=s, padding=p, dilation=dil)
        x64, w64 = x.detach().double().requires_grad_(True), w.detach().double().requires_grad_(True)
        b64 = b.detach().double().requires_grad_(True) if with_bias else None
        y64 = torch.nn.functional.conv2d(x64, w64, b64, stride=s, padding=p, dilation=dil)
        gy = torch.rand(y.shape, generator=gen).to(dev)
        y.backward(gy)
        y64.backward(gy.double())
        tol = dict(rtol=1e-5, atol=1e-5)
        torch.testing.assert_close(y.double(), y64, **tol)
        torch.testing.assert_close(x.grad.double(), x64.grad, **tol)
        torch.testing.assert_close(w.grad.double(), w64.grad, **tol)
        if with_bias:
            torch.testing.assert_close(b.grad.double(), b64.grad, **tol)


@pytest.mark.gpu
@pytest.mark.skipif(not has_cuda(), reason='needs a CUDA device')
def test_conv2d_operator_on_numpy_arrays(native):
    import mobula
    from oracle import conv as O
    mobula.op.load('Convolution')
    rng = np.random.default_rng(3)
    x = rng.standard_normal((2, 3, 9, 8)).astype(np.float32)
    w = rng.standard_normal((4, 3, 3, 3)).astype(np.float32)
    op = mobula.op.Conv2D[np.ndarray](channels=4, kernel_size=(3, 3), padding=(1, 1))
    y = op(x, w)
    ref = np.stack([(w.reshape(4, -1) @ O.numpy_im2col(x[i], (3, 3), (1, 1), (1, 1), (1, 1)).reshape(27, -1)).reshape(4, 9, 8)
                    for i in range(2)])
    np.testing.assert_allclose(y, ref, rtol=1e-6, atol=1e-6)
    dx, dw = op.backward(np.ones_like(y))
    cols = [w.reshape(4, -1).T @ np.ones((4, 72), np.float32) for _ in range(2)]
    ref_dx = np.stack([O.numpy_col2im(c.reshape(27, 9, 8), (3, 9, 8), (3, 3), (1, 1), (1, 1), (1, 1)) for c in cols])
    np.testing.assert_allclose(dx, ref_dx, rtol=1e-5, atol=1e-5)
    assert dw.shape == w.shape
