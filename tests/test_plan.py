"""Prepared ROI tables (mobula_roialign_prepare_f32): one grouping of the ROI
rows, shared by forward and backward -- `-m gpu`, bit-exact against the reference's vectors."""
import ctypes

import numpy as np
import pytest

from conftest import load_golden, has_cuda
from mobulaop_b200.testing import assert_bit_exact

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not has_cuda(), reason='needs a CUDA device')]


@pytest.mark.parametrize('name', ['fpn_14x14_sr2', 'wide_w280_sr2', 'edges_sr0', 'ref_value_test'])
@pytest.mark.parametrize('shuffle', [False, True])
def test_forward_and_backward_share_one_plan(native, oracle, name, shuffle):
    import torch
    from mobulaop_b200 import _native
    g = load_golden(name)
    rois_np, y_ref, dy_np, dx_ref = g['rois'], g['y'], g['dy'], g['dx']
    if shuffle:          # rows of different images interleaved: the plan's partition is not the identity
        perm = np.random.default_rng(1).permutation(rois_np.shape[0])
        rois_np, y_ref, dy_np = np.ascontiguousarray(rois_np[perm]), y_ref[perm], np.ascontiguousarray(dy_np[perm])
        # (the canonical accumulation order follows the rows of the table: ask the oracle again)
        dx_ref = oracle.backward(dy_np, rois_np, g['data'].shape, g['scale'], g['sr'])
    dev = torch.device('cuda', 0)
    data = torch.from_numpy(g['data']).to(dev)
    rois = torch.from_numpy(np.ascontiguousarray(rois_np)).to(dev)
    dy = torch.from_numpy(np.ascontiguousarray(dy_np)).to(dev)
    N, C, H, W = g['data'].shape
    vp = lambda t: ctypes.c_void_p(t.data_ptr())
    stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    plan = _native.RoiPlan(0, stream, vp(rois), rois_np.shape[0], N, H, W, g['pooled'], g['scale'], g['sr'])
    rois.fill_(float('nan'))                     # the plan keeps its own copy of the table
    out = torch.full(y_ref.shape, float('nan'), device=dev)
    dx = torch.full(g['data'].shape, float('nan'), device=dev)
    launches = _native.launch_count()
    plan.forward(stream, vp(data), vp(out), C)
    plan.backward(stream, vp(dy), vp(dx), C, mode=_native.BWD_DETERMINISTIC)
    plan.forward(stream, vp(data), vp(out), C)   # a plan serves any number of calls
    torch.cuda.synchronize()
    assert _native.launch_count() > launches
    assert_bit_exact(out.cpu().numpy(), y_ref)
    assert_bit_exact(dx.cpu().numpy(), dx_ref)
    plan.close()
    plan.close()                                  # idempotent


def test_plan_errors(native):
    import torch
    from mobulaop_b200 import _native
    bogus = (ctypes.c_char * 64)()
    rc = native.mobula_roialign_forward_prepared_f32(ctypes.cast(bogus, ctypes.c_void_p), None, None, None, 4, 0)
    assert rc == _native.ERR_INVALID_ARGUMENT
    handle = ctypes.c_void_p()
    rois = torch.zeros((2, 5), device='cuda')
    rc = native.mobula_roialign_prepare_f32(0, None, ctypes.c_void_p(rois.data_ptr()), 2, -1, 8, 8, 2, 2, 1.0, 2, 0,
                                            ctypes.byref(handle))
    assert rc == _native.ERR_INVALID_ARGUMENT and not handle      # the grouping needs num_images
    rc = native.mobula_roialign_prepare_f32(0, None, ctypes.c_void_p(rois.data_ptr()), 2, 1, 8, 8, 2, 2, 1.0, 2,
                                            _native.FLAG_HOST_BUFFERS, ctypes.byref(handle))
    assert rc == _native.ERR_UNSUPPORTED
