"""The oracle (oracle/) pinned against the reference: golden vectors produced by the
unmodified reference (tests/golden/gen_golden.py), its hand-checkable known answer
 and, where present, the reference's own C++ compiled as oracle/_ref."""
import numpy as np
import pytest

from conftest import golden_names, load_golden
from mobulaop_b200.testing import assert_bit_exact, assert_scatter_close
from oracle import roialign as O


@pytest.mark.parametrize('name', golden_names())
def test_c_oracle_matches_reference_golden_bitwise(oracle, name):
    g = load_golden(name)
    y = oracle.forward(g['data'], g['rois'], g['pooled'], g['scale'], g['sr'])
    assert_bit_exact(y, g['y'])
    dx = oracle.backward(g['dy'], g['rois'], g['data'].shape, g['scale'], g['sr'])
    assert_bit_exact(dx, g['dx'])


def test_known_answer(oracle):
    # inputs: /root/reference/opzoo/ROIAlign/test_roialign.py:204-207
    data = np.arange(2 * 3 * 4 * 4, dtype=np.float32).reshape(2, 3, 4, 4)
    rois = np.array([[0, 1, 1, 3, 3], [1, 2, 2, 3, 3]], np.float32)
    y = oracle.forward(data, rois, (2, 2), 1.0, 1)
    np.testing.assert_array_equal(y[0, 0], [[7.5, 8.5], [11.5, 12.5]])
    np.testing.assert_array_equal(y[1, 0], [[59.25, 59.75], [61.25, 61.75]])
    dx = oracle.backward(np.ones_like(y), rois, data.shape, 1.0, 1)
    np.testing.assert_array_equal(
        dx[0, 0], [[0, 0, 0, 0], [0, .25, .5, .25], [0, .5, 1, .5], [0, .25, .5, .25]])


@pytest.mark.parametrize('name', golden_names())
def test_multithread_oracle_forward_is_order_independent(oracle, name):
    g = load_golden(name)
    y = oracle.forward(g['data'], g['rois'], g['pooled'], g['scale'], g['sr'], threads=4)
    assert_bit_exact(y, g['y'])
    dx = oracle.backward(g['dy'], g['rois'], g['data'].shape, g['scale'], g['sr'], threads=4)
    abs_sum = oracle.backward(np.abs(g['dy']), g['rois'], g['data'].shape, g['scale'], g['sr'])
    assert_scatter_close(dx, g['dx'], abs_sum, rtol=1e-5)


@pytest.mark.parametrize('name', ['known_answer', 'ref_value_test', 'edges_sr2', 'edges_sr0'])
def test_numpy_restatement_matches(oracle, name):
    g = load_golden(name)
    y = O.numpy_forward(g['data'], g['rois'], g['pooled'], g['scale'], g['sr'])
    assert_bit_exact(y, g['y'])


@pytest.mark.parametrize('name', ['known_answer', 'edges_sr3_p5x3'])
def test_numpy_backward_restatement_matches(name):
    g = load_golden(name)
    dx = O.numpy_backward(g['dy'], g['rois'], g['data'].shape, g['scale'], g['sr'])
    assert_bit_exact(dx, g['dx'])


@pytest.mark.skipif(not O.RefLib.available('st'), reason='oracle/_ref not built')
@pytest.mark.parametrize('name', golden_names())
def test_compiled_reference_matches_golden(oracle, name):
    g = load_golden(name)
    ref = O.RefLib('st')
    assert ref.host_threads() == 1
    assert_bit_exact(ref.forward(g['data'], g['rois'], g['pooled'], g['scale'], g['sr']), g['y'])
    assert_bit_exact(ref.backward(g['dy'], g['rois'], g['data'].shape, g['scale'], g['sr']), g['dx'])


@pytest.mark.skipif(not O.RefLib.available('st'), reason='oracle/_ref not built')
def test_oracle_matches_compiled_reference_on_random_workload(oracle):
    from mobulaop_b200 import workloads as W
    wl = W.CONFIGS['C5'].scaled(N=2, C=3, rois_per_image=60)
    wl.H, wl.W = 40, 56
    data, rois, dy = W.make_inputs(wl, seed=7)
    ref = O.RefLib('st')
    for sr in (0, 2):
        assert_bit_exact(oracle.forward(data, rois, (7, 7), wl.scale, sr),
                         ref.forward(data, rois, (7, 7), wl.scale, sr))
        assert_bit_exact(oracle.backward(dy, rois, data.shape, wl.scale, sr),
                         ref.backward(dy, rois, data.shape, wl.scale, sr))


def test_bookkeeping_agrees_with_numpy_geometry(oracle):
    g = load_golden('edges_sr0')
    gi, gf, ti, tf = oracle.bookkeeping(g['rois'], g['data'].shape[2:], g['pooled'], g['scale'],
                                        g['sr'], max_grid=8)
    geo = O.numpy_geometry(g['rois'], g['pooled'], g['scale'], g['sr'])
    np.testing.assert_array_equal(gi[:, 0], geo['batch'])
    np.testing.assert_array_equal(gi[:, 1], geo['grid_h'])
    np.testing.assert_array_equal(gi[:, 2], geo['grid_w'])
    assert_bit_exact(gf[:, 0], geo['start_w'])
    assert_bit_exact(gf[:, 3], geo['bin_h'])
    H = g['data'].shape[2]
    for r in range(len(g['rois'])):
        gh = min(int(geo['grid_h'][r]), 8)
        pos, lo, hi, frac, valid = O.numpy_axis(geo['start_h'][r], geo['bin_h'][r], g['pooled'][0],
                                                int(geo['grid_h'][r]), H)
        np.testing.assert_array_equal(ti[r, 0, :g['pooled'][0], :gh, 0],
                                      np.where(valid, lo, -1)[:, :gh])
        np.testing.assert_array_equal(ti[r, 0, :g['pooled'][0], :gh, 1],
                                      np.where(valid, hi, -1)[:, :gh])


def test_empty_inputs(oracle):
    data = np.zeros((1, 2, 4, 4), np.float32)
    rois = np.zeros((0, 5), np.float32)
    assert oracle.forward(data, rois, (2, 2), 1.0, 2).shape == (0, 2, 2, 2)
    dx = oracle.backward(np.zeros((0, 2, 2, 2), np.float32), rois, data.shape, 1.0, 2)
    assert not dx.any()


def test_unique_elements_counter():
    data = np.zeros((1, 2, 8, 8), np.float32)
    rois = np.array([[0, 0, 0, 4, 4]], np.float32)
    u = O.unique_elements_touched(rois, data.shape, (2, 2), 1.0, 1)
    # samples at 1 and 3 along both axes, integer coordinates -> lo=pos, hi=pos+1
    assert u == 2 * 16
