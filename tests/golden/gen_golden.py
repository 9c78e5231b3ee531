#!/usr/bin/env python
"""Generate the golden ROIAlign vectors in tests/golden/*.npz.

Runs ONLY in the build container: it imports the UNMODIFIED reference package from
/root/reference (NumPy adapter -> JIT-built CPU library) and records its outputs on
seeded inputs.  The GPU box has no /root/reference, so the vectors are committed.

    python tests/golden/gen_golden.py [--only-new]     # rewrites tests/golden/*.npz

Reference entry points exercised: mobula.op.load('ROIAlign')
(/root/reference/mobula/op/loader.py:623-659) and the NumPy adapter
(/root/reference/mobula/glue/numpy_glue.py:31-103) around
/root/reference/opzoo/ROIAlign/ROIAlign.py:14-42.  The reference is built with
HOST_NUM_THREADS=1 so its backward runs in the canonical ascending-index order
(/root/reference/mobula/cpp/include/context/openmp_ctx.h:50-67) and is reproducible.
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = '/root/reference'


def import_reference(build_dir):
    shim = os.path.join(build_dir, 'shim')
    os.makedirs(shim, exist_ok=True)
    # `portalocker` is not installed in the image; the reference only needs an flock
    with open(os.path.join(shim, 'portalocker.py'), 'w') as f:
        f.write('import fcntl\nLOCK_EX = fcntl.LOCK_EX\n'
                'def lock(f, flags):\n    fcntl.flock(f.fileno(), flags)\n'
                'def unlock(f):\n    fcntl.flock(f.fileno(), fcntl.LOCK_UN)\n')
    sys.path[:0] = [shim, REF]
    os.environ['CXX'] = '/usr/bin/g++'
    import mobula
    assert mobula.__file__.startswith(REF), mobula.__file__
    mobula.config.BUILD_IN_LOCAL_PATH = False
    mobula.config.BUILD_PATH = os.path.join(build_dir, 'build')
    mobula.config.HOST_NUM_THREADS = 1
    mobula.config.CXX = '/usr/bin/g++'
    mobula.op.load('ROIAlign')
    return mobula


def cases():
    rng = np.random.default_rng(20261019)
    out = {}

    # 1. the hand-checkable known answer (test_roialign.py:204-207, examples/RunROIAlign.py)
    data = np.arange(2 * 3 * 4 * 4, dtype=np.float32).reshape(2, 3, 4, 4)
    rois = np.array([[0, 1, 1, 3, 3], [1, 2, 2, 3, 3]], np.float32)
    out['known_answer'] = dict(data=data, rois=rois, pooled=(2, 2), scale=1.0, sr=1,
                               dy=np.ones((2, 3, 2, 2), np.float32))

    # 2. the reference's value-test shape (test_roialign.py:243-261), seeded here
    N, C, H, W, R, dlen = 5, 3, 16, 16, 7, 224
    data = np.arange(N * C * H * W, dtype=np.float32).reshape(N, C, H, W)
    cxy = rng.uniform(0, dlen, (R, 2)).astype(np.float32)
    wh = rng.uniform(0, dlen, (R, 2)).astype(np.float32)
    rois = np.concatenate([rng.integers(0, N, (R, 1)).astype(np.float32),
                           cxy - wh / 2, cxy + wh / 2], axis=1).astype(np.float32)
    out['ref_value_test'] = dict(data=data, rois=rois, pooled=(3, 4), scale=H * 1.0 / dlen, sr=0,
                                 dy=rng.uniform(-1, 1, (R, C, 3, 4)).astype(np.float32))

    # 3. edge cases on odd plane sizes: clamps, degenerate, outside, border
    N, C, H, W = 2, 4, 13, 17
    data = rng.standard_normal((N, C, H, W)).astype(np.float32)
    s = 0.25
    rois = np.array([
        [0, -20.0, -12.0, 30.0, 22.0],        # starts left/above the plane (y,x < 0 clamp)
        [1, 40.0, 30.0, 90.0, 70.0],          # runs past the right/bottom edge (>= H-1 clamp)
        [0, 30.0, 30.0, 10.0, 10.0],          # x2<x1, y2<y1 -> forced 1x1 (ROIAlign.cpp:37-39)
        [1, 12.0, 12.0, 12.0, 12.0],          # zero area
        [0, 400.0, 400.0, 500.0, 500.0],      # fully outside -> zeros (bilinear.h:15-18)
        [1, -300.0, -300.0, -100.0, -100.0],  # fully outside (negative)
        [0, 0.0, 0.0, 68.0, 52.0],            # exactly the plane; last samples hit y == H edge
        [1, 3.3, 7.7, 41.9, 29.1],
        [0, 63.9, 47.9, 68.1, 52.1],          # tiny box on the far corner
        [1, -4.0, 10.0, 72.0, 20.0],          # wider than the plane
    ], np.float32)
    dy = rng.standard_normal((len(rois), C, 7, 7)).astype(np.float32)
    out['edges_sr2'] = dict(data=data, rois=rois, pooled=(7, 7), scale=s, sr=2, dy=dy)
    out['edges_sr0'] = dict(data=data, rois=rois, pooled=(7, 7), scale=s, sr=0, dy=dy)
    out['edges_sr3_p5x3'] = dict(data=data, rois=rois, pooled=(5, 3), scale=s, sr=3,
                                 dy=rng.standard_normal((len(rois), C, 5, 3)).astype(np.float32))

    # 4. the benchmark file's literal workload, fewer channels
    #    (benchmark/benchmark_roi_align.py:15-27)
    N, C, H, W, R = 2, 8, 14, 14, 64
    data = np.arange(N * C * H * W, dtype=np.float32).reshape(N, C, H, W)
    rois = np.array([[i % N, 0, 0, i % H, i % W] for i in range(R)], np.float32)
    out['bench_literal_2x2_sr1'] = dict(data=data, rois=rois, pooled=(2, 2), scale=1.0, sr=1,
                                        dy=rng.standard_normal((R, C, 2, 2)).astype(np.float32))
    out['bench_literal_7x7_sr2'] = dict(data=data, rois=rois, pooled=(7, 7), scale=1.0, sr=2,
                                        dy=rng.standard_normal((R, C, 7, 7)).astype(np.float32))

    # 5. FPN-like random boxes, mask-head pooling, heavy overlap (many adds per pixel)
    N, C, H, W, R = 3, 5, 25, 34, 48
    data = rng.standard_normal((N, C, H, W)).astype(np.float32)
    size = np.exp(rng.uniform(np.log(16), np.log(160), (R, 2)))
    ctr = rng.uniform(0, 1, (R, 2)) * np.array([W / 0.25, H / 0.25])
    rois = np.concatenate([rng.integers(0, N, (R, 1)), ctr - size / 2, ctr + size / 2],
                          axis=1).astype(np.float32)
    out['fpn_14x14_sr2'] = dict(data=data, rois=rois, pooled=(14, 14), scale=0.25, sr=2,
                                dy=rng.standard_normal((R, C, 14, 14)).astype(np.float32))

    # 6. (round 2) the same edge cases on planes whose width is a multiple of 4 -- what the
    #    TMA-fed "sweep" kernels accept -- with their own generator so the vectors above stay put
    rng = np.random.default_rng(20261020)
    N, C, H, W = 2, 35, 13, 20                # 35 channels: one full 32-channel group + a ragged one
    data = rng.standard_normal((N, C, H, W)).astype(np.float32)
    rois = np.array([
        [0, -20.0, -12.0, 30.0, 22.0],
        [1, 40.0, 30.0, 100.0, 70.0],
        [0, 30.0, 30.0, 10.0, 10.0],
        [1, 12.0, 12.0, 12.0, 12.0],
        [0, 400.0, 400.0, 500.0, 500.0],
        [1, -300.0, -300.0, -100.0, -100.0],
        [0, 0.0, 0.0, 80.0, 52.0],            # exactly the plane
        [1, 3.3, 7.7, 41.9, 29.1],
        [0, 75.9, 47.9, 80.1, 52.1],          # tiny box on the far corner
        [1, -4.0, 10.0, 84.0, 20.0],          # wider than the plane
        [0, 10.0, -40.0, 30.0, 90.0],         # taller than the plane
        [1, 76.0, 0.0, 79.9, 52.0],           # last column only (x >= W-1 clamp)
    ], np.float32)
    out['edges_w20_sr2'] = dict(data=data, rois=rois, pooled=(7, 7), scale=0.25, sr=2,
                                dy=rng.standard_normal((len(rois), C, 7, 7)).astype(np.float32))
    out['edges_w20_sr0'] = dict(data=data, rois=rois, pooled=(7, 7), scale=0.25, sr=0,
                                dy=rng.standard_normal((len(rois), C, 7, 7)).astype(np.float32))
    out['edges_w20_sr3_p5x3'] = dict(data=data, rois=rois, pooled=(5, 3), scale=0.25, sr=3,
                                     dy=rng.standard_normal((len(rois), C, 5, 3)).astype(np.float32))
    # a plane wider than one TMA box (two boxes per 8-channel group), ROIs across the seam
    N, C, H, W, R = 2, 3, 11, 280, 40
    data = rng.standard_normal((N, C, H, W)).astype(np.float32)
    size = np.exp(rng.uniform(np.log(8), np.log(900), (R, 2)))
    ctr = rng.uniform(0, 1, (R, 2)) * np.array([W / 0.25, H / 0.25])
    ctr[:12, 0] = rng.uniform(130, 150, 12) / 0.25            # around pixel column 140
    rois = np.concatenate([rng.integers(0, N, (R, 1)), ctr - size / 2, ctr + size / 2],
                          axis=1).astype(np.float32)
    out['wide_w280_sr2'] = dict(data=data, rois=rois, pooled=(7, 7), scale=0.25, sr=2,
                                dy=rng.standard_normal((R, C, 7, 7)).astype(np.float32))
    out['wide_w280_sr0'] = dict(data=data, rois=rois, pooled=(4, 6), scale=0.25, sr=0,
                                dy=rng.standard_normal((R, C, 4, 6)).astype(np.float32))
    return out


def main():
    with tempfile.TemporaryDirectory() as tmp:
        mobula = import_reference(tmp)
        only_new = '--only-new' in sys.argv
        for name, c in cases().items():
            if only_new and os.path.exists(os.path.join(HERE, name + '.npz')):
                continue
            op = mobula.op.ROIAlign[np.ndarray](pooled_size=c['pooled'],
                                                spatial_scale=float(c['scale']),
                                                sampling_ratio=int(c['sr']))
            y = op(c['data'], c['rois'])
            dx, drois = op.backward(c['dy'])
            assert not drois.any()
            np.savez_compressed(
                os.path.join(HERE, name + '.npz'), data=c['data'], rois=c['rois'],
                dy=c['dy'], pooled=np.array(c['pooled'], np.int32),
                scale=np.float64(c['scale']), sr=np.int32(c['sr']),
                y=np.asarray(y, np.float32), dx=np.asarray(dx, np.float32))
            print('%-24s y%s dx%s' % (name, y.shape, dx.shape))


if __name__ == '__main__':
    main()
