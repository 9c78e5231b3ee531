import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


def golden_names():
    """ROIAlign vectors (tests/golden/gen_golden.py)."""
    return sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.endswith('.npz') and not f.startswith('conv_'))


def conv_golden_names():
    """im2col / col2im vectors (tests/golden/gen_conv_golden.py)."""
    return sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.endswith('.npz') and f.startswith('conv_'))


def load_conv_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + '.npz'))
    pair = lambda k: tuple(int(v) for v in z[k])
    return dict(im=z['im'], colg=z['colg'], kernel=pair('kernel'), pad=pair('pad'), stride=pair('stride'),
                dilation=pair('dilation'), col=z['col'], im_grad=z['im_grad'])


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + '.npz'))
    return dict(data=z['data'], rois=z['rois'], dy=z['dy'], pooled=tuple(int(v) for v in z['pooled']),
                scale=float(z['scale']), sr=int(z['sr']), y=z['y'], dx=z['dx'])


@pytest.fixture(scope='session')
def oracle():
    from oracle.roialign import COracle
    return COracle()


@pytest.fixture(scope='session')
def native():
    """The CUDA library through ctypes; building it needs nvcc, loading it needs no GPU."""
    from mobulaop_b200 import _native, build
    build.build()
    return _native.load()


_HAS_CUDA = None


def has_cuda():
    """Cached; waits out the driver's start-up on a freshly started GPU box (see cuda_ready)."""
    global _HAS_CUDA
    if _HAS_CUDA is None:
        from mobulaop_b200.testing import cuda_ready
        _HAS_CUDA = cuda_ready()
    return _HAS_CUDA
