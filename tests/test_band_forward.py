"""The banded channel-pair forward (resident_fwd2b_kernel) and the single-plane forward it replaces on
planes whose channel pair exceeds shared memory: bit-exact against the oracle, including boxes several
times the plane (bin rows no band can hold -> the kernel's global-memory path), boxes outside the
plane, degenerate boxes, a row that points at no image, accumulate and the fp16 store -- `-m gpu`.
One child process per kernel choice: MOBULA_FWD_BAND is read once per process."""
import os
import subprocess
import sys

import pytest

from conftest import has_cuda

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not has_cuda(), reason='needs a CUDA device')]

WORKER = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'band_forward_worker.py')


@pytest.mark.parametrize('band', ['1', '0'])
def test_forward_bit_exact_with_and_without_bands(native, band):
    env = dict(os.environ, MOBULA_FWD_BAND=band)
    out = subprocess.run([sys.executable, WORKER], env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and 'OK band=' + band in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
