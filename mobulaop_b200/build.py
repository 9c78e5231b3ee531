"""Ahead-of-time build of libmobula_roialign_sm100.so (the only native artefact).

The reference JIT-compiles every operator on first call and picks compiler flags per
context (/root/reference/mobula/building/build.py:82-139, with no -arch flag at all);
here there is one library, one architecture (sm_100a) and the build happens before
use: `python -m mobulaop_b200.build` or `__graft_entry__.build()`.
"""
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC_DIR = os.path.join(PKG_DIR, 'csrc')
LIB_DIR = os.path.join(PKG_DIR, 'lib')
LIB_NAME = 'libmobula_roialign_sm100.so'
LIB_PATH = os.path.join(LIB_DIR, LIB_NAME)

SOURCES = ['roialign_api.cu', 'roialign_gather.cu', 'roialign_tables.cu', 'roialign_plane.cu', 'roialign_resident.cu']
HEADERS = ['roialign_kernels.h', 'roialign_common.cuh', 'ptx_sm100.cuh',
           os.path.join('..', '..', 'include', 'mobula_roialign.h')]

NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a',   # Blackwell B200 only
    '-lineinfo', '-O3', '-std=c++17',
    '-fmad=false',                                  # see roialign_common.cuh
    '-Xcompiler', '-fPIC,-fvisibility=hidden',
    '-shared',
]


def find_nvcc():
    for cand in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError('nvcc not found (set NVCC)')


def sources():
    return [os.path.join(CSRC_DIR, s) for s in SOURCES if os.path.exists(os.path.join(CSRC_DIR, s))]


def is_stale():
    if not os.path.exists(LIB_PATH):
        return True
    built = os.path.getmtime(LIB_PATH)
    deps = sources() + [os.path.normpath(os.path.join(CSRC_DIR, h)) for h in HEADERS]
    return any(os.path.exists(d) and os.path.getmtime(d) > built for d in deps)


def build(force=False, verbose=False):
    """Compile the library in-tree; returns its path."""
    if not force and not is_stale():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    cmd = [find_nvcc()] + NVCC_FLAGS + os.environ.get('MOBULA_NVCC_EXTRA', '').split()
    if os.path.exists('/usr/bin/g++'):
        cmd += ['-ccbin', '/usr/bin/g++']     # the image's $CC wrapper lacks libgomp specs
    if verbose:
        cmd += ['-Xptxas', '-v']
    tmp = LIB_PATH + '.tmp.%d' % os.getpid()
    cmd += ['-o', tmp] + sources()
    out = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or out.returncode != 0:
        sys.stderr.write(' '.join(cmd) + '\n' + out.stdout + out.stderr)
    if out.returncode != 0:
        if os.path.exists(tmp):
            os.remove(tmp)
        raise RuntimeError('nvcc failed building %s' % LIB_NAME)
    os.replace(tmp, LIB_PATH)
    return LIB_PATH


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
