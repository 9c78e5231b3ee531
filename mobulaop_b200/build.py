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

SOURCES = ['roialign_api.cu', 'roialign_gather.cu', 'roialign_tables.cu', 'roialign_plane.cu', 'roialign_resident.cu',
           'roialign_sweep.cu', 'roialign_pull.cu', 'im2col.cu']
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


def _object_path(src):
    return os.path.join(LIB_DIR, 'obj', os.path.basename(src)[:-3] + '.o')


def _compile(nvcc, src, obj, verbose):
    cmd = [nvcc] + [f for f in NVCC_FLAGS if f != '-shared'] + os.environ.get('MOBULA_NVCC_EXTRA', '').split()
    if os.path.exists('/usr/bin/g++'):
        cmd += ['-ccbin', '/usr/bin/g++']     # the image's $CC wrapper lacks libgomp specs
    if verbose:
        cmd += ['-Xptxas', '-v']
    tmp = obj + '.tmp.%d' % os.getpid()
    cmd += ['-c', src, '-o', tmp]
    out = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or out.returncode != 0:
        sys.stderr.write(' '.join(cmd) + '\n' + out.stdout + out.stderr)
    if out.returncode != 0:
        if os.path.exists(tmp):
            os.remove(tmp)
        raise RuntimeError('nvcc failed compiling %s' % src)
    os.replace(tmp, obj)


def build(force=False, verbose=False):
    """Compile the library in-tree (one object per source, in parallel); returns its path."""
    if not force and not is_stale():
        return LIB_PATH
    from concurrent.futures import ThreadPoolExecutor
    os.makedirs(os.path.join(LIB_DIR, 'obj'), exist_ok=True)
    nvcc = find_nvcc()
    headers = [os.path.normpath(os.path.join(CSRC_DIR, h)) for h in HEADERS]
    newest_header = max(os.path.getmtime(h) for h in headers if os.path.exists(h))
    todo = []
    for src in sources():
        obj = _object_path(src)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), newest_header):
            todo.append((src, obj))
    with ThreadPoolExecutor(max_workers=max(1, min(len(todo), os.cpu_count() or 1))) as pool:
        for fut in [pool.submit(_compile, nvcc, src, obj, verbose) for src, obj in todo]:
            fut.result()
    cmd = [nvcc, '-shared', '-gencode', 'arch=compute_100a,code=sm_100a']
    if os.path.exists('/usr/bin/g++'):
        cmd += ['-ccbin', '/usr/bin/g++']
    tmp = LIB_PATH + '.tmp.%d' % os.getpid()
    cmd += ['-o', tmp] + [_object_path(src) for src in sources()]
    out = subprocess.run(cmd, capture_output=True, text=True)
    if out.returncode != 0:
        sys.stderr.write(' '.join(cmd) + '\n' + out.stdout + out.stderr)
        if os.path.exists(tmp):
            os.remove(tmp)
        raise RuntimeError('nvcc failed linking %s' % LIB_NAME)
    os.replace(tmp, LIB_PATH)
    return LIB_PATH


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
