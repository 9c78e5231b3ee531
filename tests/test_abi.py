"""The C-ABI boundary without a GPU: the library builds for sm_100a, loads, and exports
every symbol include/mobula_roialign.h declares.  No compute call is made here."""
import ctypes
import os
import re
import subprocess

import pytest

from conftest import ROOT

HEADER = os.path.join(ROOT, 'include', 'mobula_roialign.h')


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    names = re.findall(r'^[A-Za-z_][\w\s\*]*?\b(\w+)\s*\([^;{]*\)\s*;', text, flags=re.M)
    return sorted(set(names))


def test_header_lists_the_reference_symbols():
    names = declared_functions()
    # names the reference's code generator emits for float
    for sym in ('roi_align_forward_ab78de3c', 'roi_align_backward_f318a24e', 'set_device'):
        assert sym in names
    assert len(names) >= 13


def test_reference_symbol_names_follow_the_reference_hash():
    # mobula/func.py:13-50: <func>_<md5(comma-joined C type names)[:8]>
    import hashlib
    common = 'const int,const float*,const float,const int,const int,const int,const int,const int,const int,'
    fwd = common + 'const float*,float*'
    bwd = common + 'float*,const float*'
    assert 'roi_align_forward_' + hashlib.md5(fwd.encode()).hexdigest()[:8] == 'roi_align_forward_ab78de3c'
    assert 'roi_align_backward_' + hashlib.md5(bwd.encode()).hexdigest()[:8] == 'roi_align_backward_f318a24e'


def test_library_exports_every_declared_symbol(native):
    from mobulaop_b200 import _native
    declared = declared_functions()
    assert sorted(_native.EXPORTED_SYMBOLS) == declared
    for sym in declared:
        assert hasattr(native, sym), sym
    out = subprocess.run(['nm', '-D', '--defined-only', _native.library_path()],
                         capture_output=True, text=True).stdout
    exported = set(line.split()[-1] for line in out.splitlines() if ' T ' in line)
    assert set(declared) <= exported
    # nothing but the declared C symbols leaks out of the library
    assert not [s for s in exported - set(declared) if not s.startswith('_')]


def test_abi_version_and_pure_host_entry_points(native):
    assert native.mobula_roialign_abi_version() == 1
    assert isinstance(native.mobula_roialign_launch_count(), int)
    assert native.mobula_roialign_last_algo(0) in (b'none', b'gather', b'plane', b'resident', b'sweep')


def test_invalid_arguments_are_rejected_before_any_cuda_call(native):
    from mobulaop_b200 import _native
    null = ctypes.c_void_p(0)
    rc = native.mobula_roialign_forward_f32(0, null, null, null, null, 4, 1, 8, 0, 16, 7, 7, 0.25, 2, 0)
    assert rc == _native.ERR_INVALID_ARGUMENT
    assert b'bad shape' in native.mobula_last_error()
    rc = native.mobula_roialign_forward_f32(0, null, null, null, null, 4, 1, 8, 16, 16, 7, 7, 0.25, 2, 0)
    assert rc == _native.ERR_INVALID_ARGUMENT and b'null' in native.mobula_last_error()
    with pytest.raises(ValueError):
        _native.check(rc)
    # empty workloads are a no-op and never touch the device
    assert native.mobula_roialign_forward_f32(0, null, null, null, null, 0, 1, 8, 16, 16, 7, 7, 0.25, 2, 0) == 0
    assert native.mobula_roialign_backward_f32(0, null, null, null, null, 0, 0, 8, 16, 16, 7, 7, 0.25, 2, 0, 1) == 0


def test_sass_is_sm100a_only():
    from mobulaop_b200 import _native, build
    build.build()
    out = subprocess.run(['cuobjdump', '-lelf', _native.library_path()], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip('cuobjdump unavailable')
    archs = set(re.findall(r'sm_\d+a?', out.stdout))
    assert archs == {'sm_100a'}, archs


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'mobulaop_b200')
    for base, _dirs, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                text = open(os.path.join(base, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', text, flags=re.M), f
                assert 'liboracle' not in text and 'oracle/_' not in text, f
