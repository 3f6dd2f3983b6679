"""CPU: the C-ABI library loads, exports every symbol include/bbgpu.h declares, and refuses to run without a GPU."""
import ctypes
import os
import re

import pytest

from conftest import ROOT
from bayesbridge_b200 import _lib


def _declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'bbgpu.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(bb_[a-z0-9_]+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared_symbols()
    assert len(names) >= 30
    for name in names:
        assert hasattr(lib, name), name


def test_python_binding_covers_the_header():
    assert set(_declared_symbols()) == set(_lib.SIGNATURES)


def test_version_and_error_string():
    lib = _lib.load()
    assert lib.bb_version() >= 100
    assert lib.bb_last_error() is not None


def test_no_silent_cpu_fallback():
    """Without a visible GPU the context cannot be created and nothing computes."""
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    with pytest.raises(RuntimeError):
        _lib.Context(0)


def test_product_code_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'bayesbridge_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                src = open(os.path.join(dirpath, f)).read()
                assert 'import oracle' not in src and 'from oracle' not in src and 'oracle/' not in src, f
