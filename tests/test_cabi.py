"""CPU: the C-ABI shared library builds, loads and exports every symbol include/pda_b200.h declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "pda_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pda_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol():
    from probabilistic_domain_adaptation_b200 import build, _lib
    path = build.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    syms = _declared_symbols()
    assert len(syms) >= 10
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in pda_b200.h but not exported"
        assert s in _lib.SIGNATURES, f"{s} has no ctypes signature in _lib.py"
    assert set(_lib.SIGNATURES) == set(syms)
    loaded = _lib.load()
    assert loaded.pda_abi_version() == 1
    assert loaded.pda_error_string(-1).decode().startswith("unsupported")


def test_no_cpu_fallback():
    import pytest
    import torch
    from probabilistic_domain_adaptation_b200 import ops, _lib
    with pytest.raises(_lib.PdaError):
        ops.avgpool2(torch.zeros(1, 2, 2, 8, dtype=torch.bfloat16))
