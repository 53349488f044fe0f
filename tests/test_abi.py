"""The C-ABI library loads and exports every symbol include/kgv.h declares (no compute calls: no GPU here)."""
import ctypes
import os
import re

from conftest import ROOT


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "kgv.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(kgv_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    path = os.path.join(ROOT, "rusty_kaspa_b200", "libkgv.so")
    assert os.path.exists(path), "libkgv.so missing: run __graft_entry__.build()"
    lib = ctypes.CDLL(path)
    names = declared_symbols()
    assert len(names) >= 10
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/kgv.h but not exported"


def test_python_binding_table_matches_header():
    from rusty_kaspa_b200 import _lib
    bound = sorted(n for n, _, _ in _lib.SYMBOLS)
    assert bound == declared_symbols()


def test_no_cpu_fallback_without_device():
    """On a box without a CUDA device context creation must fail loudly, never fall back."""
    import torch
    import rusty_kaspa_b200 as rk
    import pytest
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    with pytest.raises(rk.KgvError):
        rk.GpuContext(0)
