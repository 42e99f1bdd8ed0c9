"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol include/ld_b200.h
declares (no compute calls without a GPU)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, 'include', 'ld_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(ldb200_\w+)\s*\(', src)))


def test_header_matches_binding():
    from ld_b200 import lib as L
    assert declared_symbols() == L.exported_symbols()


def test_library_builds_loads_and_exports():
    from ld_b200 import build, lib as L
    path = build.build()
    assert os.path.exists(path)
    so = ctypes.CDLL(path)
    for name in declared_symbols():
        assert hasattr(so, name), name
    lib = L.load()
    assert lib.ldb200_version() >= 100
    assert ctypes.sizeof(L.Pyramid) == 4 * (2 + 3 * L.MAX_LEVELS)


def test_no_fallback_when_library_missing(monkeypatch):
    from ld_b200 import lib as L
    monkeypatch.setattr(L, '_lib', None)
    monkeypatch.setattr(L, '_LIB_PATH', '/nonexistent/libld_b200.so')
    import pytest
    with pytest.raises(L.LdB200Error):
        L.load()
