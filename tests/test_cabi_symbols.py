"""The C-ABI library must build, load, and export exactly what include/p2v.h declares (no GPU needed)."""
import ctypes
import os
import re

from conftest import ROOT


def _declared():
    hdr = open(os.path.join(ROOT, 'include', 'p2v.h')).read()
    hdr = re.sub(r'/\*.*?\*/', '', hdr, flags=re.S)
    return sorted(set(re.findall(r'\b(p2v_[a-z0-9_]+)\s*\(', hdr)))


def test_library_builds_and_exports_header_symbols():
    from diff_vit_b200 import _cabi
    from diff_vit_b200.build import build
    path = build()
    assert os.path.exists(path)
    handle = ctypes.CDLL(path)
    declared = _declared()
    assert len(declared) >= 18
    for name in declared:
        assert hasattr(handle, name), name
    assert sorted(_cabi.SYMBOLS) == declared
    lib = _cabi.lib()
    assert lib.p2v_version() == 100
    assert lib.p2v_last_error() is not None
