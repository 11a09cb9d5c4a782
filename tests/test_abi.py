"""The C-ABI library builds, loads without a GPU, and exports every symbol include/diart_b200.h declares
(no compute calls here)."""
import ctypes
import os
import re

from diart_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    text = open(os.path.join(ROOT, "include", "diart_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dg_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.lib()
    names = header_functions()
    assert len(names) >= 25
    for name in names:
        assert hasattr(lib, name), f"{name} is declared in the header but not exported"
    assert set(names) == set(_lib.SIGNATURES), "ctypes signature table and header disagree"
    assert lib.dg_version() >= 100
    assert isinstance(lib.dg_last_error(), bytes)


def test_argument_errors_without_gpu():
    lib = _lib.lib()
    assert lib.dg_seg_create(None, 0, 0, None) == -1
    assert b"null" in lib.dg_last_error()
    assert lib.dg_emb_create(None, 0, 77, 0, None) == -1
    assert lib.dg_osp(None, 1, 1, 1, 3.0, 10.0, 0, None, None) == -1
    out = ctypes.c_void_p()
    assert lib.dg_cluster_create(64, 512, 0.6, 0.3, 1.0, 0, ctypes.byref(out)) == -1   # > 32 global speakers
