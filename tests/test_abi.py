"""CPU: the C-ABI library loads and exports every symbol include/x264cu.h declares; without a
GPU the product path fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "x264-i386pic_b200"))
LIB = os.path.join(ROOT, "x264-i386pic_b200", "libx264cuda.so")


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "x264cu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(x264cu_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(LIB):
        sys.path.insert(0, ROOT)
        import __graft_entry__
        __graft_entry__.build()
    return C.CDLL(LIB)


def test_exports_every_declared_symbol(lib):
    syms = declared_symbols()
    assert len(syms) >= 30
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing


def test_binding_matches_header():
    import x264cu
    L = x264cu.load()
    for s in declared_symbols():
        getattr(L, s)


def test_open_fails_loudly_without_a_device(lib):
    import x264cu
    if lib.x264cu_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(RuntimeError, match="no CUDA device"):
        x264cu.Context(640, 360)


def test_bad_configuration_is_rejected(lib):
    import x264cu
    cfg = x264cu.Cfg(8, 8, 3, 7, 1, 16, 512, 1, 1, 1, 0, 2, 16, None)
    lib.x264cu_open.restype = C.c_void_p
    assert not lib.x264cu_open(0, C.byref(cfg))
    lib.x264cu_last_error.restype = C.c_char_p
    assert b"bad configuration" in lib.x264cu_last_error()


def test_product_does_not_reference_the_oracle():
    """the shipped sources must not include, link or load anything under oracle/."""
    pkg = os.path.join(ROOT, "x264-i386pic_b200")
    for dirpath, _, files in os.walk(pkg):
        if "build" in dirpath:
            continue
        for f in files:
            if f.endswith((".cu", ".cuh", ".py", ".h", ".c", ".cpp")) or f == "Makefile":
                text = open(os.path.join(dirpath, f)).read()
                assert "x264_oracle" not in text and "libx264oracle" not in text and "libx264ref" not in text, f
