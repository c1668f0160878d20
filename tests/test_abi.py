"""The C-ABI library loads and exports every symbol include/b2slam.h declares (no compute here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "b2slam.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b2_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_documented_surface(b2):
    syms = declared_symbols()
    assert len(syms) >= 25
    assert sorted(b2._lib.EXPORTS) == syms


def test_library_exports_every_declared_symbol(b2):
    import __graft_entry__

    __graft_entry__.build()
    lib = ctypes.CDLL(b2._lib.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} missing from libb2slam.so"
    lib.b2_version.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.b2_version()


def test_result_struct_layout(b2):
    assert ctypes.sizeof(b2.IcpResult) == 16 * 8 + 8 + 8 + 4 + 4
    assert b2._lib.RESULT_DTYPE.itemsize == ctypes.sizeof(b2.IcpResult)


def test_no_cpu_fallback(b2):
    import torch

    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        b2.Engine(0)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "3d-lidar-slam_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "liboracle" not in text, f
