"""CPU-side checks of the C-ABI boundary: the library loads and exports every symbol that
include/kpms_b200.h declares (no compute calls without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "kpms_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(kpms_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    import __graft_entry__
    __graft_entry__.build()
    from keypoint_moseq_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert sorted(_lib.SIGNATURES) == declared, "ctypes prototypes out of sync with the header"
    assert _lib.load().kpms_version() >= 100


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "keypoint_moseq_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
