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


def test_supported_dims_table_and_up_front_validation():
    """latent_dim / nlags / num_states are user configuration (keypoint_moseq/io.py:72-83): the compiled pairs are
    listed by the library and anything else is refused before a kernel is launched, with the table in the message."""
    import __graft_entry__
    __graft_entry__.build()
    from keypoint_moseq_b200 import _lib
    pairs = _lib.supported_dims()
    assert len(pairs) == len(set(pairs)) >= 30
    for d in range(2, 17):
        assert (d, 3) in pairs                       # every latent_dim 2..16 at the reference's default nlags
    for want in [(10, 3), (4, 3), (2, 2), (16, 3), (10, 2), (10, 4), (4, 1)]:
        assert want in pairs
        _lib.check_model_dims(want[0], want[1], 100)
    with pytest.raises(_lib.KpmsError, match="not compiled"):
        _lib.check_model_dims(9, 2, 100)
    with pytest.raises(_lib.KpmsError, match="num_states"):
        _lib.check_model_dims(10, 3, 0)
