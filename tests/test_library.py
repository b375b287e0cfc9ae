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


def test_chunk_planner_minimises_waves_times_steps():
    """csrc/capi.cu chunks_for through kpms_plan_chunks: with fewer chains than task slots one wave of slots / N chunks
    (C2: 29 filter chunks); with about as many chains as slots (C4 on one GPU: 1200 chains, 2368 filter slots, 1184 HMM
    slots) several shorter waves instead of one half-empty or two full-length ones; never beyond len / (4 warm-up);
    always within 3 % of the brute-force optimum of the wave model; a forced count wins."""
    import __graft_entry__
    __graft_entry__.build()
    from keypoint_moseq_b200 import _lib
    lib = _lib.load()

    def cost(N, slots, length, W, C):
        return -(-N * C // slots) * (-(-length // C) + (W if C > 1 else 0))

    assert lib.kpms_plan_chunks(80, 148 * 16, 10028, 64) == 29
    assert lib.kpms_plan_chunks(80, 148 * 8, 10027, 64) == 14
    assert lib.kpms_plan_chunks(1200, 148 * 16, 10028, 64) == 7
    assert lib.kpms_plan_chunks(1200, 148 * 8, 10027, 64) == 7
    assert lib.kpms_plan_chunks(1, 148 * 16, 100, 64) == 1                       # too short to cut
    for N in (1, 4, 80, 150, 300, 550, 600, 1200, 5000):
        for slots in (148 * 4, 148 * 8, 148 * 12, 148 * 16):
            for length in (500, 10028, 100_000):
                C = lib.kpms_plan_chunks(N, slots, length, 64)
                cmax = max(1, length // 256)
                assert 1 <= C <= cmax
                best = min(cost(N, slots, length, 64, c) for c in range(1, cmax + 1) if N * c <= 64 * slots or c == 1)
                assert cost(N, slots, length, 64, C) * 100 <= best * 103, (N, slots, length, C)
    _lib.set_time_chunking(chunks=3)
    try:
        assert lib.kpms_plan_chunks(1200, 148 * 16, 10028, 64) == 3
    finally:
        _lib.set_time_chunking(chunks=0)
    assert lib.kpms_plan_chunks(1200, 148 * 16, 10028, 64) == 7
