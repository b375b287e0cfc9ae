"""Golden fixtures (tests/golden/*.npz, made by tests/golden/make_golden.py): the oracle must keep
reproducing them (CPU), and the CUDA path must match them (GPU)."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from helpers import rel_err, small_problem, tape_for  # noqa: E402
from make_golden import CASES, run_case  # noqa: E402


def _load(name):
    with np.load(os.path.join(HERE, "golden", name + ".npz")) as f:
        return {k: f[k] for k in f.files}


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_reproduces_golden(name):
    gold, out = _load(name), run_case(CASES[name])
    assert np.array_equal(out["z"], gold["z"])
    for key in ("x", "v", "h", "s", "Ab", "Q", "betas", "pi", "sigmasq", "logZ"):
        np.testing.assert_allclose(out[key], gold[key], rtol=1e-9, atol=1e-11, err_msg=key)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_cuda_matches_golden(name):
    import torch
    from keypoint_moseq_b200 import gibbs
    spec, gold = CASES[name], _load(name)
    data, _, model = small_problem(**spec["problem"])
    tape = tape_for(data, model, seed=spec["tape_seed"])
    dd = gibbs.to_device_data(data, "cuda", torch.float64)
    dm = gibbs.to_device_model(model, "cuda", torch.float64)
    out = gibbs.resample_model(dd, **dm, draws=tape, **spec["flags"])
    assert np.array_equal(out["states"]["z"].cpu().numpy(), gold["z"])
    for key in ("Ab", "Q", "betas", "pi", "sigmasq"):
        assert rel_err(out["params"][key].cpu().numpy(), gold[key]) < 1e-7, key
    if not spec["flags"].get("ar_only"):
        assert rel_err(out["states"]["x"].cpu().numpy(), gold["x"]) < 1e-7
        assert rel_err(out["states"]["v"].cpu().numpy(), gold["v"]) < 1e-7
        assert np.abs(out["states"]["s"].cpu().numpy() / gold["s"] - 1).max() < 1e-6
