"""Parity at the benchmark's own shapes and chain length.

One recording of 15 000 frames segmented at 10 000 gives the two kinds of rows every BASELINE config is
made of: a full row of T = 10 030 frames and a ragged row (5 000 valid frames, padded tail).  The whole
sweep runs in float32 with the default time chunking (about 39 speculative chunks per chain, as in
bench.py) against the float64 oracle on the same injected draws.  Bars (BASELINE.json north_star):
labels bit-exact; continuous latents, centroid, noise scales and the marginal log-likelihood within
1e-4 relative; heading within 1e-4 rad.  Every case appends its measured errors to
gpurun_out/parity_report.jsonl when that directory exists.
"""
import json
import os

import numpy as np
import pytest
import torch

import oracle as orc
from helpers import rel_err, tape_for
from keypoint_moseq_b200.synth import sample_dataset
from test_gpu_parity import _cast_problem, _np, _to_dev

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHAPES = {
    "C2": dict(d=10, L=3, K=100, k=12, D=2),      # full 2D model, the bench workload
    "C3": dict(d=10, L=3, K=100, k=16, D=3),      # full 3D model
    "C1": dict(d=4, L=3, K=100, k=10, D=2),       # the AR-only configuration's shape
}


def _report(row):
    out = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, "parity_report.jsonl"), "a") as fh:
            fh.write(json.dumps(row) + "\n")


def _problem(shape, seed=21):
    data, _, model = sample_dataset(recordings=1, frames=15_000, seg_length=10_000, seed=seed, kappa=1e4, **shape)
    assert data["Y"].shape[1] == 10_030 and (data["mask"].sum(1) == [10_030, 5_000]).all()
    tape = tape_for(data, model, seed=seed + 1)
    return _cast_problem(data, model, tape, torch.float32)


@pytest.mark.parametrize("name", ["C2", "C3", "C1"])
@pytest.mark.parametrize("flags", [dict(), dict(states_only=True)], ids=["full", "states_only"])
def test_sweep_at_benchmark_length_float32(name, flags):
    from keypoint_moseq_b200 import gibbs as g
    shape = SHAPES[name]
    data, model, tape = _problem(shape)
    st_ref, pr_ref, _ = orc.resample_model(data, model["states"], model["params"], model["hypparams"],
                                           model["noise_prior"], tape, **flags)
    dd, dm = _to_dev(data, model, torch.float32)
    out = g.resample_model(dd, **dm, draws=tape, **flags)
    torch.cuda.synchronize()
    kal, hmm = g.chunk_diagnostics("kalman_ws"), g.chunk_diagnostics("hmm_ws")
    st = {key: _np(val) for key, val in out["states"].items()}
    mask = data["mask"] > 0
    dh = np.angle(np.exp(1j * (st["h"].astype(np.float64) - st_ref["h"])))
    errs = {
        "z_mismatches": int((st["z"] != st_ref["z"]).sum()),
        "x": rel_err(st["x"], st_ref["x"]), "v": rel_err(st["v"], st_ref["v"]),
        "s": float(np.abs(st["s"] / st_ref["s"] - 1).max()),
        "h_rad": float(np.abs(dh).max()), "h_rad_valid": float(np.abs(dh[mask]).max()),
    }
    for key in ("Ab", "Q", "betas", "pi"):
        errs[key] = rel_err(_np(out["params"][key]), pr_ref[key])
    # marginal log-likelihood of the new latents under the new parameters (float64 path, float32 latents)
    mll = g.marginal_log_likelihood(dd["mask"], out["states"]["x"], out["params"]["Ab"], out["params"]["Q"],
                                    out["params"]["pi"]).item()
    mll_ref = orc.marginal_log_likelihood(data["mask"].astype(float), st_ref["x"], pr_ref["Ab"], pr_ref["Q"], pr_ref["pi"])
    errs["mll"] = abs(mll - mll_ref) / abs(mll_ref)
    _report({"test": "sweep_at_benchmark_length", "shape": name, "flags": flags, "errors": errs, "kalman": kal, "hmm": hmm})
    assert errs["z_mismatches"] == 0, errs
    assert errs["x"] < 1e-4 and errs["v"] < 1e-4 and errs["s"] < 1e-3 and errs["mll"] < 1e-4, errs
    assert errs["h_rad"] < 1e-4, errs
    for key in ("Ab", "Q", "betas", "pi"):
        assert errs[key] < 1e-7, errs
    # the time-parallel path really ran: more than one chunk, nothing fell back on this converged model
    assert kal["forward_rerun"] == 0 and kal["backward_rerun"] == 0, kal


def test_ar_only_sweep_at_c1_shape():
    """BASELINE configs[0]: the AR-HMM-only sweep at latent_dim 4 / 100 states on full-length chains."""
    from keypoint_moseq_b200 import gibbs as g
    data, model, tape = _problem(SHAPES["C1"], seed=33)
    st_ref, pr_ref, _ = orc.resample_model(data, model["states"], model["params"], model["hypparams"],
                                           model["noise_prior"], tape, ar_only=True)
    dd, dm = _to_dev(data, model, torch.float32)
    out = g.resample_model(dd, **dm, draws=tape, ar_only=True)
    hmm = g.chunk_diagnostics("hmm_ws")
    errs = {"z_mismatches": int((_np(out["states"]["z"]) != st_ref["z"]).sum())}
    for key in ("Ab", "Q", "betas", "pi"):
        errs[key] = rel_err(_np(out["params"][key]), pr_ref[key])
    _report({"test": "ar_only_c1", "errors": errs, "hmm": hmm})
    assert errs["z_mismatches"] == 0, (errs, hmm)
    assert max(errs[key] for key in ("Ab", "Q", "betas", "pi")) < 1e-7, errs
    for key in ("x", "v", "h", "s"):                       # untouched by an AR-only sweep
        assert torch.equal(out["states"][key], dm["states"][key]), key
