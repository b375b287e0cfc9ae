"""Generates the committed golden fixtures from the float64 oracle.

    python tests/golden/make_golden.py

Each fixture stores the seeds that rebuild the inputs (synthetic data + tape) and the oracle's outputs
for one taped Gibbs sweep, so that (a) the oracle is pinned against accidental change and (b) the CUDA
path can be checked without re-running the oracle.  The reference itself has no golden vectors
(SURVEY.md section 4), so these are this repo's own.
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

import oracle as orc  # noqa: E402
from helpers import small_problem, tape_for  # noqa: E402

CASES = {
    "sweep_d4_K12_2d": dict(problem=dict(seed=21, d=4, L=3, K=12, k=5, D=2, kappa=1e2), tape_seed=31, flags={}),
    "sweep_d4_K8_3d_global_noise": dict(problem=dict(seed=22, d=4, L=3, K=8, k=6, D=3, kappa=1e2), tape_seed=32,
                                        flags=dict(resample_global_noise_scale=True)),
    "sweep_d10_K100_2d": dict(problem=dict(seed=23, d=10, L=3, K=100, k=12, D=2, kappa=1e4), tape_seed=33, flags={}),
    "sweep_d4_K12_ar_only": dict(problem=dict(seed=24, d=4, L=3, K=12, k=5, D=2, kappa=1e6), tape_seed=34,
                                 flags=dict(ar_only=True)),
}


def run_case(spec):
    data, _, model = small_problem(**spec["problem"])
    tape = tape_for(data, model, seed=spec["tape_seed"])
    st, pr, logZ = orc.resample_model(data, model["states"], model["params"], model["hypparams"],
                                      model["noise_prior"], tape, **spec["flags"])
    out = {"z": st["z"].astype(np.int32), "logZ": logZ}
    for key in ("x", "v", "h", "s"):
        out[key] = st[key]
    for key in ("Ab", "Q", "betas", "pi", "sigmasq"):
        out[key] = pr[key]
    return out


def main():
    src = open(os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle", "kpms_oracle.py"), "rb").read()
    sha = hashlib.sha256(src).hexdigest()
    for name, spec in CASES.items():
        out = run_case(spec)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), oracle_sha256=np.asarray(sha), **out)
        print(name, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
