"""Shared builders for the parity tests: small seeded problems, tapes, comparisons."""
import numpy as np

import oracle as orc
from keypoint_moseq_b200.synth import sample_dataset


def small_problem(seed=0, recordings=2, frames=260, k=5, D=2, d=4, L=3, K=12, seg_length=150, kappa=1e4):
    """Ragged batch: with frames=260 and seg_length=150 each recording gives one full row and one
    short row (mask tail 0)."""
    data, metadata, model = sample_dataset(recordings=recordings, frames=frames, k=k, D=D, d=d, L=L, K=K,
                                           seed=seed, seg_length=seg_length, kappa=kappa)
    return data, metadata, model


def tape_for(data, model, seed=1):
    N, T, k, D = data["Y"].shape
    d = model["states"]["x"].shape[-1]
    L = T - model["states"]["z"].shape[1]
    K = model["params"]["pi"].shape[0]
    return orc.make_tape(np.random.default_rng(seed), N, T, k, D, d, L, K)


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def oracle_sweep(data, model, tape, **flags):
    st, pr, logZ = orc.resample_model(data, model["states"], model["params"], model["hypparams"],
                                      model["noise_prior"], tape, **flags)
    return st, pr, logZ
