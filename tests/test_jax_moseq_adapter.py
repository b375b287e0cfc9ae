"""Pins oracle/kpms_oracle.py against the reference's own engine the moment it is importable.

The arithmetic of the sweep lives in the un-vendored `jax_moseq` package (reached from
/root/reference/keypoint_moseq/fitting.py:13-16, :25, :536-538, :667-673); neither it nor jax can be installed
in the build container, so every test here SKIPS today and the oracle's header says "parity unpinned".  They
compare the deterministic sub-quantities of the sweep - nothing that depends on an RNG contract - so that one
`pip install jax-moseq` turns "unpinned" into a verdict.  Names verified from the reference's import lines are
used directly; names recalled from upstream's layout go through `_upstream`, which skips the single check if
the attribute has moved.  Conventions the oracle had to fix without being able to look
are named constants of the oracle (EPS_SHIFT, X_PRIOR_VAR, SCALE_DOF, OBSVAR_DOF) so a failure here is settled
by changing one line there and the matching KPMS_* macro in csrc/common.cuh.
"""
import importlib
import inspect
import re

import numpy as np
import pytest

jax = pytest.importorskip("jax")
jax_moseq = pytest.importorskip("jax_moseq")
jax.config.update("jax_enable_x64", True)          # the reference's default (keypoint_moseq/__init__.py:1-4)
import jax.numpy as jnp  # noqa: E402

import oracle as orc  # noqa: E402
from helpers import small_problem, tape_for  # noqa: E402

TOL = 1e-8


def _upstream(module, name):
    try:
        return getattr(importlib.import_module(module), name)
    except (ImportError, AttributeError) as e:
        pytest.skip(f"{module}.{name} not found in this jax_moseq ({e})")


@pytest.fixture(scope="module")
def problem():
    data, _, model = small_problem(seed=41, d=4, L=3, K=8, k=6, D=2, frames=240, seg_length=150)
    return data, model, tape_for(data, model)


def test_center_embedding_spans_the_same_subspace():
    """Verified import (viz.py:20).  The basis may differ by an orthogonal factor: compare projectors."""
    from jax_moseq.models.keypoint_slds import center_embedding
    for k in (4, 7, 12):
        G, Go = np.asarray(center_embedding(k)), orc.center_embedding(k)
        assert G.shape == Go.shape == (k, k - 1)
        np.testing.assert_allclose(G @ G.T, Go @ Go.T, atol=1e-12)
        np.testing.assert_allclose(np.abs(G), np.abs(Go), atol=1e-12)      # SVD-based upstream: same vectors up to sign


def test_ar_log_likelihood(problem):
    data, model, _ = problem
    fn = _upstream("jax_moseq.utils.autoregression", "ar_log_likelihood")
    st, pr = model["states"], model["params"]
    ref = np.stack([np.asarray(fn(jnp.asarray(st["x"]), (jnp.asarray(pr["Ab"][j]), jnp.asarray(pr["Q"][j]))))
                    for j in range(pr["Ab"].shape[0])], -1)
    np.testing.assert_allclose(orc.ar_log_likelihood(st["x"], pr["Ab"], pr["Q"]), ref, rtol=TOL, atol=TOL)


def test_marginal_log_likelihood_and_stateseq_marginals(problem):
    """Verified imports and call shapes (fitting.py:14, :536-538, :667-673)."""
    from jax_moseq.models.arhmm import marginal_log_likelihood, stateseq_marginals
    data, model, _ = problem
    st, pr = model["states"], model["params"]
    mll = float(marginal_log_likelihood(jnp.asarray(data["mask"]), jnp.asarray(st["x"]), jnp.asarray(pr["Ab"]),
                                        jnp.asarray(pr["Q"]), jnp.asarray(pr["pi"])))
    ours = orc.marginal_log_likelihood(data["mask"].astype(float), st["x"], pr["Ab"], pr["Q"], pr["pi"])
    # (vi) of the open points: sum over chains here; a per-frame mean upstream would show up as this ratio
    assert abs(ours - mll) < 1e-7 * abs(mll), (ours, mll, ours / mll, float(data["mask"].sum()))
    marg = np.asarray(stateseq_marginals(jnp.asarray(st["x"]), jnp.asarray(data["mask"]),
                                         **{key: jnp.asarray(val) for key, val in pr.items() if key in ("Ab", "Q", "pi")}))
    ref = orc.stateseq_marginals(st["x"], data["mask"].astype(float), pr["Ab"], pr["Q"], pr["pi"])
    valid = data["mask"][:, -ref.shape[1]:] > 0
    np.testing.assert_allclose(marg[valid], ref[valid], atol=1e-8)


def test_ar_to_lds_conventions(problem):
    """Open points (ii) and (iv): noise on the shifted blocks, placement of the lags inside the companion matrix."""
    data, model, _ = problem
    fn = _upstream("jax_moseq.utils.kalman", "ar_to_lds")
    pr = model["params"]
    d = pr["Ab"].shape[1]
    out = fn(jnp.asarray(pr["Ab"][..., :-1]), jnp.asarray(pr["Ab"][..., -1]), jnp.asarray(pr["Q"]))
    A_up, b_up, Q_up = (np.asarray(o) for o in out[:3])
    A, b, Qa = orc.ar_to_lds(pr["Ab"], pr["Q"], 0.0)
    np.testing.assert_allclose(A_up, A, atol=1e-12)
    np.testing.assert_allclose(b_up, b, atol=1e-12)
    np.testing.assert_allclose(Q_up, Qa, atol=1e-12, err_msg=f"oracle EPS_SHIFT = {orc.EPS_SHIFT}")
    assert A.shape[-1] == pr["Ab"].shape[-1] - 1 and d * (A.shape[-1] // d) == A.shape[-1]


def test_kalman_filter_moments(problem):
    """Filtered means and covariances of the lag-augmented state (open points (ii), (iii): prior covariance,
    masked steps)."""
    data, model, _ = problem
    kf = _upstream("jax_moseq.utils.kalman", "kalman_filter")
    st, pr = model["states"], model["params"]
    N, T, k, D = data["Y"].shape
    d, n = pr["Ab"].shape[1], pr["Ab"].shape[2] - 1
    L = n // d
    Ct = orc.lifted_obs_matrix(pr["Cd"], k, D)
    C = np.zeros((k * D, n))
    C[:, n - d:] = Ct[:, :-1]
    ys = orc.rotate(data["Y"] - st["v"][:, :, None, :], -st["h"]).reshape(N, T, k * D)[:, L - 1:]
    Rs = np.repeat(st["s"] * pr["sigmasq"], D, axis=-1)[:, L - 1:]
    A, B, Qa = orc.ar_to_lds(pr["Ab"], pr["Q"], 1e-3)
    m0, S0 = np.zeros(n), orc.X_PRIOR_VAR * np.eye(n)
    fm, fS = orc.kalman_filter(ys, data["mask"][:, L - 1:], st["z"], m0, S0, A, B, Qa, C, Ct[:, -1], Rs)
    sig = inspect.signature(kf)
    for row in range(N):
        args = dict(ys=jnp.asarray(ys[row]), mask=jnp.asarray(data["mask"][row, L - 1:]), zs=jnp.asarray(st["z"][row]),
                    m0=jnp.asarray(m0), S0=jnp.asarray(S0), A=jnp.asarray(A), B=jnp.asarray(B), Q=jnp.asarray(Qa),
                    C=jnp.asarray(C), D=jnp.asarray(Ct[:, -1]), Rs=jnp.asarray(Rs[row]))
        if not set(sig.parameters) <= set(args):
            pytest.skip(f"kalman_filter signature changed: {sig}")
        out = kf(**{key: args[key] for key in sig.parameters})
        up_m, up_S = np.asarray(out[-2]), np.asarray(out[-1])
        valid = data["mask"][row, L - 1:] > 0
        np.testing.assert_allclose(up_m[valid], fm[row][valid], rtol=1e-7, atol=1e-8,
                                   err_msg=f"oracle X_PRIOR_VAR = {orc.X_PRIOR_VAR}")
        np.testing.assert_allclose(up_S[valid], fS[row][valid], rtol=1e-7, atol=1e-9)


def test_regression_posterior_parameters(problem):
    """MNIW posterior (M_n, K_n, S_n, nu_n) from the sufficient statistics (open point (i))."""
    data, model, _ = problem
    st, pr, ah = model["states"], model["params"], model["hypparams"]["ar_hypparams"]
    K = pr["Ab"].shape[0]
    G = orc.ar_suffstats(st["x"], st["z"], data["mask"], K)
    ours = [orc.mniw_posterior(G[j], ah["nu_0"], ah["S_0"], ah["M_0"], ah["K_0"]) for j in range(K)]
    fn = _upstream("jax_moseq.models.arhmm.gibbs", "_resample_regression_params")
    src = inspect.getsource(fn)
    # the upstream routine draws inside; what can be compared without an RNG contract is its algebra, which the
    # source must contain in this form (K_n from K_0^-1 + S_xx, M_n = (M_0 K_0^-1 + S_yx) K_n)
    assert re.search(r"K_0_inv|inv\(K_0\)|solve\(K_0", src), "upstream MNIW update no longer inverts K_0"
    for M_n, K_n, S_n, nu_n in ours:
        assert np.all(np.linalg.eigvalsh(S_n) > 0) and np.all(np.linalg.eigvalsh(K_n) > 0) and nu_n >= ah["nu_0"]


def test_degrees_of_freedom_conventions():
    """Open point (vii): upstream may hard-code 3 where the oracle uses the keypoint dimension."""
    mod = importlib.import_module("jax_moseq.models.keypoint_slds.gibbs")
    for name, const in (("resample_scales", "SCALE_DOF"), ("resample_obs_variance", "OBSVAR_DOF")):
        fn = getattr(mod, name, None)
        if fn is None:
            pytest.skip(f"{name} not found")
        src = inspect.getsource(fn)
        hard3 = re.search(r"\+\s*3\b|3\s*\*", src) is not None
        ours = getattr(orc, const)
        assert (ours == 3) == hard3, (f"{name}: upstream {'hard-codes 3' if hard3 else 'uses the data dimension'}, "
                                      f"oracle.{const} = {ours!r} (None = keypoint dimension D)")
