"""Joint-distribution test of the oracle's whole sweep (after Geweke, "Getting it right", JASA 2004).

The reference ships no golden vectors for the sweep and jax_moseq is absent (oracle header: PARITY
UNPINNED), so the strongest statement available about `oracle.resample_model` is that it is a correct
Gibbs sampler for the keypoint-SLDS generative model the reference documents:

  betas ~ Dir(gamma/K), pi_i ~ Dir(alpha betas + kappa e_i), (Ab_j, Q_j) ~ MNIW(nu_0, S_0, M_0, K_0),
  sigmasq_k ~ nu_sigma sigmasq_0 / chi2(nu_sigma), z_0 ~ uniform, z_t ~ pi[z_{t-1}],
  x_0 ~ N(0, 10 I), x_t ~ N(A_z x_{t-1} + b_z, Q_z), s_tk ~ nu_s s_0 / chi2(nu_s), h_t ~ U(-pi, pi),
  v_0 ~ N(0, 1e6 I), v_t ~ N(v_{t-1}, sigmasq_loc I), Y_tk ~ N(R(h_t) Ybar_k(x_t) + v_t, s_tk sigmasq_k I).

Draw (parameters, states) from this prior and Y from the likelihood: the pair is a draw from the joint, so
(parameters, states) is a draw from the posterior given Y, and a sweep whose nine conditionals are all right
leaves it one.  Every function g of parameters and states therefore has E[g(after m sweeps) - g(before)] = 0.
Replicates are independent, so the paired z-score has an exact standard error (Geweke's original
successive-conditional chain tests the same invariance but mixes slowly here and needs batch-means errors).
A wrong degree of freedom, a transposed matrix, a missing posterior term or an off-by-one in time in any
resampler moves some functional far outside +-4.5 (second test: power).
One lag and no jitter, because the lag-augmented Kalman model with noisy copies (EPS_SHIFT, jitter) is an
approximation of the AR process by design and not the same joint model.  Prior draws use NumPy / SciPy
generators, not the oracle's taped samplers."""
import numpy as np
import pytest
from scipy.stats import invwishart

import oracle as orc

N, T, k, D, d, L, K = 1, 5, 3, 2, 2, 1, 2
HYP = {
    "trans_hypparams": {"num_states": K, "alpha": 2.0, "kappa": 1.5, "gamma": 3.0},
    "ar_hypparams": {"nu_0": d + 4.0, "S_0": 0.5 * np.eye(d), "M_0": np.hstack([0.6 * np.eye(d), np.zeros((d, 1))]),
                     "K_0": 0.5 * np.eye(d * L + 1)},
    "obs_hypparams": {"nu_sigma": 6.0, "sigmasq_0": 0.8, "nu_s": 5.0},
    "cen_hypparams": {"sigmasq_loc": 0.5},
}
S_PRIOR = 1.3        # noise_prior (s_0)


def draw_prior(rng, Cd):
    th, ah, oh, ch = (HYP[g] for g in ("trans_hypparams", "ar_hypparams", "obs_hypparams", "cen_hypparams"))
    betas = rng.dirichlet(np.full(K, th["gamma"] / K))
    pi = np.stack([rng.dirichlet(th["alpha"] * betas + th["kappa"] * np.eye(K)[i]) for i in range(K)])
    Q = np.stack([np.atleast_2d(invwishart.rvs(df=ah["nu_0"], scale=ah["S_0"], random_state=rng)) for _ in range(K)])
    LK = np.linalg.cholesky(ah["K_0"])
    Ab = np.stack([ah["M_0"] + np.linalg.cholesky(Q[j]) @ rng.standard_normal((d, d * L + 1)) @ LK.T for j in range(K)])
    sigmasq = oh["nu_sigma"] * oh["sigmasq_0"] / rng.chisquare(oh["nu_sigma"], size=k)
    z = np.empty((N, T - L), dtype=np.int64)
    x = np.empty((N, T, d))
    for n in range(N):
        x[n, 0] = np.sqrt(orc.X_PRIOR_VAR) * rng.standard_normal(d)
        for t in range(T - L):
            z[n, t] = rng.integers(K) if t == 0 else rng.choice(K, p=pi[z[n, t - 1]])
            j = z[n, t]
            x[n, t + 1] = Ab[j, :, :-1] @ x[n, t] + Ab[j, :, -1] + np.linalg.cholesky(Q[j]) @ rng.standard_normal(d)
    s = oh["nu_s"] * S_PRIOR / rng.chisquare(oh["nu_s"], size=(N, T, k))
    h = rng.uniform(-np.pi, np.pi, size=(N, T))
    v = np.empty((N, T, D))
    v[:, 0] = np.sqrt(orc.V_PRIOR_VAR) * rng.standard_normal((N, D))
    for t in range(1, T):
        v[:, t] = v[:, t - 1] + np.sqrt(ch["sigmasq_loc"]) * rng.standard_normal((N, D))
    states = {"x": x, "z": z, "s": s, "h": h, "v": v}
    params = {"betas": betas, "pi": pi, "Ab": Ab, "Q": Q, "sigmasq": sigmasq, "Cd": Cd}
    return states, params


def draw_data(rng, states, params):
    mean = orc.estimate_coordinates(states["x"], states["v"], states["h"], params["Cd"], k, D)
    std = np.sqrt(states["s"] * params["sigmasq"])[..., None]
    return mean + std * rng.standard_normal(mean.shape)


def functionals(states, params, Y):
    x, z, s, h, v = (states[n] for n in ("x", "z", "s", "h", "v"))
    Ab, Q, pi, betas, sig = (params[n] for n in ("Ab", "Q", "pi", "betas", "sigmasq"))
    dv = np.diff(v, axis=1)
    ldq = np.log(np.linalg.det(Q))
    out = {
        "log sig": np.log(sig).mean(), "log sig^2": (np.log(sig) ** 2).mean(),
        "log s": np.log(s).mean(), "log s^2": (np.log(s) ** 2).mean(),
        "logdet Q": ldq.mean(), "logdet Q0*Q1": ldq[0] * ldq[1],
        "A00": Ab[:, 0, 0].mean(), "A10": Ab[:, 1, 0].mean(), "b0": Ab[:, 0, -1].mean(), "Ab^2": (Ab ** 2).mean(),
        "A00*logdetQ": (Ab[:, 0, 0] * ldq).mean(),
        "pi diag": np.diag(pi).mean(), "pi01": pi[0, 1], "beta0": betas[0], "beta0^2": betas[0] ** 2,
        "beta0*pi10": betas[0] * pi[1, 0],
        "occ z=0": (z == 0).mean(), "switch rate": (z[:, 1:] != z[:, :-1]).mean(),
        "z first==0": float(z[0, 0] == 0), "z last==0": float(z[0, -1] == 0),
        "sin h": np.sin(h).mean(), "cos 2h": np.cos(2 * h).mean(),
        "dv^2": (dv ** 2).mean(), "dv lag1": (dv[:, 1:] * dv[:, :-1]).mean(),
        "log s * log sig": (np.log(s).mean((0, 1)) * np.log(sig)).mean(),
        "log s * resid": (np.log(s) * np.log1p(x[:, :, :1] ** 2)).mean(),
    }
    # fit of the assigned label / scale: standardised residuals have a known law under the joint
    sq = orc.compute_squared_error(Y, x, v, h, params["Cd"]) / (s * sig)
    out["obs resid"] = np.log1p(sq).mean()
    out["obs resid * log s"] = (np.log1p(sq) * np.log(s)).mean()
    for t in range(T - 1):
        for name, j in (("own", z[0, t]), ("other", 1 - z[0, t])):
            r = x[0, t + 1] - Ab[j, :, :-1] @ x[0, t] - Ab[j, :, -1]
            out[f"ar resid {name} {t}"] = np.log1p(r @ np.linalg.solve(Q[j], r))
    out["logdet Q[z]"] = ldq[z[0]].mean()
    out["logdet Q[z last]"] = ldq[z[0, -1]]
    out["log pi[z,z']"] = np.log(pi[z[0, :-1], z[0, 1:]]).mean()
    for t in range(T):                                            # per frame: catches off-by-one in time
        out[f"log1p x{t}^2"] = np.log1p(x[:, t] ** 2).mean()
        out[f"tanh x{t}"] = np.tanh(x[:, t] / 3).mean()
        out[f"cos h{t}"] = np.cos(h[:, t]).mean()
        out[f"log s{t}"] = np.log(s[:, t]).mean()
    for t in range(T - 1):
        out[f"log1p dx{t}^2"] = np.log1p((x[:, t + 1] - x[:, t]) ** 2).mean()
        out[f"dx{t} given z"] = (np.tanh(x[:, t + 1, 0] - x[:, t, 0]) * (2.0 * (z[:, t] == 0) - 1.0)).mean()
    return out


def invariance_z_scores(replicates, sweeps, seed, **sweep_kwargs):
    rng = np.random.default_rng(seed)
    Cd = np.random.default_rng(99).standard_normal(((k - 1) * D, d + 1))
    mask = np.ones((N, T))
    kw = dict(resample_global_noise_scale=True, jitter=0.0)
    kw.update(sweep_kwargs)
    diffs, names = [], None
    for _ in range(replicates):
        states, params = draw_prior(rng, Cd)
        Y = draw_data(rng, states, params)
        before = functionals(states, params, Y)
        for _ in range(sweeps):
            tape = orc.make_tape(rng, N, T, k, D, d, L, K)
            states, params, _ = orc.resample_model({"Y": Y, "mask": mask}, states, params, HYP, S_PRIOR, tape, **kw)
        after = functionals(states, params, Y)
        names = list(before)
        diffs.append([after[n] - before[n] for n in names])
    diffs = np.array(diffs)
    zs = diffs.mean(0) / (diffs.std(0, ddof=1) / np.sqrt(replicates))
    return dict(zip(names, zs))


def test_whole_sweep_leaves_the_joint_distribution_invariant():
    zs = invariance_z_scores(replicates=4000, sweeps=2, seed=7)
    worst = max(zs, key=lambda n: abs(zs[n]))
    assert abs(zs[worst]) < 4.5, {n: round(float(v), 2) for n, v in zs.items() if abs(v) > 3}
    assert len(zs) == 31 + 4 * T + 4 * (T - 1)


@pytest.mark.parametrize("broken", ["scales_dof", "hmm_shifted"])
def test_invariance_test_has_power_against_a_wrong_conditional(broken, monkeypatch):
    """The same statistic flags sweeps that are subtly wrong: the scale draw with its degrees of freedom off by
    the data dimension; the label sampler reading the likelihood one frame late.  (Errors that only tilt the label
    posterior a little - say a dropped log-determinant - need about four times the replicates at this model size.)"""
    import oracle.kpms_oracle as mod
    if broken == "scales_dof":
        def wrong(Y, x, v, h, Cd, sigmasq, nu_s, s_0, g_s):
            variance = mod.compute_squared_error(Y, x, v, h, Cd) / sigmasq + s_0 * nu_s
            return variance / (2.0 * mod.gamma_mt(np.full(variance.shape, nu_s / 2.0), g_s))
        monkeypatch.setattr(mod, "resample_scales", wrong)
    else:
        real = mod.ar_log_likelihood
        monkeypatch.setattr(mod, "ar_log_likelihood", lambda x, Ab, Q: np.roll(real(x, Ab, Q), 1, axis=1))
    zs = invariance_z_scores(replicates=1500, sweeps=2, seed=8)
    assert max(abs(v) for v in zs.values()) > 6.0, broken


def test_ar_only_sweep_with_two_lags_and_three_states_is_invariant():
    """The AR-HMM part (transitions, AR parameters, labels given x) is an exact Gibbs sampler for any number
    of lags: x generated with Ab laid out [oldest lag | ... | newest lag | offset] - the layout M_0 = identity
    on the newest lag implies - must leave labels and parameters at their joint law under
    `resample_model(ar_only=True)`.  Pins the lag order of `get_lags` against the likelihood AND the
    sufficient statistics at once, and the sticky-HDP update with three states."""
    K3, L2, T2 = 3, 2, 9
    n = d * L2
    M_0 = np.hstack([np.zeros((d, n - d)), 0.5 * np.eye(d), np.zeros((d, 1))])
    hyp = {"trans_hypparams": {"num_states": K3, "alpha": 3.0, "kappa": 2.0, "gamma": 2.0},
           "ar_hypparams": {"nu_0": d + 4.0, "S_0": 0.4 * np.eye(d), "M_0": M_0, "K_0": 0.3 * np.eye(n + 1)},
           "obs_hypparams": HYP["obs_hypparams"], "cen_hypparams": HYP["cen_hypparams"]}
    th, ah = hyp["trans_hypparams"], hyp["ar_hypparams"]
    rng = np.random.default_rng(21)
    LK = np.linalg.cholesky(ah["K_0"])
    mask = np.ones((1, T2))

    def g(z, pr):
        Ab, Q, pi, betas = pr["Ab"], pr["Q"], pr["pi"], pr["betas"]
        ldq = np.log(np.linalg.det(Q))
        out = {"pi diag": np.diag(pi).mean(), "pi01": pi[0, 1], "pi21": pi[2, 1], "beta0": betas[0], "beta2^2": betas[2] ** 2,
               "logdet Q": ldq.mean(), "A old": Ab[:, 0, 0].mean(), "A new": Ab[:, 0, n - d].mean(), "A new off": Ab[:, 0, n - d + 1].mean(),
               "A old^2": (Ab[:, :, :n - d] ** 2).mean(), "A new^2": (Ab[:, :, n - d:n] ** 2).mean(), "b^2": (Ab[:, :, -1] ** 2).mean(),
               "occ0": (z == 0).mean(), "occ2": (z == 2).mean(), "switch": (z[:, 1:] != z[:, :-1]).mean(),
               "logdet Q[z]": ldq[z[0]].mean(), "log pi[z,z']": np.log(pi[z[0, :-1], z[0, 1:]]).mean(),
               "log beta[z]": np.log(betas[z[0]]).mean()}
        for t in range(T2 - L2):
            j = z[0, t]
            phi = np.concatenate([x[0, t], x[0, t + 1], [1.0]])
            r = x[0, t + 2] - Ab[j] @ phi
            out[f"resid own {t}"] = np.log1p(r @ np.linalg.solve(Q[j], r))
        return out

    diffs = []
    R = 5000
    for _ in range(R):
        betas = rng.dirichlet(np.full(K3, th["gamma"] / K3))
        pi = np.stack([rng.dirichlet(th["alpha"] * betas + th["kappa"] * np.eye(K3)[i]) for i in range(K3)])
        Q = np.stack([invwishart.rvs(df=ah["nu_0"], scale=ah["S_0"], random_state=rng) for _ in range(K3)])
        Ab = np.stack([M_0 + np.linalg.cholesky(Q[j]) @ rng.standard_normal((d, n + 1)) @ LK.T for j in range(K3)])
        x = np.empty((1, T2, d))
        z = np.empty((1, T2 - L2), dtype=np.int64)
        x[0, :L2] = rng.standard_normal((L2, d)) * 2.0
        for t in range(T2 - L2):
            z[0, t] = rng.integers(K3) if t == 0 else rng.choice(K3, p=pi[z[0, t - 1]])
            j = z[0, t]
            phi = np.concatenate([x[0, t], x[0, t + 1], [1.0]])           # oldest lag first
            x[0, t + L2] = Ab[j] @ phi + np.linalg.cholesky(Q[j]) @ rng.standard_normal(d)
        pr = {"betas": betas, "pi": pi, "Ab": Ab, "Q": Q}
        before = g(z, pr)
        st = {"x": x, "z": z}
        for _ in range(2):
            tape = orc.make_tape(rng, 1, T2, k, D, d, L2, K3)
            st, pr, _ = orc.resample_model({"Y": None, "mask": mask}, st, pr, hyp, S_PRIOR, tape, ar_only=True)
        after = g(st["z"], pr)
        diffs.append([after[m] - before[m] for m in before])
    diffs = np.array(diffs)
    zs = dict(zip(before, diffs.mean(0) / (diffs.std(0, ddof=1) / np.sqrt(R))))
    worst = max(zs, key=lambda m: abs(zs[m]))
    assert abs(zs[worst]) < 4.5, {m: round(float(v), 2) for m, v in zs.items() if abs(v) > 3}
