"""Pins the float64 oracle against closed-form / brute-force answers (the reference ships no
tests or golden vectors)."""
import itertools

import numpy as np
import pytest

import oracle as orc


def _rand_spd(rng, n, scale=1.0):
    a = rng.standard_normal((n, n))
    return scale * (a @ a.T / n + 0.3 * np.eye(n))


def _dense_joint(ys, omask, zs, m0, S0, A, B, Q, C, D, Rs):
    """Posterior mean/cov of the stacked states by one dense Gaussian solve (single chain)."""
    T, n = ys.shape[0], m0.shape[0]
    Lam = np.zeros((T * n, T * n))
    eta = np.zeros(T * n)
    S0i = np.linalg.inv(S0)
    Lam[:n, :n] += S0i
    eta[:n] += S0i @ m0
    for t in range(T):
        sl = slice(t * n, (t + 1) * n)
        if omask[t]:
            Lam[sl, sl] += (C.T / Rs[t]) @ C
            eta[sl] += (C.T / Rs[t]) @ (ys[t] - D)
        if t < T - 1 and omask[t]:
            nx = slice((t + 1) * n, (t + 2) * n)
            Qi = np.linalg.inv(Q[zs[t]])
            Az, Bz = A[zs[t]], B[zs[t]]
            Lam[sl, sl] += Az.T @ Qi @ Az
            Lam[sl, nx] -= Az.T @ Qi
            Lam[nx, sl] -= Qi @ Az
            Lam[nx, nx] += Qi
            eta[sl] -= Az.T @ Qi @ Bz
            eta[nx] += Qi @ Bz
    cov = np.linalg.inv(Lam)
    return cov @ eta, cov


def test_kalman_sample_matches_dense_gaussian():
    rng = np.random.default_rng(0)
    T, n, m, K = 5, 4, 3, 3
    A = rng.standard_normal((K, n, n)) * 0.4
    B = rng.standard_normal((K, n)) * 0.2
    Q = np.stack([_rand_spd(rng, n, 0.3) for _ in range(K)])
    C = rng.standard_normal((m, n))
    D = rng.standard_normal(m)
    ys = rng.standard_normal((1, T, m))
    Rs = rng.uniform(0.2, 1.5, (1, T, m))
    zs = rng.integers(K, size=(1, T - 1))
    omask = np.ones((1, T), dtype=int)
    m0, S0 = rng.standard_normal(n), _rand_spd(rng, n, 2.0)
    mean, cov = _dense_joint(ys[0], omask[0], zs[0], m0, S0, A, B, Q, C, D, Rs[0])
    x0 = orc.kalman_sample(ys, omask, zs, m0, S0, A, B, Q, C, D, Rs, np.zeros((1, T, n)))
    np.testing.assert_allclose(x0.reshape(-1), mean, rtol=1e-8, atol=1e-10)
    # the map w -> x is affine; its linear part M satisfies M M^T = posterior covariance
    M = np.zeros((T * n, T * n))
    for i in range(T * n):
        w = np.zeros((1, T * n))
        w[0, i] = 1.0
        xi = orc.kalman_sample(ys, omask, zs, m0, S0, A, B, Q, C, D, Rs, w.reshape(1, T, n))
        M[:, i] = xi.reshape(-1) - mean
    np.testing.assert_allclose(M @ M.T, cov, rtol=1e-7, atol=1e-10)


def test_kalman_masked_tail_carries_no_information():
    rng = np.random.default_rng(1)
    T, Tv, n, m, K = 7, 4, 3, 2, 2
    A = rng.standard_normal((K, n, n)) * 0.4
    B = rng.standard_normal((K, n)) * 0.2
    Q = np.stack([_rand_spd(rng, n, 0.3) for _ in range(K)])
    C, D = rng.standard_normal((m, n)), rng.standard_normal(m)
    ys, Rs = rng.standard_normal((1, T, m)), rng.uniform(0.2, 1.5, (1, T, m))
    zs = rng.integers(K, size=(1, T - 1))
    omask = np.zeros((1, T), dtype=int)
    omask[:, :Tv] = 1
    m0, S0 = np.zeros(n), 10 * np.eye(n)
    x = orc.kalman_sample(ys, omask, zs, m0, S0, A, B, Q, C, D, Rs, np.zeros((1, T, n)))
    # valid frames: last valid frame's transition to the first padded frame is the only extra
    # factor and it is marginalised out exactly, so the mean over frames < Tv equals the
    # posterior of the truncated chain
    mean, _ = _dense_joint(ys[0, :Tv], np.ones(Tv, int), zs[0, :Tv - 1], m0, S0, A, B, Q, C, D, Rs[0, :Tv])
    np.testing.assert_allclose(x[0, :Tv].reshape(-1), mean, rtol=1e-8, atol=1e-10)
    # padded frames repeat the carried state
    for t in range(Tv + 1, T):
        np.testing.assert_array_equal(x[0, t], x[0, Tv])


def test_hmm_ffbs_matches_enumeration():
    rng = np.random.default_rng(2)
    T, K, S = 4, 3, 300_000
    pi = rng.dirichlet(np.ones(K), size=K)
    ll1 = rng.standard_normal((T, K)) * 2.0
    logp = {}
    for path in itertools.product(range(K), repeat=T):
        lp = -np.log(K) + ll1[0, path[0]]
        for t in range(1, T):
            lp += np.log(pi[path[t - 1], path[t]]) + ll1[t, path[t]]
        logp[path] = lp
    mx = max(logp.values())
    Z = sum(np.exp(v - mx) for v in logp.values())
    logZ_exact = np.log(Z) + mx
    ll = np.broadcast_to(ll1, (S, T, K)).copy()
    z, logZ = orc.sample_hmm_stateseq(pi, ll, np.ones((S, T)), rng.uniform(size=(S, T)))
    np.testing.assert_allclose(logZ[0], logZ_exact, rtol=1e-12)
    codes = (z * K ** np.arange(T)[::-1]).sum(1)
    freq = np.bincount(codes, minlength=K ** T) / S
    exact = np.array([np.exp(logp[p] - logZ_exact) for p in itertools.product(range(K), repeat=T)])
    assert np.abs(freq - exact).max() < 5 * np.sqrt(exact.max() / S)


def test_stateseq_marginals_match_enumeration():
    rng = np.random.default_rng(3)
    T, K, d, L = 5, 3, 2, 2
    x = rng.standard_normal((1, T + L, d))
    Ab = rng.standard_normal((K, d, d * L + 1)) * 0.3
    Q = np.stack([_rand_spd(rng, d, 0.5) for _ in range(K)])
    pi = rng.dirichlet(np.ones(K), size=K)
    mask = np.ones((1, T + L))
    ll = orc.ar_log_likelihood(x, Ab, Q)[0]
    post = np.zeros((T, K))
    tot = 0.0
    for path in itertools.product(range(K), repeat=T):
        p = np.exp(ll[0, path[0]]) / K
        for t in range(1, T):
            p *= pi[path[t - 1], path[t]] * np.exp(ll[t, path[t]])
        tot += p
        for t in range(T):
            post[t, path[t]] += p
    sm = orc.stateseq_marginals(x, mask, Ab, Q, pi)[0]
    np.testing.assert_allclose(sm, post / tot, rtol=1e-9)
    np.testing.assert_allclose(orc.marginal_log_likelihood(mask, x, Ab, Q, pi), np.log(tot), rtol=1e-10)


def test_ar_log_likelihood_matches_scipy():
    from scipy.stats import multivariate_normal
    rng = np.random.default_rng(4)
    K, d, L, T = 3, 3, 2, 6
    x = rng.standard_normal((2, T, d))
    Ab = rng.standard_normal((K, d, d * L + 1)) * 0.3
    Q = np.stack([_rand_spd(rng, d, 0.5) for _ in range(K)])
    ll = orc.ar_log_likelihood(x, Ab, Q)
    for nn in range(2):
        for t in range(L, T):
            phi = np.concatenate([x[nn, t - 2], x[nn, t - 1]])
            for j in range(K):
                mu = Ab[j, :, :-1] @ phi + Ab[j, :, -1]
                ref = multivariate_normal(mu, Q[j]).logpdf(x[nn, t])
                np.testing.assert_allclose(ll[nn, t - L, j], ref, rtol=1e-10)


def test_gamma_and_vonmises_moments():
    rng = np.random.default_rng(5)
    S = 200_000
    tape = np.empty((S, orc.GAMMA_TAPE))
    tape[:, :orc.GAMMA_R] = rng.standard_normal((S, orc.GAMMA_R))
    tape[:, orc.GAMMA_R:] = rng.uniform(1e-12, 1, (S, orc.GAMMA_R + 1))
    for a in (0.3, 1.0, 3.5, 50.0):
        g = orc.gamma_mt(np.full(S, a), tape)
        assert abs(g.mean() - a) < 5 * np.sqrt(a / S) + 1e-3 * a
        assert abs(g.var() - a) < 0.05 * a
    from scipy.special import i0, i1
    u = rng.uniform(1e-12, 1, (S, orc.VM_R, 3))
    for kappa in (0.5, 4.0, 200.0):
        th = orc.vonmises_bf(np.full(S, 0.7), np.full(S, kappa), u)
        assert np.all((th >= -np.pi) & (th <= np.pi))
        R = np.mean(np.cos(th - 0.7))
        assert abs(R - i1(kappa) / i0(kappa)) < 5e-3
        assert abs(np.mean(np.sin(th - 0.7))) < 5e-3


def test_location_ffbs_matches_dense_gaussian():
    rng = np.random.default_rng(6)
    N, T, k, D, d = 1, 6, 4, 2, 3
    Y = rng.standard_normal((N, T, k, D)) * 3
    x = rng.standard_normal((N, T, d))
    h = rng.uniform(-3, 3, (N, T))
    s = rng.uniform(0.5, 2, (N, T, k))
    Cd = rng.standard_normal(((k - 1) * D, d + 1))
    sigmasq = rng.uniform(0.5, 1.5, k)
    mask = np.ones((N, T), int)
    sl = 0.5
    v = orc.resample_location(Y, mask, x, h, s, Cd, sigmasq, sl, np.zeros((N, T, D)))
    Yrot = orc.rotate(orc.estimate_coordinates(x, np.zeros((N, T, D)), np.zeros((N, T)), Cd, k, D), h)
    prec = 1 / (s * sigmasq)
    Lam = np.zeros((T, T))
    eta = np.zeros((T, D))
    Lam[0, 0] += 1 / orc.V_PRIOR_VAR
    for t in range(T):
        Lam[t, t] += prec[0, t].sum()
        eta[t] += ((Y[0, t] - Yrot[0, t]) * prec[0, t][:, None]).sum(0)
        if t < T - 1:
            Lam[t, t] += 1 / sl
            Lam[t + 1, t + 1] += 1 / sl
            Lam[t, t + 1] -= 1 / sl
            Lam[t + 1, t] -= 1 / sl
    np.testing.assert_allclose(v[0], np.linalg.solve(Lam, eta), rtol=1e-7)


def test_mniw_posterior_mean_identity():
    """With zero noise tapes the draw collapses to M_n and S_n-scaled Q; check M_n against
    the ridge-regression closed form."""
    rng = np.random.default_rng(7)
    d, L, K, N, T = 2, 2, 2, 1, 400
    n = d * L
    x = rng.standard_normal((N, T, d))
    z = rng.integers(K, size=(N, T - L))
    mask = np.ones((N, T), int)
    S_0, K_0 = 0.01 * np.eye(d), 10.0 * np.eye(n + 1)
    M_0 = np.zeros((d, n + 1))
    M_0[:, n - d:n] = np.eye(d)
    G = orc.ar_suffstats(x, z, mask, K)
    tape = np.zeros((d, orc.GAMMA_TAPE))
    tape[:, orc.GAMMA_R:] = 0.5
    Ab, Q = orc.sample_mniw_from_stats(G[0], d + 2, S_0, M_0, K_0, np.zeros((d, n + 1)), np.zeros((d, d)), tape)
    phi = np.concatenate([orc.get_lags(x, L), np.ones((N, T - L, 1))], -1)[0][z[0] == 0]
    y = x[0, L:][z[0] == 0]
    K0i = np.linalg.inv(K_0)
    Mn = (M_0 @ K0i + y.T @ phi) @ np.linalg.inv(K0i + phi.T @ phi)
    np.testing.assert_allclose(Ab, Mn, rtol=1e-9, atol=1e-12)
    assert np.all(np.linalg.eigvalsh(Q) > 0)


def test_transition_resampler_is_a_distribution_and_counts():
    rng = np.random.default_rng(8)
    K, N, T, L = 5, 3, 60, 3
    z = rng.integers(K, size=(N, T - L))
    mask = np.ones((N, T), int)
    mask[1, 40:] = 0
    Nij = orc.count_transitions(z, mask, K)
    assert Nij.sum() == (T - L - 1) * 2 + (40 - L - 1)
    tape = orc.make_tape(rng, N, T, 4, 2, 2, L, K)
    betas, pi = orc.resample_hdp_transitions(z, mask, np.full(K, 1 / K), 5.7, 100.0, 1e3,
                                             tape["u_crp"], tape["u_bin"], tape["g_beta"], tape["g_pi"])
    np.testing.assert_allclose(betas.sum(), 1)
    np.testing.assert_allclose(pi.sum(1), 1)
    assert np.all(np.diag(pi) > 0.5)


def test_padding_content_does_not_reach_valid_frames_or_parameters():
    """Rows are padded to a common length with mask 0 (util.batch).  Whatever sits in the padded frames - the
    keypoints, the noise prior, the old states - must not change any parameter draw nor any state at a valid
    frame (same draws): the likelihood terms, sufficient statistics and filters all gate on the mask."""
    from helpers import oracle_sweep, small_problem, tape_for
    data, _, model = small_problem(seed=3)
    mask = np.asarray(data["mask"]) > 0
    assert (~mask).any()
    tape = tape_for(data, model, seed=5)
    flags = dict(resample_global_noise_scale=True)
    st0, pr0, lz0 = oracle_sweep(data, model, tape, **flags)

    rng = np.random.default_rng(9)
    data2 = {k_: np.array(v, copy=True) for k_, v in data.items()}
    data2["Y"][~mask] = rng.standard_normal(data2["Y"][~mask].shape) * 50.0
    model2 = {"states": {k_: np.array(v, copy=True) for k_, v in model["states"].items()},
              "params": model["params"], "hypparams": model["hypparams"],
              "noise_prior": np.array(model["noise_prior"], copy=True)}
    model2["noise_prior"][~mask] = 7.0
    for name in ("v", "h", "s"):
        pad = model2["states"][name][~mask]
        model2["states"][name][~mask] = np.abs(rng.standard_normal(pad.shape)) + 0.5
    st1, pr1, lz1 = oracle_sweep(data2, model2, tape, **flags)

    for name in ("Ab", "Q", "pi", "betas", "sigmasq"):
        assert np.array_equal(pr0[name], pr1[name]), name
    assert np.array_equal(lz0, lz1)
    L = mask.shape[1] - st0["z"].shape[1]
    assert np.array_equal(st0["z"][mask[:, L:]], st1["z"][mask[:, L:]])
    for name in ("x", "v", "h", "s"):
        assert np.array_equal(st0[name][mask], st1[name][mask]), name
