"""float64 NumPy restatement of the keypoint-SLDS Gibbs sweep with draw tapes.

TEST INFRASTRUCTURE (see oracle/__init__.py).  PARITY UNPINNED.

Every sampler takes its randomness as explicit arrays ("tapes") of standard
normals / uniforms so that the CUDA kernels can be fed the very same draws.

Upstream being restated (jax-moseq, not vendored by the reference; reached
from /root/reference/keypoint_moseq/fitting.py:13-16, :25, :245-248):
  jax_moseq/models/keypoint_slds/{gibbs,alignment}.py
  jax_moseq/models/slds/gibbs.py
  jax_moseq/models/arhmm/gibbs.py
  jax_moseq/utils/{kalman,autoregression,distributions,transitions}.py
Shapes follow the checkpoint layout the reference reads/writes
(/root/reference/keypoint_moseq/io.py:595-605, :701-711).

Conventions fixed here (DESIGN.md "open points"):
  * Ab = [A | b], lags ordered oldest -> newest, bias last.
  * Q_aug = blkdiag(EPS_SHIFT*I, Q) + jitter*I on all n diagonals.
  * initial augmented state N(0, X_PRIOR_VAR*I) at frame L-1.
  * frame t is conditioned on iff mask[t]==1; the transition t->t+1 is applied
    iff mask[t]==1, otherwise the filter/sampler carries its state through.
  * HMM: masked frames contribute log-likelihood 0, uniform initial distribution.
  * categorical: c=cumsum(p); r=c[-1]*(1-u); z=#{i: c_i<r}.
  * gamma: Marsaglia-Tsang with GAMMA_R taped attempts (fallback d on exhaustion).
  * von Mises: Best-Fisher with VM_R taped attempts (fallback mode on exhaustion).
"""
import numpy as np

EPS_SHIFT = 1e-2      # noise on the shifted (copy) blocks of the augmented AR state
X_PRIOR_VAR = 10.0    # prior variance of the first augmented state
V_PRIOR_VAR = 1e6     # diffuse prior variance of the centroid random walk
GAMMA_R = 6           # taped Marsaglia-Tsang attempts per gamma draw
VM_R = 8              # taped Best-Fisher attempts per von Mises draw
GAMMA_TAPE = 2 * GAMMA_R + 1   # [normals R | uniforms R | boost uniform]
# degrees of freedom a keypoint adds to the scaled-inverse-chi-square posteriors of s and sigmasq: None = its
# dimension D (what the model implies); upstream may hard-code 3 (open point (vii), tests/test_jax_moseq_adapter.py).
# The kernels use KPMS_OBS_DOF in csrc/common.cuh: change both together.
SCALE_DOF = None
OBSVAR_DOF = None

__all__ = [n for n in dir() if n.isupper()] + [
    "center_embedding", "lifted_obs_matrix", "rotate", "estimate_coordinates",
    "get_lags", "ar_log_likelihood", "hmm_filter", "sample_hmm_stateseq",
    "resample_discrete_stateseqs", "marginal_log_likelihood", "stateseq_marginals",
    "ar_to_lds", "kalman_filter", "kalman_sample", "resample_continuous_stateseqs",
    "compute_squared_error", "gamma_mt", "vonmises_bf", "resample_scales",
    "resample_obs_variance", "obs_variance_suffstats", "resample_heading",
    "resample_location", "ar_suffstats", "count_transitions",
    "resample_ar_params", "mniw_posterior", "sample_mniw_from_stats", "resample_hdp_transitions",
    "make_tape", "resample_model",
]


# ----------------------------------------------------------------------------
# alignment (jax_moseq/models/keypoint_slds/alignment.py; used by the reference
# at viz.py:197, util.py:498,508)
# ----------------------------------------------------------------------------
def center_embedding(k):
    """(k, k-1) orthonormal basis of the zero-mean subspace (SVD, as upstream)."""
    return np.linalg.svd(np.eye(k) - np.ones((k, k)) / k)[0][:, :-1]


def lifted_obs_matrix(Cd, k, D):
    """(Gamma kron I_D) @ Cd -> (k*D, d+1): C~ and d~ of the lifted observation model."""
    Gamma = center_embedding(k)
    return np.kron(Gamma, np.eye(D)) @ Cd


def rotate(P, h):
    """Rotate points P (..., k, D) by angle h (...) in the xy-plane: y = R(h) p."""
    c, s = np.cos(h)[..., None], np.sin(h)[..., None]
    out = P.copy()
    out[..., 0] = c * P[..., 0] - s * P[..., 1]
    out[..., 1] = s * P[..., 0] + c * P[..., 1]
    return out


def _ybar(x, Cd, k, D):
    """Centred, aligned pose Ybar (..., k, D) = Gamma . reshape(Cd [x;1])."""
    Ct = lifted_obs_matrix(Cd, k, D)
    flat = x @ Ct[:, :-1].T + Ct[:, -1]
    return flat.reshape(*x.shape[:-1], k, D)


def estimate_coordinates(x, v, h, Cd, k, D):
    """Y = R(h) Ybar(x) + v (reference use: docs/source/advanced.rst:21-26)."""
    return rotate(_ybar(x, Cd, k, D), h) + v[..., None, :]


# ----------------------------------------------------------------------------
# AR-HMM: log-likelihoods and HMM FFBS (jax_moseq/utils/autoregression.py,
# jax_moseq/utils/distributions.py sample_hmm_stateseq, dynamax hmm_filter)
# ----------------------------------------------------------------------------
def get_lags(x, nlags):
    """(..., T, d) -> (..., T-nlags, d*nlags), oldest lag first."""
    T = x.shape[-2]
    return np.concatenate([x[..., i:T - nlags + i, :] for i in range(nlags)], axis=-1)


def ar_log_likelihood(x, Ab, Q):
    """ll[n,t,j] = log N(x_{t+L}; A_j phi + b_j, Q_j), shape (N, T-L, K)."""
    K, d = Ab.shape[0], Ab.shape[1]
    L = Ab.shape[2] // d
    phi = get_lags(x, L)
    y = x[..., L:, :]
    out = np.empty(y.shape[:-1] + (K,))
    for j in range(K):
        mu = phi @ Ab[j, :, :-1].T + Ab[j, :, -1]
        Lq = np.linalg.cholesky(Q[j])
        r = np.linalg.solve(Lq, (y - mu).reshape(-1, d).T).T.reshape(y.shape)
        out[..., j] = (-0.5 * (r ** 2).sum(-1) - np.log(np.diag(Lq)).sum()
                       - 0.5 * d * np.log(2 * np.pi))
    return out


def hmm_filter(pi, ll):
    """Scaled forward filter, batched. ll (N,T,K) -> logZ (N,), filtered (N,T,K)."""
    N, T, K = ll.shape
    pred = np.full((N, K), 1.0 / K)
    filt = np.empty((N, T, K))
    logZ = np.zeros(N)
    for t in range(T):
        mx = ll[:, t].max(-1)
        q = pred * np.exp(ll[:, t] - mx[:, None])
        s = q.sum(-1)
        filt[:, t] = q / s[:, None]
        logZ += np.log(s) + mx
        pred = filt[:, t] @ pi
    return logZ, filt


def _categorical(p, u):
    c = np.cumsum(p, axis=-1)
    r = c[..., -1] * (1.0 - u)
    return (c < r[..., None]).sum(-1)


def sample_hmm_stateseq(pi, ll, mask, u):
    """HMM FFBS. ll (N,T,K), mask (N,T), u (N,T) uniforms -> z (N,T) int, logZ (N,)."""
    ll = ll * mask[..., None]
    logZ, filt = hmm_filter(pi, ll)
    N, T, K = ll.shape
    z = np.empty((N, T), dtype=np.int64)
    z[:, -1] = _categorical(filt[:, -1], u[:, -1])
    for t in range(T - 2, -1, -1):
        p = filt[:, t] * pi[:, z[:, t + 1]].T
        p = p / p.sum(-1, keepdims=True)
        z[:, t] = _categorical(p, u[:, t])
    return z, logZ


def resample_discrete_stateseqs(x, mask, Ab, Q, pi, u_z):
    L = Ab.shape[2] // Ab.shape[1]
    ll = ar_log_likelihood(x, Ab, Q)
    return sample_hmm_stateseq(pi, ll, mask[:, L:].astype(float), u_z)


def marginal_log_likelihood(mask, x, Ab, Q, pi):
    """Sum over chains of the HMM forward log-normaliser (fitting.py:667-673)."""
    L = Ab.shape[2] // Ab.shape[1]
    ll = ar_log_likelihood(x, Ab, Q) * mask[:, L:, None]
    return hmm_filter(pi, ll)[0].sum()


def stateseq_marginals(x, mask, Ab, Q, pi):
    """Smoothed marginals (N, T-L, K) (fitting.py:536-538)."""
    L = Ab.shape[2] // Ab.shape[1]
    ll = ar_log_likelihood(x, Ab, Q) * mask[:, L:, None]
    _, filt = hmm_filter(pi, ll)
    N, T, K = ll.shape
    sm = np.empty_like(filt)
    sm[:, -1] = filt[:, -1]
    for t in range(T - 2, -1, -1):
        pred = filt[:, t] @ pi
        ratio = np.where(pred > 0, sm[:, t + 1] / np.where(pred > 0, pred, 1), 0.0)
        sm[:, t] = filt[:, t] * (ratio @ pi.T)
        sm[:, t] /= sm[:, t].sum(-1, keepdims=True)
    return sm


# ----------------------------------------------------------------------------
# Kalman FFBS (jax_moseq/utils/kalman.py: ar_to_lds, kalman_filter, kalman_sample)
# ----------------------------------------------------------------------------
def ar_to_lds(Ab, Q, jitter):
    """Companion form. Returns A (K,n,n), b (K,n), Qa (K,n,n)."""
    K, d = Ab.shape[0], Ab.shape[1]
    n = Ab.shape[2] - 1
    A = np.zeros((K, n, n))
    A[:, :n - d, d:] = np.eye(n - d)
    A[:, n - d:, :] = Ab[:, :, :-1]
    b = np.zeros((K, n))
    b[:, n - d:] = Ab[:, :, -1]
    Qa = np.zeros((K, n, n))
    Qa[:, :n - d, :n - d] = np.eye(n - d) * EPS_SHIFT
    Qa[:, n - d:, n - d:] = Q
    Qa = Qa + np.eye(n) * jitter
    return A, b, Qa


def _sym(S):
    return 0.5 * (S + np.swapaxes(S, -1, -2))


def kalman_filter(ys, omask, zs, m0, S0, A, B, Q, C, D, Rs):
    """Information-form condition + predict, batched over chains.

    ys (N,T,m), omask (N,T), zs (N,T-1), Rs (N,T,m) diagonal obs variances.
    Returns filtered means (N,T,n) and covariances (N,T,n,n).
    """
    N, T, _ = ys.shape
    n = m0.shape[0]
    m_pred = np.tile(m0, (N, 1))
    S_pred = np.tile(S0, (N, 1, 1))
    fm = np.empty((N, T, n))
    fS = np.empty((N, T, n, n))
    for t in range(T):
        on = omask[:, t] > 0
        Sinv = np.linalg.inv(S_pred)
        CtRi = C.T[None] / Rs[:, t][:, None, :]                       # (N,n,m)
        S_c = _sym(np.linalg.inv(Sinv + CtRi @ C))
        m_c = (S_c @ ((Sinv @ m_pred[..., None])[..., 0]
                      + (CtRi @ (ys[:, t] - D)[..., None])[..., 0])[..., None])[..., 0]
        m_c = np.where(on[:, None], m_c, m_pred)
        S_c = np.where(on[:, None, None], S_c, S_pred)
        fm[:, t], fS[:, t] = m_c, S_c
        if t < T - 1:
            Az, Bz, Qz = A[zs[:, t]], B[zs[:, t]], Q[zs[:, t]]
            m_n = (Az @ m_c[..., None])[..., 0] + Bz
            S_n = _sym(Az @ S_c @ np.swapaxes(Az, -1, -2) + Qz)
            m_pred = np.where(on[:, None], m_n, m_c)
            S_pred = np.where(on[:, None, None], S_n, S_c)
    return fm, fS


def kalman_sample(ys, omask, zs, m0, S0, A, B, Q, C, D, Rs, w):
    """FFBS draw; w (N,T,n) standard normals. Returns xs (N,T,n)."""
    fm, fS = kalman_filter(ys, omask, zs, m0, S0, A, B, Q, C, D, Rs)
    N, T, n = fm.shape
    Qinv = np.linalg.inv(Q)
    xs = np.empty((N, T, n))
    x = fm[:, -1] + (np.linalg.cholesky(fS[:, -1]) @ w[:, -1][..., None])[..., 0]
    xs[:, -1] = x
    for t in range(T - 2, -1, -1):
        on = omask[:, t] > 0
        Az, Bz, Qi = A[zs[:, t]], B[zs[:, t]], Qinv[zs[:, t]]
        AtQi = np.swapaxes(Az, -1, -2) @ Qi
        Sinv = np.linalg.inv(fS[:, t])
        S_c = _sym(np.linalg.inv(Sinv + AtQi @ Az))
        m_c = (S_c @ ((Sinv @ fm[:, t][..., None])[..., 0]
                      + (AtQi @ (x - Bz)[..., None])[..., 0])[..., None])[..., 0]
        x_new = m_c + (np.linalg.cholesky(S_c) @ w[:, t][..., None])[..., 0]
        x = np.where(on[:, None], x_new, x)
        xs[:, t] = x
    return xs


def resample_continuous_stateseqs(Y, mask, v, h, s, z, Cd, sigmasq, Ab, Q, jitter, w_x):
    """x | rest. w_x (N, T-L+1, n). Returns x (N,T,d)."""
    N, T, k, D = Y.shape
    d = Ab.shape[1]
    n = Ab.shape[2] - 1
    L = n // d
    Ct = lifted_obs_matrix(Cd, k, D)
    C = np.zeros((k * D, n))
    C[:, n - d:] = Ct[:, :-1]
    Dv = Ct[:, -1]
    ys = rotate(Y - v[:, :, None, :], -h).reshape(N, T, k * D)[:, L - 1:]
    Rs = np.repeat(s * sigmasq, D, axis=-1)[:, L - 1:]
    A, B, Qa = ar_to_lds(Ab, Q, jitter)
    xi = kalman_sample(ys, mask[:, L - 1:], z, np.zeros(n), X_PRIOR_VAR * np.eye(n),
                       A, B, Qa, C, Dv, Rs, w_x)
    head = xi[:, 0, :n - d].reshape(N, L - 1, d)
    return np.concatenate([head, xi[:, :, n - d:]], axis=1)


# ----------------------------------------------------------------------------
# bounded-attempt gamma / von Mises samplers driven by tapes
# ----------------------------------------------------------------------------
def gamma_mt(a, tape):
    """Gamma(a, 1) by Marsaglia-Tsang. tape (..., GAMMA_TAPE)."""
    a = np.asarray(a, dtype=np.float64)
    boost = a < 1.0
    a1 = np.where(boost, a + 1.0, a)
    dd = a1 - 1.0 / 3.0
    c = 1.0 / np.sqrt(9.0 * dd)
    out = dd.copy()
    done = np.zeros(a.shape, dtype=bool)
    for r in range(GAMMA_R):
        x = tape[..., r]
        u = tape[..., GAMMA_R + r]
        vv = 1.0 + c * x
        v3 = vv * vv * vv
        ok = vv > 0
        with np.errstate(divide="ignore", invalid="ignore"):
            acc = ok & (np.log(u) < 0.5 * x * x + dd - dd * v3 + dd * np.log(np.where(ok, v3, 1.0)))
        take = acc & ~done
        out = np.where(take, dd * v3, out)
        done |= acc
    with np.errstate(divide="ignore"):
        bf = np.where(boost, tape[..., 2 * GAMMA_R] ** (1.0 / np.where(boost, a, 1.0)), 1.0)
    return out * bf


def vonmises_bf(mu, kappa, tape):
    """von Mises(mu, kappa) by Best-Fisher. tape (..., VM_R, 3) uniforms."""
    kappa = np.asarray(kappa, dtype=np.float64)
    small = kappa < 1e-8
    kk = np.where(small, 1.0, kappa)
    tau = 1.0 + np.sqrt(1.0 + 4.0 * kk * kk)
    rho = (tau - np.sqrt(2.0 * tau)) / (2.0 * kk)
    r = (1.0 + rho * rho) / (2.0 * rho)
    dev = np.zeros(kappa.shape)
    done = np.zeros(kappa.shape, dtype=bool)
    for a in range(VM_R):
        u1, u2, u3 = tape[..., a, 0], tape[..., a, 1], tape[..., a, 2]
        zc = np.cos(np.pi * u1)
        f = (1.0 + r * zc) / (r + zc)
        c = kk * (r - f)
        with np.errstate(divide="ignore", invalid="ignore"):
            acc = (c * (2.0 - c) - u2 > 0) | (np.log(c / u2) + 1.0 - c >= 0)
        th = np.where(u3 > 0.5, 1.0, -1.0) * np.arccos(np.clip(f, -1.0, 1.0))
        take = acc & ~done
        dev = np.where(take, th, dev)
        done |= acc
    dev = np.where(small, np.pi * (2.0 * tape[..., 0, 0] - 1.0), dev)
    out = mu + dev
    return out - 2.0 * np.pi * np.floor((out + np.pi) / (2.0 * np.pi))


# ----------------------------------------------------------------------------
# per-frame / per-keypoint resamplers (jax_moseq/models/keypoint_slds/gibbs.py)
# ----------------------------------------------------------------------------
def compute_squared_error(Y, x, v, h, Cd):
    N, T, k, D = Y.shape
    Yhat = estimate_coordinates(x, v, h, Cd, k, D)
    return ((Y - Yhat) ** 2).sum(-1)


def resample_scales(Y, x, v, h, Cd, sigmasq, nu_s, s_0, g_s):
    """s | rest. g_s (N,T,k,GAMMA_TAPE). s = (e2/sig2 + nu_s s0) / (2 Gamma((nu_s+D)/2))."""
    D = Y.shape[-1] if SCALE_DOF is None else SCALE_DOF
    sqerr = compute_squared_error(Y, x, v, h, Cd)
    degs = nu_s + D
    variance = sqerr / sigmasq + s_0 * nu_s
    return variance / (2.0 * gamma_mt(np.full(variance.shape, degs / 2.0), g_s))


def obs_variance_suffstats(Y, mask, x, v, h, s, Cd):
    sqerr = compute_squared_error(Y, x, v, h, Cd)
    return (sqerr / s * mask[..., None]).sum((0, 1)), mask.sum()


def resample_obs_variance(Y, mask, x, v, h, s, Cd, nu_sigma, sigmasq_0, g_sig):
    """sigmasq | rest. g_sig (k, GAMMA_TAPE)."""
    D = Y.shape[-1] if OBSVAR_DOF is None else OBSVAR_DOF
    S, n_valid = obs_variance_suffstats(Y, mask, x, v, h, s, Cd)
    degs = nu_sigma + D * n_valid
    variance = S + nu_sigma * sigmasq_0
    return variance / (2.0 * gamma_mt(np.full(S.shape, degs / 2.0), g_sig))


def resample_heading(Y, v, x, s, Cd, sigmasq, u_h):
    """h | rest. u_h (N,T,VM_R,3)."""
    N, T, k, D = Y.shape
    Yb = _ybar(x, Cd, k, D)[..., :2]
    Yc = (Y - v[:, :, None, :])[..., :2]
    wgt = 1.0 / (s * sigmasq)
    kc = ((Yb[..., 0] * Yc[..., 0] + Yb[..., 1] * Yc[..., 1]) * wgt).sum(-1)
    ks = ((Yb[..., 0] * Yc[..., 1] - Yb[..., 1] * Yc[..., 0]) * wgt).sum(-1)
    return vonmises_bf(np.arctan2(ks, kc), np.sqrt(kc * kc + ks * ks), u_h)


def resample_location(Y, mask, x, h, s, Cd, sigmasq, sigmasq_loc, w_v):
    """v | rest. w_v (N,T,D). Scalar-variance random-walk FFBS per chain."""
    N, T, k, D = Y.shape
    Yrot = rotate(_ybar(x, Cd, k, D), h)
    prec = 1.0 / (s * sigmasq)
    gsq = 1.0 / prec.sum(-1)
    mu = ((Y - Yrot) * prec[..., None]).sum(-2) * gsq[..., None]
    fm = np.empty((N, T, D))
    fP = np.empty((N, T))
    m_pred = np.zeros((N, D))
    P_pred = np.full(N, V_PRIOR_VAR)
    for t in range(T):
        on = mask[:, t] > 0
        P_c = 1.0 / (1.0 / P_pred + 1.0 / gsq[:, t])
        m_c = P_c[:, None] * (m_pred / P_pred[:, None] + mu[:, t] / gsq[:, t][:, None])
        P_c = np.where(on, P_c, P_pred)
        m_c = np.where(on[:, None], m_c, m_pred)
        fm[:, t], fP[:, t] = m_c, P_c
        m_pred = m_c
        P_pred = np.where(on, P_c + sigmasq_loc, P_c)
    vs = np.empty((N, T, D))
    vcur = fm[:, -1] + np.sqrt(fP[:, -1])[:, None] * w_v[:, -1]
    vs[:, -1] = vcur
    for t in range(T - 2, -1, -1):
        on = mask[:, t] > 0
        Sg = 1.0 / (1.0 / fP[:, t] + 1.0 / sigmasq_loc)
        mc = Sg[:, None] * (fm[:, t] / fP[:, t][:, None] + vcur / sigmasq_loc)
        vnew = mc + np.sqrt(Sg)[:, None] * w_v[:, t]
        vcur = np.where(on[:, None], vnew, vcur)
        vs[:, t] = vcur
    return vs


# ----------------------------------------------------------------------------
# parameter updates (jax_moseq/models/arhmm/gibbs.py, utils/transitions.py,
# utils/distributions.py)
# ----------------------------------------------------------------------------
def ar_suffstats(x, z, mask, K):
    """Per-state Gram of f=[phi(n); 1; y(d)] over valid frames: (K, n+1+d, n+1+d)."""
    d = x.shape[-1]
    L = x.shape[1] - z.shape[1]
    phi = get_lags(x, L)
    f = np.concatenate([phi, np.ones(phi.shape[:-1] + (1,)), x[:, L:]], axis=-1)
    f = f.reshape(-1, f.shape[-1])
    zz = z.reshape(-1)
    mm = mask[:, L:].reshape(-1) > 0
    G = np.zeros((K, f.shape[-1], f.shape[-1]))
    for j in range(K):
        fj = f[(zz == j) & mm]
        G[j] = fj.T @ fj
    return G


def count_transitions(z, mask, K):
    """N_ij over consecutive pairs whose both frames are valid."""
    L = mask.shape[1] - z.shape[1]
    m = mask[:, L:] > 0
    ok = m[:, :-1] & m[:, 1:]
    a = z[:, :-1][ok]
    b = z[:, 1:][ok]
    Nij = np.zeros((K, K), dtype=np.int64)
    np.add.at(Nij, (a, b), 1)
    return Nij


def mniw_posterior(G, nu_0, S_0, M_0, K_0):
    """Posterior (M_n, K_n, S_n, nu_n) of one state's MNIW from its Gram of f = [phi; 1; y]."""
    p = K_0.shape[0]                       # n+1
    Sxx, Syx, Syy, cnt = G[:p, :p], G[p:, :p], G[p:, p:], G[p - 1, p - 1]
    K0i = np.linalg.inv(K_0)
    Kni = K0i + Sxx
    Kn = _sym(np.linalg.inv(Kni))
    Mn = (M_0 @ K0i + Syx) @ Kn
    Sn = _sym(S_0 + Syy + M_0 @ K0i @ M_0.T - Mn @ Kni @ Mn.T)
    return Mn, Kn, Sn, nu_0 + cnt


def sample_mniw_from_stats(G, nu_0, S_0, M_0, K_0, w_G, w_B, g_chi):
    """One state's (Ab, Q) from its Gram. w_G (d,n+1), w_B (d,d), g_chi (d,GAMMA_TAPE)."""
    d = S_0.shape[0]
    Mn, Kn, Sn, nu = mniw_posterior(G, nu_0, S_0, M_0, K_0)
    chi2 = 2.0 * gamma_mt((nu - np.arange(d)) / 2.0, g_chi)
    Z = np.diag(np.sqrt(chi2)) + np.tril(w_B, -1)
    Ls = np.linalg.cholesky(Sn)
    Tm = Ls @ np.linalg.inv(Z).T
    Qs = _sym(Tm @ Tm.T)
    Ab = Mn + np.linalg.cholesky(Qs) @ w_G @ np.linalg.cholesky(Kn).T
    return Ab, Qs


def resample_ar_params(x, z, mask, K, nu_0, S_0, M_0, K_0, w_G, w_B, g_chi):
    """(Ab, Q) | x, z. Tapes: w_G (K,d,n+1), w_B (K,d,d), g_chi (K,d,GAMMA_TAPE)."""
    G = ar_suffstats(x, z, mask, K)
    out = [sample_mniw_from_stats(G[j], nu_0, S_0, M_0, K_0, w_G[j], w_B[j], g_chi[j])
           for j in range(K)]
    return np.stack([o[0] for o in out]), np.stack([o[1] for o in out])


def resample_hdp_transitions(z, mask, betas, alpha, kappa, gamma, u_crp, u_bin, g_beta, g_pi,
                             counts=None):
    """(betas, pi) | z: weak-limit sticky HDP-HMM.

    u_crp, u_bin: flat uniform tapes (length >= number of counted transitions);
    g_beta (K,GAMMA_TAPE); g_pi (K,K,GAMMA_TAPE).
    """
    K = betas.shape[0]
    Nij = count_transitions(z, mask, K) if counts is None else counts
    flat = Nij.reshape(-1)
    starts = np.concatenate([[0], np.cumsum(flat)[:-1]]).reshape(K, K)
    conc = alpha * betas[None, :] + kappa * np.eye(K)
    m = np.zeros((K, K), dtype=np.int64)
    for i in range(K):
        for j in range(K):
            nn = Nij[i, j]
            if nn:
                r = np.arange(nn)
                m[i, j] = (u_crp[starts[i, j]:starts[i, j] + nn] < conc[i, j] / (r + conc[i, j])).sum()
    rho = kappa / (alpha + kappa)
    p_over = rho / (rho + betas * (1.0 - rho))
    dstart = np.concatenate([[0], np.cumsum(np.diag(Nij))[:-1]])
    w = np.array([(u_bin[dstart[i]:dstart[i] + m[i, i]] < p_over[i]).sum() for i in range(K)])
    mbar = m - np.diag(w)
    gb = gamma_mt(gamma / K + mbar.sum(0), g_beta)
    new_betas = gb / gb.sum()
    gp = gamma_mt(alpha * new_betas[None, :] + kappa * np.eye(K) + Nij, g_pi)
    pi = gp / gp.sum(1, keepdims=True)
    return new_betas, pi


# ----------------------------------------------------------------------------
# whole sweep (jax_moseq/models/keypoint_slds/gibbs.py resample_model, reached
# from /root/reference/keypoint_moseq/fitting.py:25)
# ----------------------------------------------------------------------------
def make_tape(rng, N, T, k, D, d, L, K, n_trans=None):
    """All draws one sweep can consume, as float64 arrays."""
    n = d * L
    tot = N * (T - L) if n_trans is None else n_trans

    def gam(*shape):
        t = np.empty(shape + (GAMMA_TAPE,))
        t[..., :GAMMA_R] = rng.standard_normal(shape + (GAMMA_R,))
        t[..., GAMMA_R:] = rng.uniform(1e-12, 1.0, shape + (GAMMA_R + 1,))
        return t

    return {
        "u_z": rng.uniform(1e-12, 1.0, (N, T - L)),
        "w_x": rng.standard_normal((N, T - L + 1, n)),
        "g_s": gam(N, T, k),
        "u_h": rng.uniform(1e-12, 1.0, (N, T, VM_R, 3)),
        "w_v": rng.standard_normal((N, T, D)),
        "w_G": rng.standard_normal((K, d, n + 1)),
        "w_B": rng.standard_normal((K, d, d)),
        "g_chi": gam(K, d),
        "u_crp": rng.uniform(1e-12, 1.0, tot),
        "u_bin": rng.uniform(1e-12, 1.0, tot),
        "g_beta": gam(K),
        "g_pi": gam(K, K),
        "g_sig": gam(k),
    }


def resample_model(data, states, params, hypparams, noise_prior, tape, ar_only=False,
                   states_only=False, resample_global_noise_scale=False,
                   resample_local_noise_scale=True, fix_heading=False, jitter=1e-3):
    """One Gibbs sweep in float64 with taped draws. Returns (states, params, logZ)."""
    Y, mask = data["Y"], data["mask"]
    st = dict(states)
    pr = dict(params)
    th, ah = hypparams["trans_hypparams"], hypparams["ar_hypparams"]
    oh, ch = hypparams["obs_hypparams"], hypparams["cen_hypparams"]
    K = int(th["num_states"])
    if not states_only:
        pr["betas"], pr["pi"] = resample_hdp_transitions(
            st["z"], mask, pr["betas"], th["alpha"], th["kappa"], th["gamma"],
            tape["u_crp"], tape["u_bin"], tape["g_beta"], tape["g_pi"])
        pr["Ab"], pr["Q"] = resample_ar_params(
            st["x"], st["z"], mask, K, ah["nu_0"], ah["S_0"], ah["M_0"], ah["K_0"],
            tape["w_G"], tape["w_B"], tape["g_chi"])
    st["z"], logZ = resample_discrete_stateseqs(st["x"], mask, pr["Ab"], pr["Q"], pr["pi"],
                                                tape["u_z"])
    if ar_only:
        return st, pr, logZ
    if not states_only and resample_global_noise_scale:
        pr["sigmasq"] = resample_obs_variance(Y, mask, st["x"], st["v"], st["h"], st["s"],
                                              pr["Cd"], oh["nu_sigma"], oh["sigmasq_0"],
                                              tape["g_sig"])
    if resample_local_noise_scale:
        st["s"] = resample_scales(Y, st["x"], st["v"], st["h"], pr["Cd"], pr["sigmasq"],
                                  oh["nu_s"], noise_prior, tape["g_s"])
    st["x"] = resample_continuous_stateseqs(Y, mask, st["v"], st["h"], st["s"], st["z"],
                                            pr["Cd"], pr["sigmasq"], pr["Ab"], pr["Q"],
                                            jitter, tape["w_x"])
    if not fix_heading:
        st["h"] = resample_heading(Y, st["v"], st["x"], st["s"], pr["Cd"], pr["sigmasq"],
                                   tape["u_h"])
    st["v"] = resample_location(Y, mask, st["x"], st["h"], st["s"], pr["Cd"], pr["sigmasq"],
                                ch["sigmasq_loc"], tape["w_v"])
    return st, pr, logZ
