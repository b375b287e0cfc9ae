"""Synthetic keypoint data sampled from the keypoint-SLDS generative process.

Used by the parity tests and bench.py (there is no network for real datasets).
The process and the default hyper-parameters follow the reference's config
defaults (/root/reference/keypoint_moseq/io.py:62-86) and the batch layout of
`format_data` (/root/reference/keypoint_moseq/util.py:1071-1088).
"""
import numpy as np

from .util import batch

__all__ = ["center_embedding", "default_hypparams", "sample_dataset", "CONFIGS"]

# BASELINE.json configs as (recordings, frames, keypoints k, D, latent d, nlags L, states K)
CONFIGS = {
    "C1": dict(recordings=4, frames=10_000, k=10, D=2, d=4, L=3, K=100),
    "C2": dict(recordings=20, frames=36_000, k=12, D=2, d=10, L=3, K=100),
    "C3": dict(recordings=50, frames=108_000, k=16, D=3, d=10, L=3, K=100),
    "C4": dict(recordings=200, frames=54_000, k=12, D=2, d=10, L=3, K=100),
}


def center_embedding(k):
    """(k, k-1) orthonormal basis of zero-mean k-tuples (as upstream: SVD of the
    centring matrix; reference use at /root/reference/keypoint_moseq/viz.py:197)."""
    return np.linalg.svd(np.eye(k) - np.ones((k, k)) / k)[0][:, :-1]


def default_hypparams(d, L, K, kappa=1e4):
    """hypparams dict with the layout of a reference checkpoint."""
    n = d * L
    M_0 = np.zeros((d, n + 1))
    M_0[:, n - d:n] = np.eye(d)
    return {
        "trans_hypparams": {"num_states": int(K), "gamma": 1e3, "alpha": 5.7, "kappa": float(kappa)},
        "ar_hypparams": {"latent_dim": int(d), "nlags": int(L), "S_0_scale": 0.01, "K_0_scale": 10.0,
                         "S_0": 0.01 * np.eye(d), "K_0": 10.0 * np.eye(n + 1), "M_0": M_0,
                         "nu_0": int(d + 2), "num_states": int(K)},
        "obs_hypparams": {"sigmasq_0": 0.1, "sigmasq_C": 0.1, "nu_sigma": 1e5, "nu_s": 5},
        "cen_hypparams": {"sigmasq_loc": 0.5},
    }


def _stable_ar(rng, d, L):
    n = d * L
    A = np.zeros((d, n))
    A[:, n - d:] = 0.85 * np.eye(d)
    A += rng.standard_normal((d, n)) * 0.08
    comp = np.zeros((n, n))
    comp[:n - d, d:] = np.eye(n - d)
    for _ in range(50):
        comp[n - d:] = A
        rad = np.abs(np.linalg.eigvals(comp)).max()
        if rad < 0.97:
            break
        A *= 0.95 / rad
    b = rng.standard_normal(d) * 0.05
    R = rng.standard_normal((d, d))
    Q = 0.04 * (R @ R.T / d + 0.5 * np.eye(d))
    return np.concatenate([A, b[:, None]], axis=1), Q


def sample_dataset(recordings=4, frames=10_000, k=10, D=2, d=4, L=3, K=100, seed=0, kappa=1e4,
                   kappa_gen=100.0, seg_length=None, max_seg_length=10_000, dtype=np.float64, data_seed=None):
    """Draw params, states and keypoints; batch them like `format_data`.

    Returns (data, metadata, model) with NumPy leaves:
      data  = {"Y" (N,T,k,D), "conf" (N,T,k), "mask" (N,T)}
      model = {"seed", "states", "params", "hypparams", "noise_prior"}
    `seed` draws the parameters; `data_seed` (default: continue the same stream) draws the states and
    keypoints, so that several shards of one cohort share one generative model.
    """
    rng = np.random.default_rng(seed)
    hyp = default_hypparams(d, L, K, kappa)
    th = hyp["trans_hypparams"]
    betas = rng.dirichlet(np.full(K, th["gamma"] / K))
    pi = np.stack([rng.dirichlet(th["alpha"] * betas + kappa_gen * np.eye(K)[i] + 1e-3)
                   for i in range(K)])
    AbQ = [_stable_ar(rng, d, L) for _ in range(K)]
    Ab = np.stack([a for a, _ in AbQ])
    Q = np.stack([q for _, q in AbQ])
    Lq = np.linalg.cholesky(Q)
    Cmat = np.linalg.qr(rng.standard_normal(((k - 1) * D, d)))[0] * 3.0
    d0 = rng.standard_normal((k - 1) * D) * 5.0
    Cd = np.concatenate([Cmat, d0[:, None]], axis=1)
    sigmasq = 0.1 * rng.uniform(0.5, 2.0, k)
    Gamma = center_embedding(k)
    nu_s = hyp["obs_hypparams"]["nu_s"]
    sig_loc = hyp["cen_hypparams"]["sigmasq_loc"]
    cum = np.cumsum(pi, axis=1)
    if data_seed is not None:
        rng = np.random.default_rng(data_seed)

    R_, T = recordings, frames
    u = rng.random((R_, T))
    z = np.empty((R_, T), dtype=np.int64)
    z[:, 0] = rng.integers(K, size=R_)
    for t in range(1, T):
        z[:, t] = np.minimum((cum[z[:, t - 1]] < u[:, t, None]).sum(1), K - 1)
    x = np.zeros((R_, T, d))
    noise = rng.standard_normal((R_, T, d))
    x[:, :L] = noise[:, :L] * 0.3
    for t in range(L, T):
        a = Ab[z[:, t]]
        x[:, t] = (np.einsum("rij,rj->ri", a[:, :, :-1], x[:, t - L:t].reshape(R_, -1)) + a[:, :, -1]
                   + np.einsum("rij,rj->ri", Lq[z[:, t]], noise[:, t]))
    h = np.cumsum(rng.standard_normal((R_, T)) * 0.1, axis=1)
    h = h - 2 * np.pi * np.floor((h + np.pi) / (2 * np.pi))
    v = np.cumsum(rng.standard_normal((R_, T, D)) * np.sqrt(sig_loc), axis=1)
    conf = rng.beta(5.0, 1.0, (R_, T, k))
    prior = (10.0 ** (np.log10(conf + 1e-6) * -0.5 + 0.25)) ** 2
    s = nu_s * prior / rng.chisquare(nu_s, (R_, T, k))
    Ybar = np.einsum("kj,rtjc->rtkc", Gamma, (x @ Cmat.T + d0).reshape(R_, T, k - 1, D))
    c, sn = np.cos(h)[..., None], np.sin(h)[..., None]
    Yr = Ybar.copy()
    Yr[..., 0] = c * Ybar[..., 0] - sn * Ybar[..., 1]
    Yr[..., 1] = sn * Ybar[..., 0] + c * Ybar[..., 1]
    Y = Yr + v[:, :, None, :] + rng.standard_normal((R_, T, k, D)) * np.sqrt(s * sigmasq)[..., None]
    coords, confs, st = {}, {}, {n: {} for n in ("x", "v", "h", "s", "z", "prior")}
    for r in range(R_):
        key = f"rec{r:04d}"
        coords[key], confs[key] = Y[r], conf[r]
        for name, val in (("x", x), ("v", v), ("h", h), ("s", s), ("z", z), ("prior", prior)):
            st[name][key] = val[r]

    keys = sorted(coords.keys())
    if seg_length is None:
        seg_length = min(frames, max_seg_length)
    Yb, mask, metadata = batch(coords, seg_length=seg_length, keys=keys)
    Yb = Yb.astype(np.float64)
    Yb += np.random.default_rng(42).uniform(-0.1, 0.1, Yb.shape)
    conf_b = batch(confs, seg_length=seg_length, keys=keys)[0] + 1e-3
    bt = {n: batch(st[n], seg_length=seg_length, keys=keys)[0] for n in st}
    data = {"Y": Yb.astype(dtype), "conf": conf_b.astype(dtype), "mask": mask}
    model = {
        "seed": np.array([0, seed], dtype=np.uint32),
        "states": {"x": bt["x"].astype(dtype), "v": bt["v"].astype(dtype), "h": bt["h"].astype(dtype),
                   "s": bt["s"].astype(dtype), "z": bt["z"][:, L:].astype(np.int64)},
        "params": {"Ab": Ab.astype(dtype), "Q": Q.astype(dtype), "betas": betas.astype(dtype),
                   "pi": pi.astype(dtype), "Cd": Cd.astype(dtype), "sigmasq": sigmasq.astype(dtype)},
        "hypparams": hyp,
        "noise_prior": bt["prior"].astype(dtype),
    }
    return data, metadata, model
