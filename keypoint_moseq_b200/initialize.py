"""Model initialisation from raw keypoints: `fit_pca` and the parameter / hyper-parameter part of
`init_model`.

The reference re-exports `fit_pca` from jax_moseq (/root/reference/keypoint_moseq/__init__.py:21) and
forwards `init_model` to `jax_moseq.models.keypoint_slds.init_model`
(/root/reference/keypoint_moseq/fitting.py:63-106); the notebook workflow is
`pca = kpms.fit_pca(**data, **config())`, `model = kpms.init_model(data, pca=pca, **config())`.
jax_moseq is not vendored, so the conventions below are restated from the in-tree evidence:
`pca.mean_` has (k-1)*D entries and `center_embedding(k)` maps them back to k keypoints
(/root/reference/keypoint_moseq/viz.py:189-204), `Cd` is ((k-1)*D, latent_dim + 1) with the offset in the
last column (docs/source/advanced.rst:21-26), hyper-parameter names and defaults come from
`generate_config` (/root/reference/keypoint_moseq/io.py:62-86).  Host-side NumPy / scikit-learn: this
runs once per fit, upstream of the Gibbs sweep.
"""
import numpy as np

from .synth import center_embedding

__all__ = ["align_egocentric", "preprocess_for_pca", "fit_pca", "init_hyperparams", "init_params",
           "noise_prior_from_confidence", "default_config"]


def default_config(**overrides):
    """The modelling entries of the reference's `config.yml` with the defaults of `generate_config`
    (/root/reference/keypoint_moseq/io.py:62-133): the four hyper-parameter groups, the error estimator and
    the flags `format_data` / `fit_pca` / `init_model` read.  Use as `init_model(data, pca=pca,
    **default_config())`; like `update_config`, an override matches a key at any nesting level."""
    cfg = {
        "error_estimator": {"slope": -0.5, "intercept": 0.25},
        "obs_hypparams": {"sigmasq_0": 0.1, "sigmasq_C": 0.1, "nu_sigma": 1e5, "nu_s": 5},
        "ar_hypparams": {"latent_dim": 10, "nlags": 3, "S_0_scale": 0.01, "K_0_scale": 10.0},
        "trans_hypparams": {"num_states": 100, "gamma": 1e3, "alpha": 5.7, "kappa": 1e6},
        "cen_hypparams": {"sigmasq_loc": 0.5},
        "conf_pseudocount": 1e-3, "whiten": True, "fix_heading": False,
        "added_noise_level": 0.1, "PCA_fitting_num_frames": 1000000, "conf_threshold": 0.5,
    }
    for key, val in overrides.items():
        hit = False
        if key in cfg:
            cfg[key], hit = val, True
        for group in cfg.values():
            if isinstance(group, dict) and key in group:
                group[key], hit = val, True
        if not hit:
            cfg[key] = val
    return cfg


def _np(a):
    try:
        import torch
        if isinstance(a, torch.Tensor):
            return a.detach().cpu().numpy()
    except Exception:  # pragma: no cover
        pass
    return np.asarray(a)


def align_egocentric(Y, anterior_idxs=None, posterior_idxs=None, fix_heading=False):
    """Centroid, heading and egocentrically aligned keypoints.

    v = mean over keypoints; h = angle (xy-plane) of the posterior->anterior axis, 0 when the index
    lists are missing or `fix_heading`; Y_aligned = (Y - v) rotated by -h (inverse_rigid_transform,
    /root/reference/keypoint_moseq/util.py:498-508)."""
    Y = _np(Y).astype(np.float64)
    v = Y.mean(-2)
    if fix_heading or anterior_idxs is None or posterior_idxs is None or not len(anterior_idxs) or not len(posterior_idxs):
        h = np.zeros(Y.shape[:-2])
    else:
        ant = Y[..., list(anterior_idxs), :].mean(-2)
        pos = Y[..., list(posterior_idxs), :].mean(-2)
        h = np.arctan2(ant[..., 1] - pos[..., 1], ant[..., 0] - pos[..., 0])
    Yc = Y - v[..., None, :]
    c, s = np.cos(h)[..., None], np.sin(h)[..., None]
    Ya = Yc.copy()
    Ya[..., 0] = c * Yc[..., 0] + s * Yc[..., 1]
    Ya[..., 1] = -s * Yc[..., 0] + c * Yc[..., 1]
    return Ya, v, h


def preprocess_for_pca(Y, anterior_idxs=None, posterior_idxs=None, fix_heading=False):
    """(..., k, D) keypoints -> (..., (k-1)*D) aligned, centred coordinates in the zero-mean subspace."""
    Ya, v, h = align_egocentric(Y, anterior_idxs, posterior_idxs, fix_heading)
    k, D = Ya.shape[-2:]
    Gamma = center_embedding(k)                                   # (k, k-1), Gamma' Gamma = I
    flat = np.einsum("kj,...kc->...jc", Gamma, Ya).reshape(*Ya.shape[:-2], (k - 1) * D)
    return flat, v, h


def fit_pca(Y, mask, conf=None, anterior_idxs=None, posterior_idxs=None, conf_threshold=0.5, verbose=False,
            PCA_fitting_num_frames=1000000, fix_heading=False, seed=0, **kwargs):
    """PCA of the aligned, centred keypoints (`kpms.fit_pca(**data, **config())`).

    Frames enter when `mask` is set and, if `conf` is given, every keypoint's confidence exceeds
    `conf_threshold`; at most `PCA_fitting_num_frames` of them (random subset).  Returns a fitted
    `sklearn.decomposition.PCA` with (k-1)*D features."""
    from sklearn.decomposition import PCA
    flat, _, _ = preprocess_for_pca(Y, anterior_idxs, posterior_idxs, fix_heading)
    keep = _np(mask) > 0
    if conf is not None:
        keep = keep & (_np(conf) > conf_threshold).all(-1)
    rows = flat[keep]
    if rows.shape[0] == 0:
        raise ValueError("fit_pca: no frame passes the mask / confidence filter")
    n = int(min(PCA_fitting_num_frames, rows.shape[0]))
    if n < rows.shape[0]:
        rows = rows[np.random.default_rng(seed).choice(rows.shape[0], n, replace=False)]
    if verbose:
        print(f"PCA: fitting on {rows.shape[0]} frames")
    return PCA(n_components=min(rows.shape)).fit(rows)


def init_hyperparams(trans_hypparams, ar_hypparams, obs_hypparams, cen_hypparams, **kwargs):
    """Fill the derived hyper-parameters (layout of a reference checkpoint):
    S_0 = S_0_scale I, K_0 = K_0_scale I, M_0 = identity on the newest lag, nu_0 = latent_dim + 2."""
    ar = dict(ar_hypparams)
    d, L = int(ar["latent_dim"]), int(ar["nlags"])
    K = int(trans_hypparams["num_states"])
    n = d * L
    M_0 = np.zeros((d, n + 1))
    M_0[:, n - d:n] = np.eye(d)
    ar.setdefault("S_0", float(ar["S_0_scale"]) * np.eye(d))
    ar.setdefault("K_0", float(ar["K_0_scale"]) * np.eye(n + 1))
    ar.setdefault("M_0", M_0)
    ar.setdefault("nu_0", d + 2)
    ar["num_states"] = K
    return {"trans_hypparams": dict(trans_hypparams, num_states=K), "ar_hypparams": ar,
            "obs_hypparams": dict(obs_hypparams), "cen_hypparams": dict(cen_hypparams)}


def init_params(pca, hypparams, k, flat=None, whiten=True, seed=0, **kwargs):
    """Initial parameters: `Cd` from the PCA (whitened latents when `whiten`), `sigmasq` = 1, and prior
    draws for betas / pi (weak-limit sticky HDP) and Ab / Q (MNIW).

    `flat` = the preprocessed coordinates the latents will be read from (masked rows), used for the
    whitening transform; without it the PCA's own explained variances are used."""
    from scipy.stats import invwishart
    th, ar = hypparams["trans_hypparams"], hypparams["ar_hypparams"]
    d, K = int(ar["latent_dim"]), int(th["num_states"])
    if d > pca.components_.shape[0]:
        raise ValueError(f"latent_dim {d} exceeds the {pca.components_.shape[0]} fitted principal components")
    rng = np.random.default_rng(seed)
    comps = pca.components_[:d]                                   # (d, (k-1)D), orthonormal rows
    if whiten:
        if flat is not None and flat.shape[0] > d:
            lat = (flat - pca.mean_) @ comps.T
            cov = np.cov(lat, rowvar=False).reshape(d, d)
        else:
            cov = np.diag(pca.explained_variance_[:d])
        Wc = np.linalg.cholesky(cov + 1e-12 * np.eye(d))
        Cmat = comps.T @ Wc                                       # y = C x + mean with cov(x) = I
    else:
        Cmat = comps.T
    Cd = np.concatenate([Cmat, pca.mean_[:, None]], axis=1)
    betas = rng.dirichlet(np.full(K, th["gamma"] / K))
    pi = np.stack([rng.dirichlet(th["alpha"] * betas + th["kappa"] * np.eye(K)[i] + 1e-8) for i in range(K)])
    S_0, K_0, M_0, nu_0 = (np.asarray(ar[key], dtype=np.float64) if key != "nu_0" else float(ar[key])
                           for key in ("S_0", "K_0", "M_0", "nu_0"))
    LK = np.linalg.cholesky(K_0)
    Ab, Q = [], []
    for _ in range(K):
        q = np.atleast_2d(invwishart.rvs(df=nu_0, scale=S_0, random_state=rng))
        g = rng.standard_normal(M_0.shape)
        Ab.append(M_0 + np.linalg.cholesky(q) @ g @ LK.T)
        Q.append(q)
    return {"Ab": np.stack(Ab), "Q": np.stack(Q), "betas": betas, "pi": pi, "Cd": Cd, "sigmasq": np.ones(k)}


def noise_prior_from_confidence(conf, error_estimator=None):
    """noise_prior = (10 ** (slope * log10(conf) + intercept)) ** 2 (error_estimator of the config,
    /root/reference/keypoint_moseq/io.py:62-66); ones without an estimator."""
    conf = _np(conf).astype(np.float64)
    if error_estimator is None:
        return np.ones_like(conf)
    slope, intercept = float(error_estimator["slope"]), float(error_estimator["intercept"])
    return (10.0 ** (np.log10(np.maximum(conf, 1e-6)) * slope + intercept)) ** 2
