"""fit_model / apply_model through the reference's signatures on the GPU, checkpoint layout and resume,
NaN guard, full-size properties and the 2-GPU sharded sweep."""
import os
import subprocess
import sys
import warnings

import numpy as np
import pytest
import torch

from helpers import small_problem

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fit_model_checkpoints_and_resume(tmp_path):
    from keypoint_moseq_b200 import fitting, io as kio
    data, meta, model = small_problem(seed=11, d=4, L=3, K=12, k=5, D=2, kappa=1e2)
    proj = str(tmp_path)
    out, name = fitting.fit_model(model, data, meta, proj, "m1", num_iters=4, save_every_n_iters=2, ar_only=True)
    assert set(out) == {"seed", "states", "params", "hypparams", "noise_prior"}
    ck = kio.load_hdf5(os.path.join(proj, "m1", "checkpoint.h5"))
    assert sorted(ck["model_snapshots"], key=int) == ["0", "2", "4"]
    assert set(ck["data"]) == {"Y", "conf", "mask"} and isinstance(ck["metadata"], tuple)
    snap = ck["model_snapshots"]["4"]
    assert snap["states"]["z"].shape == (data["Y"].shape[0], data["Y"].shape[1] - 3)
    assert snap["params"]["Ab"].shape == (12, 4, 13) and snap["seed"].dtype == np.uint32
    np.testing.assert_array_equal(snap["states"]["z"], out["states"]["z"].cpu().numpy())
    # ar_only leaves the continuous states untouched
    np.testing.assert_array_equal(snap["states"]["x"], model["states"]["x"])
    # resume from iteration 2 with the full model: later snapshots are dropped, new ones written
    m2, d2, md2, it = kio.load_checkpoint(proj, "m1", iteration=2)
    out2, _ = fitting.fit_model(m2, d2, md2, proj, "m1", start_iter=2, num_iters=5, save_every_n_iters=-1)
    ck2 = kio.load_hdf5(os.path.join(proj, "m1", "checkpoint.h5"))
    assert sorted(ck2["model_snapshots"], key=int) == ["0", "2", "5"]
    assert torch.isfinite(out2["states"]["x"]).all()
    assert not np.array_equal(ck2["model_snapshots"]["5"]["states"]["x"], model["states"]["x"])


def test_async_checkpoints_write_the_same_snapshots(tmp_path):
    """async_checkpoints=True copies each snapshot to pinned memory on a side stream while the next sweeps are
    already queued (util.AsyncHostCopy) and writes it on a background thread: every snapshot must equal the one
    the synchronous path writes (the reference's blocking save, keypoint_moseq/fitting.py:266-275), i.e. later
    sweeps must not leak into an earlier snapshot."""
    from keypoint_moseq_b200 import fitting, io as kio
    data, meta, model = small_problem(seed=14, d=4, L=3, K=12, k=5, D=2, kappa=1e2, frames=1200, seg_length=600)
    proj = str(tmp_path)
    for name, flag in (("sync", False), ("async", True)):
        fitting.fit_model(model, data, meta, proj, name, num_iters=7, save_every_n_iters=1, dtype=torch.float32,
                          async_checkpoints=flag)
    a = kio.load_hdf5(os.path.join(proj, "sync", "checkpoint.h5"))["model_snapshots"]
    b = kio.load_hdf5(os.path.join(proj, "async", "checkpoint.h5"))["model_snapshots"]
    assert sorted(a, key=int) == sorted(b, key=int) == [str(i) for i in range(8)]
    for it in a:
        for key in ("x", "v", "h", "s", "z"):
            np.testing.assert_array_equal(a[it]["states"][key], b[it]["states"][key], err_msg=f"{it}/{key}")
            assert a[it]["states"][key].dtype == b[it]["states"][key].dtype
        for key in ("Ab", "Q", "pi", "betas"):
            np.testing.assert_array_equal(a[it]["params"][key], b[it]["params"][key], err_msg=f"{it}/{key}")
        np.testing.assert_array_equal(a[it]["seed"], b[it]["seed"])
    assert not np.array_equal(a["3"]["states"]["x"], a["4"]["states"]["x"])


def test_fit_model_is_deterministic_and_float32_runs(tmp_path):
    from keypoint_moseq_b200 import fitting
    data, meta, model = small_problem(seed=12, d=4, L=3, K=12, k=5, D=2, kappa=1e2)
    a, _ = fitting.fit_model(model, data, meta, num_iters=3, save_every_n_iters=None, dtype=torch.float32)
    b, _ = fitting.fit_model(model, data, meta, num_iters=3, save_every_n_iters=None, dtype=torch.float32)
    for key in ("x", "v", "h", "s", "z"):
        assert torch.equal(a["states"][key], b["states"][key]), key
    assert a["states"]["x"].dtype == torch.float32 and a["params"]["Ab"].dtype == torch.float64


def test_init_model_from_raw_keypoints_then_fit(tmp_path):
    """The notebook workflow without JAX: fit_pca -> init_model(data, pca=pca, **config) -> AR-only
    sweeps -> full sweeps.  Starts from prior draws, so the verified time chunks run their fallbacks."""
    from keypoint_moseq_b200 import fitting, initialize
    data, meta, truth = small_problem(seed=21, d=4, L=3, K=12, k=6, D=2, kappa=1e2, frames=500, seg_length=250)
    config = {
        "trans_hypparams": {"num_states": 12, "gamma": 1e3, "alpha": 5.7, "kappa": 1e4},
        "ar_hypparams": {"latent_dim": 4, "nlags": 3, "S_0_scale": 0.01, "K_0_scale": 10.0},
        "obs_hypparams": {"sigmasq_0": 0.1, "sigmasq_C": 0.1, "nu_sigma": 1e5, "nu_s": 5},
        "cen_hypparams": {"sigmasq_loc": 0.5},
        "error_estimator": {"slope": -0.5, "intercept": 0.25},
        "anterior_idxs": [0, 1], "posterior_idxs": [4, 5], "whiten": True, "fix_heading": False,
    }
    pca = initialize.fit_pca(**data, **config, conf_threshold=0.0)
    model = fitting.init_model(data, pca=pca, **config, seed=np.array([0, 5], dtype=np.uint32))
    assert set(model) == {"seed", "states", "params", "hypparams", "noise_prior"}
    assert model["params"]["Cd"].shape == (10, 5) and model["hypparams"]["ar_hypparams"]["nu_0"] == 6
    assert tuple(model["states"]["z"].shape) == (data["Y"].shape[0], data["Y"].shape[1] - 3)
    m, _ = fitting.fit_model(model, data, meta, num_iters=5, ar_only=True, save_every_n_iters=None)
    m = fitting.update_hypparams(m, kappa=1e3)
    m, _ = fitting.fit_model(m, data, meta, num_iters=5, save_every_n_iters=None)
    for key in ("x", "v", "h", "s"):
        assert torch.isfinite(m["states"][key]).all(), key
    from keypoint_moseq_b200 import gibbs
    import oracle as orc
    Yhat = orc.estimate_coordinates(*(m["states"][q].cpu().numpy().astype(float) for q in ("x", "v", "h")),
                                    m["params"]["Cd"].cpu().numpy(), 6, 2)
    resid = (data["Y"] - Yhat)[data["mask"] > 0]
    assert np.sqrt((resid ** 2).mean()) < 2.0, "the fitted model stops reconstructing the keypoints"


def test_apply_model_and_results_layout(tmp_path):
    from keypoint_moseq_b200 import fitting, io as kio
    data, meta, model = small_problem(seed=13, d=4, L=3, K=12, k=5, D=2, kappa=1e2)
    proj = str(tmp_path)
    res, m = fitting.apply_model(model, data, meta, proj, "m", num_iters=3, return_model=True,
                                 anterior_idxs=[0], posterior_idxs=[4])
    assert sorted(res) == ["rec0000", "rec0001"]
    for rec in res.values():
        assert rec["syllable"].shape == (260,) and rec["latent_state"].shape == (260, 4)
        assert rec["centroid"].shape == (260, 2) and rec["heading"].shape == (260,)
        assert rec["syllable"].min() >= 0 and rec["syllable"].max() < 12
    # parameters are fixed in apply_model (states_only sweeps)
    np.testing.assert_array_equal(m["params"]["Ab"].cpu().numpy(), model["params"]["Ab"])
    saved = kio.load_results(proj, "m")
    np.testing.assert_array_equal(saved["rec0001"]["syllable"], res["rec0001"]["syllable"])
    with pytest.raises(RuntimeError):
        fitting.apply_model(model, data, meta, proj, "m", num_iters=1)
    fitting.apply_model(model, data, meta, proj, "m", num_iters=1, overwrite=True)
    marg = fitting.estimate_syllable_marginals(model, data, meta, burn_in_iters=1, num_samples=2, steps_per_sample=1)
    assert marg["rec0000"].shape == (260, 12)
    np.testing.assert_allclose(marg["rec0000"].sum(1), 1.0, atol=1e-9)


def test_nan_guard_stops_fitting(tmp_path):
    from keypoint_moseq_b200 import fitting
    data, meta, model = small_problem(seed=14, d=4, L=3, K=12, k=5, D=2, kappa=1e2)
    data = dict(data, Y=data["Y"].copy())
    data["Y"][1, 20, 2, 0] = np.nan
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        out, _ = fitting.fit_model(model, data, meta, num_iters=5, save_every_n_iters=None)
    assert any("NaNs encountered" in str(x.message) for x in w)
    # the last good model (the input) is returned
    np.testing.assert_array_equal(out["states"]["x"].cpu().numpy(), model["states"]["x"])


def test_location_aware_is_rejected():
    from keypoint_moseq_b200 import fitting
    data, meta, model = small_problem(seed=15, d=4, L=3, K=12, k=5, D=2)
    with pytest.raises(NotImplementedError):
        fitting.fit_model(model, data, meta, num_iters=1, save_every_n_iters=None, location_aware=True)


def test_unsupported_shape_fails_loudly():
    from keypoint_moseq_b200 import _lib, gibbs
    data, meta, model = small_problem(seed=16, d=3, L=2, K=6, k=5, D=2)
    dd, dm = gibbs.to_device_data(data), gibbs.to_device_model(model)
    with pytest.raises(_lib.KpmsError, match="not compiled"):
        gibbs.resample_model(dd, **dm)
    # the raw entry point refuses as well (status -3, message from the library)
    x, z = dm["states"]["x"], dm["states"]["z"]
    with pytest.raises(_lib.KpmsError, match="unsupported"):
        gibbs.sufficient_statistics(x, z, dd["mask"], 6)


def test_full_size_properties_c2():
    """BASELINE config C2 (80 chains x 10 030 frames, float32 states): size-independent invariants."""
    from keypoint_moseq_b200 import gibbs
    from keypoint_moseq_b200.synth import CONFIGS, sample_dataset
    cfg = CONFIGS["C2"]
    data, meta, model = sample_dataset(**cfg, seed=3)
    dd, dm = gibbs.to_device_data(data, "cuda", torch.float32), gibbs.to_device_model(model, "cuda", torch.float32)
    K, d, L = cfg["K"], cfg["d"], cfg["L"]
    packed = gibbs.sufficient_statistics(dm["states"]["x"], dm["states"]["z"], dd["mask"], K)
    gram, counts, _ = gibbs.unpack_statistics(packed, K, d, L)
    mask = data["mask"]
    valid = (mask[:, L:] > 0)
    F = d * L + d + 1
    assert gram[:, F - 1, F - 1].sum().item() == valid.sum()                      # every valid frame counted once
    assert counts.sum().item() == (valid[:, 1:] & valid[:, :-1]).sum()           # every valid pair counted once
    np.testing.assert_allclose(gram.cpu().numpy(), np.swapaxes(gram.cpu().numpy(), 1, 2))
    out = gibbs.resample_model(dd, **dm)
    out2 = gibbs.resample_model(dd, **dm)
    st = out["states"]
    for key in ("x", "v", "h", "s"):
        assert torch.isfinite(st[key]).all(), key
        assert torch.equal(st[key], out2["states"][key]), key                     # same seed -> same sweep
    assert st["z"].min().item() >= 0 and st["z"].max().item() < K
    assert (st["s"] > 0).all() and st["h"].abs().max().item() <= np.pi + 1e-5
    pi = out["params"]["pi"]
    np.testing.assert_allclose(pi.sum(1).cpu().numpy(), 1.0, atol=1e-12)
    assert abs(out["params"]["betas"].sum().item() - 1) < 1e-12
    ev = torch.linalg.eigvalsh(out["params"]["Q"])
    assert (ev > 0).all()
    # the resampled trajectory still explains the data at the noise scale
    Ct = gibbs.lifted_obs_matrix(out["params"]["Cd"], cfg["k"], cfg["D"]).float()
    n0 = 3
    x, v, h = st["x"][:n0], st["v"][:n0], st["h"][:n0]
    Yb = (x @ Ct[:, :-1].T + Ct[:, -1]).reshape(n0, -1, cfg["k"], cfg["D"])
    c, s_ = torch.cos(h)[..., None], torch.sin(h)[..., None]
    Yr = torch.stack([c * Yb[..., 0] - s_ * Yb[..., 1], s_ * Yb[..., 0] + c * Yb[..., 1]], -1) + v[:, :, None, :]
    resid = (dd["Y"][:n0] - Yr)[dd["mask"][:n0] > 0]
    assert resid.pow(2).mean().sqrt().item() < 1.5


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_sharded_sweep_matches_single_gpu(tmp_path):
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29617",
                          os.path.join(ROOT, "tools", "dist_check.py")], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "dist_check ok" in out.stdout


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_sharded_fit_and_apply(tmp_path):
    """`fit_model` / `apply_model` with group=WORLD on two real GPUs (checkpoint cadence, whole model on every rank,
    per-recording results); the host logic alone is covered on the CPU by tests/test_host.py with a stub sweep."""
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29631",
                          os.path.join(ROOT, "tools", "dist_fit_check.py")], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "dist_fit_check ok" in out.stdout
