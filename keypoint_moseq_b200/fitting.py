"""`fit_model` / `apply_model` drivers with the reference's signatures and checkpoint behaviour,
calling the B200 sweep instead of `jax_moseq.models.keypoint_slds.resample_model`.

Mirrors /root/reference/keypoint_moseq/fitting.py: `_wrapped_resample` (:23-44), `init_model`
(:63-106), `fit_model` (:109-287), `apply_model` (:290-425), `estimate_syllable_marginals`
(:428-559), `update_hypparams` (:562-612), `expected_marginal_likelihoods` (:615-678).
"""
import os
import warnings
from datetime import datetime
from textwrap import fill

import numpy as np
import torch

from . import _lib, gibbs
from .dist import gather_rows, shard_rows, shard_tree
from .io import SnapshotWriter, delete_snapshots_after, extract_results, load_checkpoint, save_hdf5
from .util import AsyncHostCopy, NanGuard, check_for_nans, get_nlags, to_numpy_tree, unbatch

try:  # progress bars are optional
    import tqdm
    _trange = tqdm.trange
except Exception:  # pragma: no cover
    def _trange(*args, **kwargs):
        class _Bar:
            def __init__(self, it):
                self.it = it

            def __enter__(self):
                return self

            def __exit__(self, *exc):
                return False

            def __iter__(self):
                return iter(self.it)

            def close(self):
                pass
        return _Bar(range(*args))

__all__ = ["fit_model", "apply_model", "init_model", "init_states", "update_hypparams",
           "estimate_syllable_marginals", "expected_marginal_likelihoods", "StopResampling"]


NAN_CHECK_LAG = 4      # sweeps by which the per-sweep NaN check trails the sweep being queued (0 = synchronous)


class StopResampling(Exception):
    pass


def _report_nans(model, pbar=None):
    _, _, messages = check_for_nans(model)
    if pbar is not None:
        pbar.close()
    text = ["\nEarly termination of fitting: NaNs encountered"] + [f"  - {m}" for m in messages]
    text.append("\nFor additional information, see https://keypoint-moseq.readthedocs.io/en/latest/"
                "troubleshooting.html#nans-during-fitting")
    warnings.warn("\n".join(text))


def _wrapped_resample(resample_func, data, model, pbar=None, guard=None, **resample_options):
    """One guarded sweep: Ctrl-C and NaNs end fitting and keep the last good model (fitting.py:23-44).

    With a `NanGuard` the NaN check is pipelined: this sweep's flag is queued and the flag of the sweep
    `guard.lag` iterations back is read; `guard.clean` always holds the newest model known to be clean,
    and that is what the caller returns when StopResampling is raised."""
    try:
        new = resample_func(data, **model, **resample_options)
    except KeyboardInterrupt:
        print("Early termination of fitting: user interruption")
        raise StopResampling()
    if guard is None:
        if check_for_nans(new)[0]:
            _report_nans(new, pbar)
            raise StopResampling()
        return new
    guard.submit(new)
    failed, clean = guard.collect()
    if clean is not None:
        guard.clean = clean
    if failed is not None:
        _report_nans(failed, pbar)
        raise StopResampling()
    return new


def _drain_guard(guard, pbar=None):
    """Synchronous end of the pipelined check: True when every queued sweep was clean."""
    failed, clean = guard.collect(keep=0)
    if clean is not None:
        guard.clean = clean
    if failed is not None:
        _report_nans(failed, pbar)
        return False
    return True


def _set_parallel_flag(parallel_message_passing):
    """The reference picks the time-parallel Kalman sampler on GPU (fitting.py:47-60).  Here the
    backward pass is always time-parallel, so every value is accepted and normalised to a bool."""
    if parallel_message_passing == "force" or parallel_message_passing is None:
        return True
    return bool(parallel_message_passing)


def _widen_labels(out):
    """int32 labels of the kernels -> the int64 the reference's checkpoints hold."""
    states = out.get("states") if isinstance(out, dict) else None
    if states is not None and "z" in states:
        states["z"] = np.asarray(states["z"]).astype(np.int64)
    return out


def _host_model(model):
    return _widen_labels(to_numpy_tree(model))


class _Shards:
    """One fit spread over the ranks of a process group: every rank calls `fit_model` / `apply_model`
    with the SAME full data and model plus `group=`; the rows are dealt out with `dist.shard_rows` (whole
    recordings together, balanced by valid frames), each rank sweeps its own rows, the sufficient statistics are
    all-reduced inside the sweep, and states are gathered only when a snapshot is due or the call returns.
    Rank 0 alone touches the files.  Without a group (or with one rank) every method is the identity."""

    def __init__(self, group, data, metadata):
        self.group, self.world, self.rank, self.rows = group, 1, 0, None
        if group is not None:
            import torch.distributed as dist
            self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world > 1:
            self.N = int(data["mask"].shape[0])
            self.rows_per_rank = shard_rows(data["mask"], self.world, list(metadata[0]))
            if min(len(rows) for rows in self.rows_per_rank) == 0:
                raise ValueError(f"{self.world} ranks but only {self.N} row(s): every rank needs at least one row; "
                                 "use fewer GPUs or a shorter segment length")
            self.rows = self.rows_per_rank[self.rank]

    @property
    def writes(self):
        return self.rank == 0

    def _per_row(self, leaf):
        shape = getattr(leaf, "shape", ())
        return self.rows is not None and len(shape) >= 1 and int(shape[0]) == self.N

    def agree(self, obj):
        """Rank 0's value of a small Python object on every rank (the timestamped model name)."""
        if self.world == 1:
            return obj
        import torch.distributed as dist
        box = [obj]
        dist.broadcast_object_list(box, src=dist.get_global_rank(self.group, 0), group=self.group)
        return box[0]

    def split_data(self, data):
        return data if self.rows is None else shard_tree(dict(data), self.rows)

    def split_model(self, model):
        if self.rows is None:
            return model
        out = dict(model, states=shard_tree(dict(model["states"]), self.rows))
        if self._per_row(model.get("noise_prior")):
            out["noise_prior"] = shard_tree(model["noise_prior"], self.rows)
        return out

    def join_model(self, host_model):
        """Collective: the full host model from every rank's rows (parameters are identical on all ranks)."""
        if self.rows is None:
            return host_model
        out = dict(host_model, states={key: gather_rows(val, self.rows_per_rank, self.group)
                                       for key, val in host_model["states"].items()})
        prior = host_model.get("noise_prior")
        if getattr(prior, "shape", ()) and int(prior.shape[0]) == len(self.rows):
            out["noise_prior"] = gather_rows(prior, self.rows_per_rank, self.group)
        return out


def init_states(data, params, hypparams, seed, noise_prior=None, anterior_idxs=None, posterior_idxs=None,
                error_estimator=None, dtype=torch.float64, device="cuda", **kwargs):
    """Data-driven initial states for fixed parameters (what `apply_model` needs, fitting.py:387-394):
    centroid = keypoint mean, heading from the posterior->anterior axis (0 when the indices are not
    given), x by least squares on the lifted observation operator, unit scales, then one HMM draw of z.
    Returns (states, noise_prior)."""
    dd = gibbs.to_device_data(data, device, dtype)
    Y, mask = dd["Y"], dd["mask"]
    N, T, k, D = Y.shape
    pr = {key: gibbs._dev(val, torch.float64, Y.device) for key, val in params.items()}
    v = Y.mean(2)
    if anterior_idxs is not None and posterior_idxs is not None and len(anterior_idxs) and len(posterior_idxs):
        ant = Y[:, :, list(anterior_idxs)].mean(2)
        pos = Y[:, :, list(posterior_idxs)].mean(2)
        h = torch.atan2(ant[..., 1] - pos[..., 1], ant[..., 0] - pos[..., 0])
    else:
        h = torch.zeros((N, T), dtype=dtype, device=Y.device)
    Ct = gibbs.lifted_obs_matrix(pr["Cd"], k, D)
    Yc = (Y - v[:, :, None, :]).to(torch.float64)
    c, s_ = torch.cos(h).to(torch.float64)[..., None], torch.sin(h).to(torch.float64)[..., None]
    Yr = Yc.clone()
    Yr[..., 0] = c * Yc[..., 0] + s_ * Yc[..., 1]
    Yr[..., 1] = -s_ * Yc[..., 0] + c * Yc[..., 1]
    resid = Yr.reshape(N, T, k * D) - Ct[:, -1]
    x = resid @ torch.linalg.pinv(Ct[:, :-1]).T
    if noise_prior is None:
        if error_estimator is not None and "conf" in dd:
            slope, intercept = error_estimator["slope"], error_estimator["intercept"]
            noise_prior = (10.0 ** (torch.log10(dd["conf"] + 1e-6) * slope + intercept)) ** 2
        else:
            noise_prior = torch.ones((N, T, k), dtype=dtype, device=Y.device)
    noise_prior = gibbs._dev(noise_prior, dtype, Y.device)
    z, _ = gibbs.resample_discrete_stateseqs(x.to(dtype).contiguous(), mask, pr["Ab"], pr["Q"], pr["pi"],
                                             gibbs.seed_to_u64(seed))
    states = {"x": x.to(dtype).contiguous(), "v": v.contiguous(), "h": h.contiguous(),
              "s": torch.ones((N, T, k), dtype=dtype, device=Y.device), "z": z}
    return states, noise_prior


def init_model(data=None, states=None, params=None, hypparams=None, noise_prior=None, seed=None, pca=None,
               whiten=True, location_aware=False, allo_hypparams=None, trans_hypparams=None, ar_hypparams=None,
               obs_hypparams=None, cen_hypparams=None, error_estimator=None, anterior_idxs=None,
               posterior_idxs=None, fix_heading=False, PCA_fitting_num_frames=1000000, conf_threshold=0.5,
               **kwargs):
    """Model-dict constructor (fitting.py:63-106; `kpms.init_model(data, pca=pca, **config())`).

    Anything not supplied is initialised: hyper-parameters from the four `*_hypparams` dicts of the
    config, parameters from the PCA (`Cd`, whitened latents when `whiten`) and prior draws
    (`initialize.init_params`; the PCA is fitted here when neither `params` nor `pca` is given), states
    from the data for the given parameters (`init_states`).  Passing `params` / `hypparams` of a fitted
    model re-initialises only the states, which is what `apply_model` does (fitting.py:387-394)."""
    from . import initialize
    if location_aware:
        raise NotImplementedError("location_aware=True (allo_keypoint_slds) is not implemented by the B200 sweep")
    seed = np.array([0, 0], dtype=np.uint32) if seed is None else seed
    seed_int = int(gibbs.seed_to_u64(seed) & 0x7FFFFFFF)
    if hypparams is None:
        missing = [n for n, v in (("trans_hypparams", trans_hypparams), ("ar_hypparams", ar_hypparams),
                                  ("obs_hypparams", obs_hypparams), ("cen_hypparams", cen_hypparams)) if v is None]
        if missing:
            raise ValueError(f"init_model needs `hypparams` or the config entries {missing}")
        hypparams = initialize.init_hyperparams(trans_hypparams, ar_hypparams, obs_hypparams, cen_hypparams)
    arh = hypparams["ar_hypparams"]
    _lib.check_model_dims(arh["latent_dim"], arh["nlags"],
                          arh.get("num_states", hypparams["trans_hypparams"].get("num_states", 1)))
    # (a model's own `hypparams` win over the config groups, as when apply_model re-initialises the states of a
    # fitted model with `**config()` in the call, fitting.py:387-394)
    if params is None:
        if data is None:
            raise ValueError("init_model needs `data` to initialise parameters")
        Y, mask = to_numpy_tree(data["Y"]), to_numpy_tree(data["mask"])
        if pca is None:
            pca = initialize.fit_pca(Y, mask, conf=to_numpy_tree(data["conf"]) if "conf" in data else None,
                                     anterior_idxs=anterior_idxs, posterior_idxs=posterior_idxs,
                                     conf_threshold=conf_threshold, PCA_fitting_num_frames=PCA_fitting_num_frames,
                                     fix_heading=fix_heading, seed=seed_int)
        flat, _, _ = initialize.preprocess_for_pca(Y, anterior_idxs, posterior_idxs, fix_heading)
        params = initialize.init_params(pca, hypparams, Y.shape[2], flat=flat[np.asarray(mask) > 0], whiten=whiten,
                                        seed=seed_int)
    if states is None:
        if noise_prior is None and data is not None and "conf" in data and error_estimator is not None:
            noise_prior = initialize.noise_prior_from_confidence(to_numpy_tree(data["conf"]), error_estimator)
        states, noise_prior = init_states(data, params, hypparams, seed, noise_prior=noise_prior,
                                          anterior_idxs=None if fix_heading else anterior_idxs,
                                          posterior_idxs=None if fix_heading else posterior_idxs,
                                          error_estimator=error_estimator, **kwargs)
    elif noise_prior is None:
        noise_prior = np.ones(np.asarray(to_numpy_tree(states["s"])).shape)
    return {"seed": seed, "states": states, "params": params, "hypparams": hypparams, "noise_prior": noise_prior}


def fit_model(model, data, metadata, project_dir=None, model_name=None, num_iters=50, start_iter=0,
              verbose=False, ar_only=False, parallel_message_passing=None, jitter=0.001,
              generate_progress_plots=True, save_every_n_iters=25, location_aware=False, **kwargs):
    """Fit a model to data: `num_iters - start_iter + 1` Gibbs sweeps with periodic checkpoints
    (signature and checkpoint behaviour of fitting.py:109-287).  Returns (model, model_name).

    Extra keyword arguments understood here: `dtype` (torch.float32 / torch.float64 for the states and
    the continuous-path kernels, default float64 like the reference's x64 mode), `hmm_dtype`, `group`,
    `nan_check_lag`, and `async_checkpoints` (default True: the snapshot is copied to pinned host memory on a
    side stream, `util.AsyncHostCopy`, and written by `io.SnapshotWriter` on a background thread while the next
    sweeps run - all writes have finished, and any write error has been raised, when fit_model returns; measured at
    C2: 87 ms per snapshot against 159 ms; False: snapshots are written before the next sweep is launched, as in
    the reference).
    With `group` (a torch.distributed process group, one rank per GPU) every rank passes the same full data and
    model; rows are sharded inside, rank 0 writes the gathered snapshots and every rank returns the whole model
    (see `_Shards`).
    """
    if location_aware:
        raise NotImplementedError("location_aware=True (allo_keypoint_slds) is not implemented by the B200 sweep")
    if generate_progress_plots and save_every_n_iters == 0:
        warnings.warn(fill("The `generate_progress_plots` option requires that `save_every_n_iters` be greater "
                           "than 0. Progress plots will not be generated."))
        generate_progress_plots = False
    shards = _Shards(kwargs.get("group"), data, metadata)
    if model_name is None:
        model_name = shards.agree(str(datetime.now().strftime("%Y_%m_%d-%H_%M_%S")))
    checkpoint_path = None
    if save_every_n_iters is not None:
        savedir = os.path.join(project_dir, model_name)
        checkpoint_path = os.path.join(savedir, "checkpoint.h5")
        if shards.writes:
            os.makedirs(savedir, exist_ok=True)
            print(fill(f"Outputs will be saved to {savedir}"))
            if not os.path.exists(checkpoint_path):
                save_hdf5(checkpoint_path, {"model_snapshots": {f"{start_iter}": _host_model(model)},
                                            "metadata": (np.asarray(metadata[0]), np.asarray(metadata[1])),
                                            "data": to_numpy_tree(data)})
            else:
                delete_snapshots_after(checkpoint_path, start_iter)
    data, model = shards.split_data(data), shards.split_model(model)

    parallel_message_passing = _set_parallel_flag(parallel_message_passing)
    dtype = kwargs.pop("dtype", torch.float64)
    device = kwargs.pop("device", "cuda")
    extra = {key: kwargs[key] for key in ("hmm_dtype", "group", "fix_heading", "resample_global_noise_scale",
                                          "resample_local_noise_scale") if key in kwargs}
    data_dev = gibbs.to_device_data(data, device, dtype)
    model = gibbs.to_device_model(model, device, dtype)
    resample_func = gibbs.resample_model

    # NaN check pipelined by one sweep (kwarg nan_check_lag, 0 = synchronous as in the reference): the model
    # returned after a NaN is the last one that was checked clean, as in fitting.py:30-44, :263-264
    guard = NanGuard(lag=int(kwargs.pop("nan_check_lag", NAN_CHECK_LAG)), group=shards.group if shards.world > 1 else None)
    guard.clean = model
    use_writer = kwargs.pop("async_checkpoints", True) and checkpoint_path and shards.writes
    writer = SnapshotWriter(save=save_hdf5) if use_writer else None
    try:
        model = _fit_loop(model, data_dev, resample_func, guard, writer, checkpoint_path, start_iter, num_iters,
                          save_every_n_iters, ar_only, verbose, jitter, parallel_message_passing, extra, shards)
    finally:
        if writer is not None:
            writer.close()
    if shards.world > 1:                                  # every rank returns the whole model, as with one GPU
        model = gibbs.to_device_model(shards.join_model(_host_model(model)), device, dtype)
        gibbs.release_graphs()                            # captured sweeps hold the communicator (see gibbs.release_graphs)
    return model, model_name


def _fit_loop(model, data_dev, resample_func, guard, writer, checkpoint_path, start_iter, num_iters,
              save_every_n_iters, ar_only, verbose, jitter, parallel_message_passing, extra, shards):
    with _trange(start_iter, num_iters + 1, ncols=72) as pbar:
        for iteration in pbar:
            try:
                model = _wrapped_resample(resample_func, data_dev, model, pbar=pbar, guard=guard, ar_only=ar_only,
                                          verbose=verbose, jitter=jitter,
                                          parallel_message_passing=parallel_message_passing, **extra)
            except StopResampling:
                model = guard.clean
                break
            if save_every_n_iters is not None and iteration > start_iter:
                if iteration == num_iters or (save_every_n_iters > 0 and iteration % save_every_n_iters == 0):
                    if not _drain_guard(guard, pbar):           # never checkpoint an unchecked sweep
                        model = guard.clean
                        break
                    if writer is not None and shards.world == 1:
                        # copy on a side stream into pinned memory; the writer thread waits for it, the sweeps go on
                        writer.submit(checkpoint_path, AsyncHostCopy(model, _widen_labels).result,
                                      f"model_snapshots/{iteration}")
                        continue
                    snapshot = shards.join_model(_host_model(model))      # collective when sharded
                    if writer is not None:
                        writer.submit(checkpoint_path, snapshot, f"model_snapshots/{iteration}")
                    elif shards.writes:
                        save_hdf5(checkpoint_path, snapshot, f"model_snapshots/{iteration}", exist_ok=True)
                    # progress plots (viz.plot_progress) are outside the sweep's scope and are skipped
        else:
            if not _drain_guard(guard, pbar):
                model = guard.clean
    return model


def apply_model(model, data, metadata, project_dir=None, model_name=None, num_iters=500, ar_only=False,
                save_results=True, verbose=False, results_path=None, parallel_message_passing=None,
                return_model=False, location_aware=False, overwrite=False, **kwargs):
    """Apply a fitted model to new data: states are re-initialised from the data with the parameters
    fixed and resampled `num_iters` times with `states_only=True`; results are extracted and optionally
    saved (fitting.py:290-425)."""
    if location_aware:
        raise NotImplementedError("location_aware=True (allo_keypoint_slds) is not implemented by the B200 sweep")
    parallel_message_passing = _set_parallel_flag(parallel_message_passing)
    dtype = kwargs.pop("dtype", torch.float64)
    device = kwargs.pop("device", "cuda")
    extra = {key: kwargs.pop(key) for key in ("hmm_dtype", "group", "fix_heading") if key in kwargs}
    if "fix_heading" in extra:                # the flag shapes the initial heading (0) AND freezes it in every sweep
        kwargs["fix_heading"] = extra["fix_heading"]
    shards = _Shards(extra.get("group"), data, metadata)
    if save_results and results_path is None:
        assert project_dir is not None and model_name is not None, fill(
            "The `save_results` option requires either a `results_path` or the `project_dir` and "
            "`model_name` arguments")
        results_path = os.path.join(project_dir, model_name, "results.h5")
    data_dev = gibbs.to_device_data(shards.split_data(data), device, dtype)
    model = init_model(data=data_dev, seed=model["seed"], params=model["params"], hypparams=model["hypparams"],
                       dtype=dtype, device=device, **kwargs)
    model = gibbs.to_device_model(model, device, dtype)
    # pipelined NaN check, verdict shared by the ranks of a sharded call (a rank stopping alone would desert its
    # peers in the collectives below)
    guard = NanGuard(lag=int(kwargs.pop("nan_check_lag", NAN_CHECK_LAG)), group=shards.group if shards.world > 1 else None)
    guard.clean = model
    with _trange(num_iters, ncols=72) as pbar:
        for _ in pbar:
            try:
                model = _wrapped_resample(gibbs.resample_model, data_dev, model, pbar=pbar, guard=guard, ar_only=ar_only,
                                          states_only=True, verbose=verbose,
                                          parallel_message_passing=parallel_message_passing, **extra)
            except StopResampling:
                model = guard.clean
                break
        else:
            if not _drain_guard(guard, pbar):
                model = guard.clean
    if shards.world > 1:          # states-only sweeps exchange nothing; the rows meet again here, rank 0 saves
        model = gibbs.to_device_model(shards.join_model(_host_model(model)), device, dtype)
        gibbs.release_graphs()
    results = extract_results(model, metadata, project_dir, model_name, save_results and shards.writes, results_path,
                              overwrite=overwrite)
    return (results, model) if return_model else results


def estimate_syllable_marginals(model, data, metadata, burn_in_iters=200, num_samples=100, steps_per_sample=10,
                                return_samples=False, verbose=False, parallel_message_passing=None,
                                location_aware=False, **kwargs):
    """Marginal syllable distributions by averaging HMM smoother marginals over Gibbs samples of the
    states with fixed parameters (fitting.py:428-559).  Returns {recording: (T - nlags, K) array}
    (and the samples when `return_samples`)."""
    if location_aware:
        raise NotImplementedError("location_aware=True (allo_keypoint_slds) is not implemented by the B200 sweep")
    parallel_message_passing = _set_parallel_flag(parallel_message_passing)
    dtype = kwargs.pop("dtype", torch.float64)
    device = kwargs.pop("device", "cuda")
    extra = {key: kwargs.pop(key) for key in ("hmm_dtype", "group") if key in kwargs}
    if "fix_heading" in kwargs:               # initial heading 0 (init_model) and frozen in every sweep
        extra["fix_heading"] = kwargs["fix_heading"]
    shards = _Shards(extra.get("group"), data, metadata)          # rows sharded over the ranks, joined at the end
    data_dev = gibbs.to_device_data(shards.split_data(data), device, dtype)
    model = init_model(data=data_dev, seed=model["seed"], params=model["params"], hypparams=model["hypparams"],
                       dtype=dtype, device=device, **kwargs)
    model = gibbs.to_device_model(model, device, dtype)
    total = burn_in_iters + num_samples * steps_per_sample
    acc, n_acc, samples = None, 0, []
    guard = NanGuard(lag=int(kwargs.pop("nan_check_lag", NAN_CHECK_LAG)), group=shards.group if shards.world > 1 else None)
    guard.clean = model
    with _trange(total, ncols=72) as pbar:
        for it in pbar:
            try:
                model = _wrapped_resample(gibbs.resample_model, data_dev, model, pbar=pbar, guard=guard, states_only=True,
                                          verbose=verbose, parallel_message_passing=parallel_message_passing, **extra)
            except StopResampling:
                break
            if it >= burn_in_iters and (it - burn_in_iters) % steps_per_sample == 0:
                if not _drain_guard(guard, pbar):               # only sweeps checked clean are sampled (all ranks agree)
                    break
                p = model["params"]
                marg = gibbs.stateseq_marginals(model["states"]["x"], data_dev["mask"], p["Ab"], p["Q"], p["pi"])
                acc = marg.clone() if acc is None else acc + marg
                n_acc += 1
                if return_samples:
                    samples.append(model["states"]["z"].cpu().numpy())
    nlags = get_nlags(model["params"]["Ab"])
    keys, bounds = list(metadata[0]), np.asarray(metadata[1]) + np.array([nlags, 0])
    if acc is None:
        raise RuntimeError("estimate_syllable_marginals: the sweeps stopped (NaNs or interruption) before the first "
                           "sample was taken")
    est = (acc / n_acc).cpu().numpy()          # n_acc == num_samples unless the sweeps stopped early
    if shards.world > 1:
        est = gather_rows(est, shards.rows_per_rank, shards.group)
        if samples:                                             # one collective for all samples (same count on every rank)
            stacked = gather_rows(np.moveaxis(np.asarray(samples), 0, 1), shards.rows_per_rank, shards.group)
            samples = list(np.moveaxis(stacked, 1, 0))
    marginals = unbatch(est, keys, bounds)
    marginals = {k_: np.pad(v[nlags:], ((nlags, 0), (0, 0)), mode="edge") for k_, v in marginals.items()}
    if return_samples:
        smp = unbatch(np.moveaxis(np.asarray(samples), 0, 2), keys, bounds)
        smp = {k_: np.pad(v[nlags:], ((nlags, 0), (0, 0)), mode="edge") for k_, v in smp.items()}
        return marginals, smp
    return marginals


def update_hypparams(model_dict, **kwargs):
    """Edit scalar hyper-parameters in place, casting to the old type (fitting.py:562-612)."""
    assert "hypparams" in model_dict, fill("The inputted model/checkpoint does not contain any hyperparams")
    not_updated = list(kwargs.keys())
    for group in model_dict["hypparams"]:
        for k, v in kwargs.items():
            if k in model_dict["hypparams"][group]:
                old = model_dict["hypparams"][group][k]
                if not np.isscalar(old):
                    print(fill(f"{k} cannot be updated since it is not a scalar hyperparam"))
                else:
                    if not isinstance(v, type(old)):
                        warnings.warn(f"'{k}' with {type(v)} will be cast to {type(old)}")
                    model_dict["hypparams"][group][k] = type(old)(v)
                    not_updated.remove(k)
    if len(not_updated) > 0:
        warnings.warn(fill(f"The following hypparams were not found {not_updated}"))
    return model_dict


def expected_marginal_likelihoods(project_dir=None, model_names=None, checkpoint_paths=None):
    """Expected marginal likelihood score of each model's (Ab, Q, pi) on the other models' latent
    trajectories (fitting.py:615-678).  Returns (scores, standard_errors)."""
    if checkpoint_paths is None:
        assert project_dir is not None and model_names is not None, fill(
            "Must provide either `checkpoint_paths` or `project_dir` and `model_names`")
        checkpoint_paths = [os.path.join(project_dir, name, "checkpoint.h5") for name in model_names]
    xs, params, data = [], [], None
    for path in checkpoint_paths:
        model, data, _, _ = load_checkpoint(path=path)
        xs.append(model["states"]["x"])
        params.append(model["params"])
    M = len(xs)
    mlls = np.zeros((M, M))
    for i in range(M):
        for j in range(M):
            if i != j:
                mlls[i, j] = gibbs.marginal_log_likelihood(data["mask"], xs[j], params[i]["Ab"], params[i]["Q"],
                                                           params[i]["pi"]).item()
    scores = mlls.sum(1) / (M - 1)
    variances = (mlls ** 2).sum(1) / (M - 1) - scores ** 2
    return scores, np.sqrt(variances / (M - 1))
