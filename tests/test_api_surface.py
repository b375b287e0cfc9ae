"""The drop-in boundary, pinned against the reference itself: names, parameter order and defaults of every
reference function this repo mirrors (parsed from /root/reference by tests/golden/make_signatures.py into
tests/golden/reference_signatures.json) must be accepted unchanged here.  Extra keyword parameters with
defaults are allowed; nothing the reference accepts may be missing, renamed, reordered or re-defaulted."""
import ast
import inspect
import json
import os

import pytest

GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_signatures.json")))


def _ours(qualname):
    mod, name = qualname.split(".")
    module = __import__(f"keypoint_moseq_b200.{mod}", fromlist=[name])
    return getattr(module, name)


@pytest.mark.parametrize("qualname", sorted(k for k in GOLD if not k.startswith("__")))
def test_reference_call_signature_is_accepted(qualname):
    ref = GOLD[qualname]
    sig = inspect.signature(_ours(qualname))
    params = list(sig.parameters.values())
    names = [p.name for p in params]
    has_varkw = any(p.kind is inspect.Parameter.VAR_KEYWORD for p in params)
    last = -1
    for rp in ref["params"]:
        if ref["varargs"] and rp["name"] not in names:
            continue                      # reference forwards *args to jax_moseq (init_model): keywords are checked below
        assert rp["name"] in names, f"{qualname}: parameter `{rp['name']}` of the reference (line {ref['line']}) is missing"
        idx = names.index(rp["name"])
        if not ref["varargs"]:
            assert idx > last, f"{qualname}: `{rp['name']}` is out of the reference's positional order"
            last = idx
        ours = params[idx]
        if rp["default"] is None:
            continue
        assert ours.default is not inspect.Parameter.empty, f"{qualname}: `{rp['name']}` lost its default"
        assert ours.default == ast.literal_eval(rp["default"]), (
            f"{qualname}: default of `{rp['name']}` is {ours.default!r}, the reference has {rp['default']}")
    if ref["varkw"]:
        assert has_varkw, f"{qualname}: the reference swallows unknown keywords (**{ref['varkw']}), so must we"


def test_default_config_matches_generate_config():
    """The default hyper-parameters are the reference's (generate_config, io.py:62-133), value for value."""
    from keypoint_moseq_b200.initialize import default_config
    ref = GOLD["__generate_config_defaults__"]
    cfg = default_config()
    assert set(ref) <= set(cfg)
    for key, val in ref.items():
        assert cfg[key] == val, key
    assert default_config(kappa=1e4, latent_dim=4)["trans_hypparams"]["kappa"] == 1e4
    assert default_config(latent_dim=4)["ar_hypparams"]["latent_dim"] == 4


def test_host_helpers_match_the_reference_functions():
    """tests/golden/reference_host_helpers.npz was produced by executing the reference's own pure-NumPy
    helpers (tests/golden/make_host_golden.py): the segment-length rule, keypoint interpolation and
    bodypart reindexing here must return exactly what the reference returns."""
    import warnings

    import numpy as np

    from keypoint_moseq_b200 import util
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_host_helpers.npz"))
    pos = 0
    for n_seq, want in zip(g["seg_cases"], g["seg_lengths_out"]):
        lens = g["seg_lengths_in"][pos:pos + n_seq]
        pos += n_seq
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            assert util.find_optimal_segment_length(lens) == want, lens
    pos = 0
    for n_seq, (mx, pad, frag), want in zip(g["alt_cases"], g["alt_params"], g["alt_out"]):
        lens = g["alt_in"][pos:pos + n_seq]
        pos += n_seq
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            assert util.find_optimal_segment_length(lens, int(mx), int(pad), int(frag)) == want, lens
    np.testing.assert_array_equal(util.interpolate_keypoints(g["interp_coords"], g["interp_outliers"]), g["interp_out"])
    np.testing.assert_array_equal(util.reindex_by_bodyparts(g["reindex_in"], list("abcde"), ["d", "a", "e"]),
                                  g["reindex_out"])


def test_update_hypparams_matches_the_reference_function():
    """Inputs and outputs of the reference's own update_hypparams (executed by make_host_golden.py): same
    values, same types after the cast, same number of warnings, same refusal of non-scalar entries."""
    import contextlib
    import io
    import warnings

    import numpy as np

    from keypoint_moseq_b200.fitting import update_hypparams
    recs = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden",
                                       "reference_update_hypparams.json")))
    for rec in recs:
        model = {"hypparams": {"trans_hypparams": {"num_states": 100, "gamma": 1e3, "alpha": 5.7, "kappa": 1e6},
                               "ar_hypparams": {"latent_dim": 10, "nlags": 3, "S_0_scale": 0.01, "K_0_scale": 10.0,
                                                "S_0": np.eye(2), "nu_0": 12},
                               "obs_hypparams": {"sigmasq_0": 0.1, "nu_s": 5},
                               "cen_hypparams": {"sigmasq_loc": 0.5}}}
        with warnings.catch_warnings(record=True) as w, contextlib.redirect_stdout(io.StringIO()) as so:
            warnings.simplefilter("always")
            out = update_hypparams(model, **rec["kwargs"])["hypparams"]
        for key, want in rec["result"].items():
            g_, k_ = key.split("/")
            got = out[g_][k_]
            assert type(got).__name__ == rec["types"][key], key
            assert (np.asarray(got).tolist() if not np.isscalar(got) else got) == want, key
        assert len(w) == rec["n_warnings"], rec["kwargs"]
        assert bool(so.getvalue()) == rec["printed"], rec["kwargs"]


def test_format_data_matches_the_reference_function():
    """The reference's own format_data, executed with this repo's `batch` in place of the un-vendored
    jax_moseq one and an identity `device_put` (make_host_golden.py), against util.format_data: reindexing,
    interpolation, confidence clamp + pseudocount and the Uniform(+-0.1) noise of default_rng(42) agree
    to the last bit."""
    import warnings

    import numpy as np

    from keypoint_moseq_b200.util import format_data
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_host_helpers.npz"))
    edges = np.concatenate([[0], np.cumsum(g["fd_lengths"])])
    coords = {f"rec{i}": g["fd_coords"][edges[i]:edges[i + 1]].copy() for i in range(3)}
    conf = {f"rec{i}": g["fd_conf"][edges[i]:edges[i + 1]].copy() for i in range(3)}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        data, (keys, bounds) = format_data(coords, conf, bodyparts=["p0", "p1", "p2", "p3", "p4", "p5"],
                                           use_bodyparts=["p5", "p0", "p3", "p1"], seg_length=80, device="cpu")
    np.testing.assert_array_equal(data["Y"].numpy(), g["fd_Y"])
    np.testing.assert_array_equal(data["conf"].numpy(), g["fd_conf_out"])
    np.testing.assert_array_equal(data["mask"].numpy(), g["fd_mask"])
    assert list(keys) == list(g["fd_keys"])
    np.testing.assert_array_equal(np.asarray(bounds), g["fd_bounds"])


def test_extract_results_matches_the_reference_function():
    """The reference's own extract_results (run by make_host_golden.py with this repo's `unbatch` substituted
    for jax_moseq's): syllables left-padded by repeating the first label nlags times, per-recording
    stitching of latent state, centroid and heading."""
    import numpy as np

    from keypoint_moseq_b200.io import extract_results
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_host_helpers.npz"))
    model = {"states": {k: g["er_" + k] for k in ("x", "v", "h", "z")}}
    res = extract_results(model, (list(g["er_keys"]), g["er_bounds"]), save_results=False)
    names = sorted({k.split("/")[1] for k in g.files if k.startswith("er_out/")})
    assert sorted(res) == names
    for rec in names:
        for field in ("syllable", "latent_state", "centroid", "heading"):
            np.testing.assert_array_equal(np.asarray(res[rec][field]), g[f"er_out/{rec}/{field}"])


@pytest.mark.parametrize("async_checkpoints", [False, True])
def test_fit_model_loop_matches_the_reference_loop(tmp_path, monkeypatch, async_checkpoints):
    """fit_model's control flow against the reference's own loop (fit_model + _wrapped_resample executed by
    make_host_golden.py with stubs below the boundary): the same snapshots are written, and after a NaN
    sweep the model from before it comes back - although the NaN check here trails the sweeps by
    NAN_CHECK_LAG and may let a few extra sweeps run, it never lets an unchecked sweep be saved or returned."""
    from keypoint_moseq_b200 import fitting
    recs = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_fit_loop.json")))
    saves = []

    def stub_resample(data, count=0, nan_at=None, **kw):
        import numpy as np
        new = count + 1
        bad = nan_at is not None and new >= nan_at
        return {"count": new, "nan_at": nan_at, "x": np.array([float("nan") if bad else 0.0])}

    monkeypatch.setattr(fitting.gibbs, "resample_model", stub_resample)
    monkeypatch.setattr(fitting.gibbs, "to_device_data", lambda d, *a, **k: d)
    monkeypatch.setattr(fitting.gibbs, "to_device_model", lambda m, *a, **k: m)
    monkeypatch.setattr(fitting, "_host_model", lambda m: m)
    monkeypatch.setattr(fitting, "to_numpy_tree", lambda t: t)
    monkeypatch.setattr(fitting, "save_hdf5", lambda path, d, datapath=None, **kw: saves.append(datapath or "init"))
    import warnings
    for i, rec in enumerate(recs):
        saves.clear()
        kw = dict(rec["case"])
        nan_at = kw.pop("nan_at", None)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            model, name = fitting.fit_model({"count": 0, "nan_at": nan_at}, {}, ([], []), str(tmp_path / f"p{i}"), "m",
                                            generate_progress_plots=False, async_checkpoints=async_checkpoints, **kw)
        assert name == "m"
        assert saves == rec["saves"], rec["case"]          # background writes are complete on return, in order
        assert model["count"] == rec["returned_count"], rec["case"]


def test_apply_model_loop_matches_the_reference_loop(monkeypatch):
    """apply_model's control flow against the reference's own (executed with stubs by make_host_golden.py):
    the states are re-initialised, exactly num_iters states-only sweeps run with the reference's keywords
    (no jitter), results are extracted with the reference's path / overwrite handling."""
    from keypoint_moseq_b200 import fitting
    recs = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_apply_loop.json")))
    calls = []

    def stub_resample(data, count=0, **kw):
        calls.append(kw)
        return {"count": count + 1}

    monkeypatch.setattr(fitting.gibbs, "resample_model", stub_resample)
    monkeypatch.setattr(fitting.gibbs, "to_device_data", lambda d, *a, **k: d)
    monkeypatch.setattr(fitting.gibbs, "to_device_model", lambda m, *a, **k: m)
    monkeypatch.setattr(fitting, "init_model", lambda **kw: {"count": 0})
    monkeypatch.setattr(fitting, "check_for_nans", lambda m: (False, [], []))
    monkeypatch.setattr(fitting, "extract_results",
                        lambda model, metadata, project_dir, model_name, save_results, results_path, overwrite=False:
                        {"count": model["count"], "save_results": save_results, "results_path": results_path,
                         "overwrite": overwrite})
    for rec in recs:
        calls.clear()
        res = fitting.apply_model({"seed": 1, "params": 2, "hypparams": 3}, {}, ([], []), "/proj", "name", **rec["case"])
        back = None
        if isinstance(res, tuple):
            res, back = res
        assert len(calls) == rec["sweeps"]
        for key, val in rec["kwargs_per_sweep"].items():
            if key == "parallel_message_passing":
                continue                   # normalised to a bool here: the backward pass is always time-parallel
            assert calls[0][key] == val, key
        assert "jitter" not in calls[0]
        assert res == rec["results"], rec["case"]
        assert (None if back is None else back["count"]) == rec["model_count"]


def test_estimate_syllable_marginals_loop_matches_the_reference_loop(monkeypatch):
    """estimate_syllable_marginals' control flow against the reference's own (executed with stubs by
    make_host_golden.py): burn-in, which sweeps are sampled, the average over samples, the nlags shift of the
    bounds and the edge padding of marginals and label samples."""
    import sys
    import numpy as np
    import torch
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    from keypoint_moseq_b200 import fitting
    recs = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_marginals_loop.json")))
    N, TZ, K, L, D = 2, 7, 3, 2, 1

    def pattern(count):
        base = np.arange(N * TZ * K, dtype=np.float64).reshape(N, TZ, K)
        return (base + 1.0) * (1.0 + 0.01 * count)

    def model_at(count):
        z = (np.arange(N * TZ).reshape(N, TZ) + count) % K
        return {"count": count, "seed": 1,
                "states": {"x": torch.zeros(N, TZ + L, D), "z": torch.as_tensor(z)},
                "params": {"Ab": torch.zeros(K, D, D * L + 1), "Q": torch.zeros(K, D, D), "pi": torch.eye(K)},
                "hypparams": {"trans_hypparams": {"num_states": K}}}

    calls, sampled_at = [], []

    def stub_resample(data, count=0, **kw):
        assert kw.get("states_only") is True and "jitter" not in kw
        calls.append(count + 1)
        return model_at(count + 1)

    def stub_marginals(x, mask, Ab, Q, pi, **kw):
        sampled_at.append(calls[-1])
        return torch.as_tensor(pattern(calls[-1]))

    monkeypatch.setattr(fitting.gibbs, "resample_model", stub_resample)
    monkeypatch.setattr(fitting.gibbs, "stateseq_marginals", stub_marginals)
    monkeypatch.setattr(fitting.gibbs, "to_device_data", lambda d, *a, **k: d)
    monkeypatch.setattr(fitting.gibbs, "to_device_model", lambda m, *a, **k: m)
    monkeypatch.setattr(fitting, "init_model", lambda **kw: model_at(0))
    monkeypatch.setattr(fitting, "check_for_nans", lambda m: (False, [], []))
    metadata = (["a", "b"], np.array([[0, TZ + L], [0, TZ + L - 2]]))
    data = {"mask": torch.ones(N, TZ + L)}
    for rec in recs:
        calls.clear()
        sampled_at.clear()
        res = fitting.estimate_syllable_marginals(model_at(0), data, metadata, **rec["case"])
        smp = None
        if isinstance(res, tuple):
            res, smp = res
        assert len(calls) == rec["sweeps"] and sampled_at == rec["sampled_at"], rec["case"]
        assert set(res) == set(rec["marginals"])
        for key, val in rec["marginals"].items():
            np.testing.assert_allclose(res[key], np.array(val), rtol=1e-13)
        assert (smp is None) == (rec["samples"] is None)
        if smp is not None:
            for key, val in rec["samples"].items():
                assert np.array_equal(np.asarray(smp[key]), np.array(val)), key


def test_expected_marginal_likelihoods_matches_the_reference_function(monkeypatch):
    """Which (parameters of model i, trajectory of model j) pairs are scored, and how scores and standard errors
    are formed from them (fitting.py:615-678, executed from the reference source with a stub scorer)."""
    import numpy as np
    from keypoint_moseq_b200 import fitting
    recs = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_eml.json")))

    class Scalar(float):
        def item(self):
            return float(self)

    def score(mask, x, Ab, Q, pi):
        return Scalar(np.sin(float(np.asarray(Ab).sum()) * 1.3 + float(np.asarray(x).sum()) * 0.7) * 100.0 - 500.0)

    def checkpoint(path=None):
        i = int(os.path.basename(os.path.dirname(path))[1:])
        return ({"states": {"x": np.full((2, 3), float(i))},
                 "params": {"Ab": np.full((2, 2), 10.0 + i), "Q": np.zeros(1), "pi": np.zeros(1)}},
                {"mask": np.ones((2, 3))}, None, 0)

    monkeypatch.setattr(fitting, "load_checkpoint", checkpoint)
    monkeypatch.setattr(fitting.gibbs, "marginal_log_likelihood", score)
    for rec in recs:
        if "model_names" in rec:
            sc, se = fitting.expected_marginal_likelihoods("/proj", rec["model_names"])
        else:
            sc, se = fitting.expected_marginal_likelihoods(checkpoint_paths=rec["checkpoint_paths"])
        np.testing.assert_allclose(sc, rec["scores"], rtol=1e-13)
        np.testing.assert_allclose(se, rec["standard_errors"], rtol=1e-10)
    with pytest.raises(AssertionError):
        fitting.expected_marginal_likelihoods()


# ---------------------------------------------------------------------------------------------------
# HDF5 tree layout: our writer against the reference's `_savetree_hdf5`, each reader on the other's tree
# ---------------------------------------------------------------------------------------------------
def _io_golden():
    here = os.path.dirname(os.path.abspath(__file__))
    return json.load(open(os.path.join(here, "golden", "reference_hdf5_layout.json")))


def _with_fake_h5py(monkeypatch):
    import fake_h5py
    from keypoint_moseq_b200 import io as kio
    fake_h5py.reset()
    monkeypatch.setattr(kio, "h5py", fake_h5py)
    monkeypatch.setattr(kio, "HAVE_H5PY", True)
    return fake_h5py, kio


def _untag(t):
    import numpy as np
    (kind, val), = [(k, v) for k, v in t.items() if k not in ("dtype", "shape")]
    if kind == "dict":
        return {k: _untag(v) for k, v in val}
    if kind == "list":
        return [_untag(v) for v in val]
    if kind == "tuple":
        return tuple(_untag(v) for v in val)
    if kind == "ndarray":
        dt = str if t["dtype"] == "U" else np.dtype(t["dtype"])
        return np.array(val, dtype=dt).reshape(t["shape"])
    return val


def _same_tree(a, b, path=""):
    import numpy as np
    assert type(a) is type(b) or (isinstance(a, (int, float, bool, np.generic)) and isinstance(b, (int, float, bool, np.generic))), (path, type(a), type(b))
    if isinstance(a, dict):
        assert list(a.keys()) == list(b.keys()), path
        for k in a:
            _same_tree(a[k], b[k], f"{path}/{k}")
    elif isinstance(a, (list, tuple)):
        assert len(a) == len(b), path
        for i, (x, y) in enumerate(zip(a, b)):
            _same_tree(x, y, f"{path}[{i}]")
    elif isinstance(a, np.ndarray):
        assert a.shape == b.shape and a.dtype.kind == b.dtype.kind, (path, a.dtype, b.dtype)
        if a.dtype.kind != "U":
            assert a.dtype == b.dtype, (path, a.dtype, b.dtype)
        assert np.array_equal(a, b), path
    else:
        assert a == b, (path, a, b)


@pytest.mark.parametrize("case", ["checkpoint", "misc", "results", "long_list"])
def test_hdf5_writer_produces_the_reference_layout(case, tmp_path, monkeypatch):
    """Same groups, `type` attributes, `arr{k}` names, dataset dtypes / shapes / values and string encoding
    as the reference's `save_hdf5` (io.py:1297-1400) - compared through an in-memory h5py."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_io_golden
    fake, kio = _with_fake_h5py(monkeypatch)
    gold = _io_golden()[case]
    tree = make_io_golden.cases()[case] if case != "long_list" else {"seq": [int(i) for i in range(12)]}
    path = str(tmp_path / "f.h5")
    kio.save_hdf5(path, tree)
    assert fake.File(path, "r").describe() == gold["stored"]
    if case == "checkpoint":
        snap = tree["model_snapshots"]["0"]
        kio.save_hdf5(path, snap, "model_snapshots/25", exist_ok=True)
        assert fake.File(path, "r").describe() == gold["stored_after_snapshot"]
        for label, kw in [("exists", {}), ("no_overwrite", {"exist_ok": True}),
                          ("overwrite", {"exist_ok": True, "overwrite": True})]:
            if gold["errors"][label]:
                with pytest.raises(AssertionError):
                    kio.save_hdf5(path, snap, "model_snapshots/25", **kw)
            else:
                kio.save_hdf5(path, snap, "model_snapshots/25", **kw)
        assert fake.File(path, "r").describe() == gold["stored_after_snapshot"]


@pytest.mark.parametrize("case", ["checkpoint", "misc", "results"])
def test_hdf5_reader_loads_a_reference_written_tree(case, tmp_path, monkeypatch):
    """A tree stored by the reference's writer (the fixture) loads to what the reference's own loader
    returns: container types, member order, scalars as Python scalars, strings decoded."""
    fake, kio = _with_fake_h5py(monkeypatch)
    gold = _io_golden()[case]
    path = str(tmp_path / "f.h5")
    fake.from_description(path, gold.get("stored_after_snapshot", gold["stored"]))
    if case == "checkpoint":
        _same_tree(kio.load_hdf5(path, "model_snapshots/25"), _untag(gold["loaded_datapath"]))
        loaded = kio.load_hdf5(path)
        del loaded["model_snapshots"]["25"]
        _same_tree(loaded, _untag(gold["loaded"]))
        assert kio._list_children(path, "model_snapshots") == ["0", "25"]
    else:
        _same_tree(kio.load_hdf5(path), _untag(gold["loaded"]))


def test_hdf5_reader_keeps_list_order_where_the_reference_scrambles_it(tmp_path, monkeypatch):
    """Documented divergence: h5py lists group members alphabetically, so the reference's loader returns a
    12-element list as arr0, arr1, arr10, arr11, arr2 ...; ours orders by the index in the name."""
    fake, kio = _with_fake_h5py(monkeypatch)
    gold = _io_golden()["long_list"]
    path = str(tmp_path / "f.h5")
    fake.from_description(path, gold["stored"])
    assert _untag(gold["loaded"])["seq"] == [0, 1, 10, 11, 2, 3, 4, 5, 6, 7, 8, 9]
    assert kio.load_hdf5(path)["seq"] == list(range(12))


def test_hdf5_writer_accepts_what_the_reference_refuses_only_for_numpy_scalars(tmp_path, monkeypatch):
    fake, kio = _with_fake_h5py(monkeypatch)
    gold = _io_golden()["refused"]
    import numpy as np
    assert gold == {"float32_scalar": "ValueError", "none": "ValueError", "set": "ValueError"}
    kio.save_hdf5(str(tmp_path / "a.h5"), {"a": np.float32(1.0)})          # superset: NumPy scalars are leaves
    for bad in (None, {1, 2}):
        fake.reset()
        with pytest.raises(ValueError):
            kio.save_hdf5(str(tmp_path / "b.h5"), {"a": bad})


def test_load_checkpoint_and_reindex_match_the_reference_functions(tmp_path, monkeypatch):
    """`load_checkpoint` (io.py:492-549) and `reindex_syllables_in_checkpoint` (io.py:552-619) executed from
    the reference source on a three-snapshot checkpoint: same snapshot choice, same returned trees, same
    permutation, and the same stored file afterwards (every snapshot permuted, data / metadata untouched)."""
    import sys
    import numpy as np
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_io_golden
    fake, kio = _with_fake_h5py(monkeypatch)
    gold = _io_golden()["reindex"]
    for label, kw in [("explicit", {"index": np.array([2, 0, 3, 1])}), ("by_runs", {}), ("by_frames", {"runlength": False})]:
        fake.reset()
        path = str(tmp_path / f"{label}.h5")
        kio.save_hdf5(path, make_io_golden.reindex_checkpoint())
        if label == "explicit":
            model, data, metadata, it = kio.load_checkpoint(path=path)
            assert int(it) == gold["latest"]["iteration"] == 10
            _same_tree(model, _untag(gold["latest"]["model"]))
            _same_tree(data, _untag(gold["latest"]["data"]))
            _same_tree(metadata, _untag(gold["latest"]["metadata"]))
            model, _, _, it = kio.load_checkpoint(path=path, iteration=5)
            assert int(it) == 5 and np.array_equal(model["seed"], _untag(gold["at_5"]["seed"]))
            with pytest.raises(AssertionError):
                kio.load_checkpoint(path=path, iteration=7)
        index = kio.reindex_syllables_in_checkpoint(path=path, **kw)
        assert np.asarray(index).tolist() == gold[label]["index"], label
        assert fake.File(path, "r").describe() == gold[label]["stored"], label
