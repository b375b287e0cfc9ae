"""Golden vectors for the host helpers, produced by RUNNING the reference's own functions.

`keypoint_moseq.util` cannot be imported here (it imports jax at module level), but the helpers below are
pure NumPy: their source is cut out of /root/reference/keypoint_moseq/util.py with `ast` and executed in a
namespace that holds NumPy and `textwrap.fill` only.  Writes tests/golden/reference_host_helpers.npz.
Run in the build container: `python tests/golden/make_host_golden.py`."""
import ast
import os
import warnings
from textwrap import fill

import numpy as np

SRC = open("/root/reference/keypoint_moseq/util.py").read()
WANT = ["_get_percent_padding", "_find_optimal_segment_length", "interpolate_along_axis", "interpolate_keypoints",
        "reindex_by_bodyparts"]
tree = ast.parse(SRC)
ns = {"np": np, "fill": fill, "warnings": warnings}
for node in tree.body:
    if isinstance(node, ast.FunctionDef) and node.name in WANT:
        exec(compile(ast.Module(body=[node], type_ignores=[]), "reference/util.py", "exec"), ns)

rng = np.random.default_rng(0)
out = {}
# segment-length rule on a spread of cohorts
cases = [np.array([36000] * 20), np.array([108000] * 50), np.array([54000] * 200), np.array([10000] * 4),
         np.array([9500, 12000, 300, 40000]), np.array([17, 23, 8]), np.array([10003, 10001, 9999]),
         rng.integers(50, 30000, size=12), rng.integers(5, 200, size=7), np.array([10004]), np.array([20006, 5])]
lens, segs = [], []
for c in cases:
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        segs.append(int(ns["_find_optimal_segment_length"](np.asarray(c))))
    lens.append(np.asarray(c))
out["seg_cases"] = np.array([len(c) for c in lens])
out["seg_lengths_in"] = np.concatenate(lens)
out["seg_lengths_out"] = np.array(segs)
# small-parameter variants
alt = [(np.array([1000, 1500, 700]), 400, 20, 4), (np.array([90, 35, 61]), 50, 10, 8), (np.array([64, 64, 65]), 64, 50, 4)]
out["alt_in"] = np.concatenate([a[0] for a in alt])
out["alt_cases"] = np.array([len(a[0]) for a in alt])
out["alt_params"] = np.array([[a[1], a[2], a[3]] for a in alt])
res = []
for a in alt:
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        res.append(int(ns["_find_optimal_segment_length"](a[0], a[1], a[2], a[3])))
out["alt_out"] = np.array(res)
# interpolation of missing keypoints
coords = rng.standard_normal((40, 5, 3)).cumsum(0)
outl = rng.uniform(size=(40, 5)) < 0.3
outl[:, 4] = True                       # a keypoint that is never observed
outl[:3, 0] = True                      # missing at the start
outl[-4:, 1] = True                     # missing at the end
out["interp_coords"], out["interp_outliers"] = coords, outl
out["interp_out"] = ns["interpolate_keypoints"](coords, outl)
# bodypart reindexing
parts = ["a", "b", "c", "d", "e"]
use = ["d", "a", "e"]
out["reindex_in"] = coords
out["reindex_out"] = ns["reindex_by_bodyparts"](coords, parts, use)
# update_hypparams (fitting.py:562-612) is pure Python as well
import contextlib
import io
import json
fsrc = open("/root/reference/keypoint_moseq/fitting.py").read()
fns = {"np": np, "fill": fill, "warnings": warnings}
for node in ast.parse(fsrc).body:
    if isinstance(node, ast.FunctionDef) and node.name == "update_hypparams":
        exec(compile(ast.Module(body=[node], type_ignores=[]), "reference/fitting.py", "exec"), fns)


def _hyp():
    return {"hypparams": {"trans_hypparams": {"num_states": 100, "gamma": 1e3, "alpha": 5.7, "kappa": 1e6},
                          "ar_hypparams": {"latent_dim": 10, "nlags": 3, "S_0_scale": 0.01, "K_0_scale": 10.0,
                                           "S_0": np.eye(2), "nu_0": 12},
                          "obs_hypparams": {"sigmasq_0": 0.1, "nu_s": 5},
                          "cen_hypparams": {"sigmasq_loc": 0.5}}}


calls = [dict(kappa=1e4), dict(kappa=7, nu_s=6.9, sigmasq_loc=2), dict(S_0=3.0, alpha=1), dict(not_a_key=1.0, gamma=5)]
records = []
for kw in calls:
    with warnings.catch_warnings(record=True) as w, contextlib.redirect_stdout(io.StringIO()) as so:
        warnings.simplefilter("always")
        res = fns["update_hypparams"](_hyp(), **kw)["hypparams"]
    flat = {f"{g}/{k}": (v if np.isscalar(v) else np.asarray(v).tolist()) for g, d in res.items() for k, v in d.items()}
    types = {f"{g}/{k}": type(v).__name__ for g, d in res.items() for k, v in d.items()}
    records.append({"kwargs": kw, "result": flat, "types": types, "n_warnings": len(w), "printed": bool(so.getvalue())})
json.dump(records, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_update_hypparams.json"), "w"),
          indent=1)
# format_data end to end (util.py:929-1089), executed from the reference with two substitutions it cannot do
# without: jax_moseq.utils.batch (un-vendored) -> this repo's util.batch, jax.device_put -> identity.
import sys
import types
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from keypoint_moseq_b200.util import batch as our_batch  # noqa: E402
ns["batch"] = our_batch
ns["jax"] = types.SimpleNamespace(device_put=lambda tree: tree)
for node in tree.body:
    if isinstance(node, ast.FunctionDef) and node.name == "format_data":
        exec(compile(ast.Module(body=[node], type_ignores=[]), "reference/util.py", "exec"), ns)
frng = np.random.default_rng(7)
fcoords = {f"rec{i}": frng.standard_normal((n_, 6, 2)).cumsum(0) for i, n_ in enumerate((130, 95, 210))}
fconf = {k_: frng.uniform(-0.1, 1.0, v_.shape[:2]) for k_, v_ in fcoords.items()}
fcoords["rec1"][20:25, 3] = np.nan
fcoords["rec2"][0:2, 0] = np.nan
fparts = ["p0", "p1", "p2", "p3", "p4", "p5"]
fuse = ["p5", "p0", "p3", "p1"]
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    fdata, (fkeys, fbounds) = ns["format_data"]({k_: v_.copy() for k_, v_ in fcoords.items()},
                                                 {k_: v_.copy() for k_, v_ in fconf.items()}, bodyparts=fparts,
                                                 use_bodyparts=fuse, seg_length=80)
out["fd_lengths"] = np.array([130, 95, 210])
out["fd_coords"] = np.concatenate([fcoords[f"rec{i}"] for i in range(3)])
out["fd_conf"] = np.concatenate([fconf[f"rec{i}"] for i in range(3)])
out["fd_Y"], out["fd_conf_out"], out["fd_mask"] = np.asarray(fdata["Y"]), np.asarray(fdata["conf"]), np.asarray(fdata["mask"])
out["fd_keys"] = np.array(list(fkeys))
out["fd_bounds"] = np.asarray(fbounds)
# extract_results (io.py:622-727) without saving, executed from the reference with this repo's `unbatch` in
# place of the un-vendored jax_moseq one and an identity `device_get`
from keypoint_moseq_b200.util import unbatch as our_unbatch  # noqa: E402
ins = {"np": np, "os": os, "fill": fill, "unbatch": our_unbatch, "jax": types.SimpleNamespace(device_get=lambda t: t)}
for node in ast.parse(open("/root/reference/keypoint_moseq/io.py").read()).body:
    if isinstance(node, ast.FunctionDef) and node.name == "extract_results":
        exec(compile(ast.Module(body=[node], type_ignores=[]), "reference/io.py", "exec"), ins)
erng = np.random.default_rng(11)
ecoords = {"m1": np.zeros((130, 1)), "m2": np.zeros((75, 1))}
_, emask, (ekeys, ebounds) = our_batch(ecoords, seg_length=50, keys=["m1", "m2"])
eN, eT = emask.shape
estates = {"x": erng.standard_normal((eN, eT, 3)), "v": erng.standard_normal((eN, eT, 2)), "h": erng.standard_normal((eN, eT)),
           "z": erng.integers(0, 7, size=(eN, eT - 2))}
eres = ins["extract_results"]({"states": estates}, (ekeys, ebounds), save_results=False)
for name_, v_ in estates.items():
    out["er_" + name_] = v_
out["er_keys"], out["er_bounds"] = np.array(list(ekeys)), np.asarray(ebounds)
for rec_, d_ in eres.items():
    for k_, v_ in d_.items():
        out[f"er_out/{rec_}/{k_}"] = np.asarray(v_)
# fit_model's loop control (fitting.py:109-287 with _wrapped_resample :23-44), executed from the reference with
# stubs for everything below the boundary: which iterations are checkpointed, how many sweeps run, which model
# comes back when a sweep produces NaNs
import tempfile
ftree = ast.parse(fsrc)
saves = []


class _Bar:
    def __init__(self, it):
        self.it = it

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def __iter__(self):
        return iter(self.it)

    def close(self):
        pass


def _stub_resample(data, count=0, nan_at=None, **kw):
    return {"count": count + 1, "nan_at": nan_at}


fit_ns = {"np": np, "fill": fill, "warnings": warnings, "os": os, "datetime": __import__("datetime").datetime,
          "save_hdf5": lambda path, d, datapath=None, **kw: saves.append(datapath or "init"),
          "_set_parallel_flag": lambda x: x, "device_put_as_scalar": lambda m: m,
          "keypoint_slds": types.SimpleNamespace(resample_model=_stub_resample),
          "allo_keypoint_slds": types.SimpleNamespace(resample_model=_stub_resample),
          "tqdm": types.SimpleNamespace(trange=lambda a, b, **kw: _Bar(range(a, b))),
          "check_for_nans": lambda m: (m["nan_at"] is not None and m["count"] >= m["nan_at"], [], ["stub"]),
          "plot_progress": lambda *a, **k: None}
for node in ftree.body:
    if (isinstance(node, ast.ClassDef) and node.name == "StopResampling") or \
            (isinstance(node, ast.FunctionDef) and node.name in ("_wrapped_resample", "fit_model")):
        exec(compile(ast.Module(body=[node], type_ignores=[]), "reference/fitting.py", "exec"), fit_ns)
fit_cases = [dict(num_iters=10, start_iter=0, save_every_n_iters=3), dict(num_iters=10, start_iter=0, save_every_n_iters=-1),
             dict(num_iters=10, start_iter=0, save_every_n_iters=None), dict(num_iters=12, start_iter=5, save_every_n_iters=5),
             dict(num_iters=10, start_iter=0, save_every_n_iters=4, nan_at=7), dict(num_iters=3, start_iter=0, save_every_n_iters=25),
             dict(num_iters=9, start_iter=0, save_every_n_iters=2, nan_at=1)]
fit_records = []
for case in fit_cases:
    saves.clear()
    kw = dict(case)
    nan_at = kw.pop("nan_at", None)
    with tempfile.TemporaryDirectory() as tmp, warnings.catch_warnings(), contextlib.redirect_stdout(io.StringIO()):
        warnings.simplefilter("ignore")
        model, name = fit_ns["fit_model"]({"count": 0, "nan_at": nan_at}, {}, ([], []), tmp, "m",
                                          generate_progress_plots=False, **kw)
    fit_records.append({"case": case, "saves": list(saves), "returned_count": model["count"]})
json.dump(fit_records, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_fit_loop.json"), "w"), indent=1)
# apply_model's control flow (fitting.py:290-425), same technique
calls = []


def _stub_resample2(data, count=0, **kw):
    calls.append({k_: v_ for k_, v_ in sorted(kw.items()) if k_ not in ("tag",)})
    return {"count": count + 1, "tag": kw.get("tag", None) or "m"}


app_ns = dict(fit_ns)
app_ns.update({"jax": types.SimpleNamespace(device_put=lambda t: t),
               "init_model": lambda **kw: {"count": 0, "tag": "init:" + ",".join(sorted(kw))},
               "keypoint_slds": types.SimpleNamespace(resample_model=_stub_resample2),
               "allo_keypoint_slds": types.SimpleNamespace(resample_model=_stub_resample2),
               "tqdm": types.SimpleNamespace(trange=lambda n, **kw: _Bar(range(n))),
               "check_for_nans": lambda m: (False, [], []),
               "extract_results": lambda model, metadata, project_dir, model_name, save_results, results_path, overwrite=False:
                   {"count": model["count"], "save_results": save_results, "results_path": results_path, "overwrite": overwrite}})
for node in ftree.body:
    if isinstance(node, ast.FunctionDef) and node.name in ("_wrapped_resample", "apply_model"):
        exec(compile(ast.Module(body=[node], type_ignores=[]), "reference/fitting.py", "exec"), app_ns)
app_cases = [dict(num_iters=7), dict(num_iters=3, ar_only=True, save_results=False, return_model=True),
             dict(num_iters=2, results_path="/x/y.h5", overwrite=True, verbose=True)]
app_records = []
for case in app_cases:
    calls.clear()
    with contextlib.redirect_stdout(io.StringIO()):
        res = app_ns["apply_model"]({"seed": 1, "params": 2, "hypparams": 3}, {}, ([], []), "/proj", "name", **case)
    model_back = None
    if isinstance(res, tuple):
        res, model_back = res
    app_records.append({"case": case, "sweeps": len(calls), "kwargs_per_sweep": calls[0], "results": res,
                        "model_count": None if model_back is None else model_back["count"]})
json.dump(app_records, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_apply_loop.json"), "w"), indent=1)
# estimate_syllable_marginals' control flow (fitting.py:428-559): which sweeps are sampled, the averaging, the
# nlags shift of the bounds and the edge padding; `unbatch` / `get_nlags` are this repo's (jax_moseq's are absent)
from keypoint_moseq_b200.util import get_nlags as _get_nlags, unbatch as _unbatch  # noqa: E402

MARG_N, MARG_TZ, MARG_K, MARG_L, MARG_D = 2, 7, 3, 2, 1


def marg_pattern(count):
    """Deterministic stand-in for the smoother marginals of the model after `count` sweeps."""
    base = np.arange(MARG_N * MARG_TZ * MARG_K, dtype=np.float64).reshape(MARG_N, MARG_TZ, MARG_K)
    return (base + 1.0) * (1.0 + 0.01 * count)


def z_pattern(count):
    return (np.arange(MARG_N * MARG_TZ).reshape(MARG_N, MARG_TZ) + count) % MARG_K


def marg_model(count):
    return {"count": count, "seed": 1,
            "states": {"x": np.zeros((MARG_N, MARG_TZ + MARG_L, MARG_D)), "z": z_pattern(count)},
            "params": {"Ab": np.zeros((MARG_K, MARG_D, MARG_D * MARG_L + 1)), "Q": np.zeros((MARG_K, MARG_D, MARG_D)),
                       "pi": np.eye(MARG_K)},
            "hypparams": {"trans_hypparams": {"num_states": MARG_K}}}


marg_calls = []


def _stub_resample3(data, count=0, **kw):
    marg_calls.append(count + 1)
    return marg_model(count + 1)


sampled_at = []


def _stub_marginals(x, mask, Ab=None, Q=None, pi=None, **kw):
    sampled_at.append(marg_calls[-1])
    return marg_pattern(marg_calls[-1])


marg_ns = dict(app_ns)
marg_ns.update({"init_model": lambda **kw: marg_model(0), "stateseq_marginals": _stub_marginals,
                "keypoint_slds": types.SimpleNamespace(resample_model=_stub_resample3),
                "get_nlags": _get_nlags, "unbatch": _unbatch})
for node in ftree.body:
    if isinstance(node, ast.FunctionDef) and node.name in ("_wrapped_resample", "estimate_syllable_marginals"):
        exec(compile(ast.Module(body=[node], type_ignores=[]), "reference/fitting.py", "exec"), marg_ns)
marg_meta = (["a", "b"], np.array([[0, MARG_TZ + MARG_L], [0, MARG_TZ + MARG_L - 2]]))
marg_data = {"mask": np.ones((MARG_N, MARG_TZ + MARG_L))}
marg_records = []
for case in [dict(burn_in_iters=3, num_samples=2, steps_per_sample=4), dict(burn_in_iters=0, num_samples=3, steps_per_sample=1, return_samples=True),
             dict(burn_in_iters=5, num_samples=1, steps_per_sample=2, return_samples=True)]:
    marg_calls.clear()
    sampled_at.clear()
    with contextlib.redirect_stdout(io.StringIO()):
        res = marg_ns["estimate_syllable_marginals"](marg_model(0), marg_data, marg_meta, **case)
    smp = None
    if isinstance(res, tuple):
        res, smp = res
    marg_records.append({"case": case, "sweeps": len(marg_calls), "sampled_at": list(sampled_at),
                         "marginals": {k_: v_.tolist() for k_, v_ in res.items()},
                         "samples": None if smp is None else {k_: np.asarray(v_).tolist() for k_, v_ in smp.items()}})
json.dump(marg_records, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_marginals_loop.json"), "w"), indent=1)
# expected_marginal_likelihoods (fitting.py:615-678): which (parameters, trajectory) pairs are scored and how
# the scores and standard errors are formed
class _Scalar(float):
    def item(self):
        return float(self)


def eml_score(mask, x, Ab, Q, pi):
    return _Scalar(np.sin(float(np.asarray(Ab).sum()) * 1.3 + float(np.asarray(x).sum()) * 0.7) * 100.0 - 500.0)


def eml_checkpoint(path):
    i = int(os.path.basename(os.path.dirname(path))[1:])
    return ({"states": {"x": np.full((2, 3), float(i))},
             "params": {"Ab": np.full((2, 2), 10.0 + i), "Q": np.zeros(1), "pi": np.zeros(1)}},
            {"mask": np.ones((2, 3))}, None, 0)


eml_ns = dict(fit_ns)
eml_ns.update({"jnp": types.SimpleNamespace(array=lambda a: a), "load_checkpoint": lambda path=None: eml_checkpoint(path),
               "marginal_log_likelihood": eml_score, "os": os,
               "tqdm": types.SimpleNamespace(trange=lambda n, **kw: range(n))})
for node in ftree.body:
    if isinstance(node, ast.FunctionDef) and node.name == "expected_marginal_likelihoods":
        exec(compile(ast.Module(body=[node], type_ignores=[]), "reference/fitting.py", "exec"), eml_ns)
eml_records = []
for names in (["m0", "m1", "m2"], ["m3", "m1", "m4", "m0", "m2"]):
    sc, se = eml_ns["expected_marginal_likelihoods"]("/proj", names)
    eml_records.append({"model_names": names, "scores": np.asarray(sc).tolist(), "standard_errors": np.asarray(se).tolist()})
sc, se = eml_ns["expected_marginal_likelihoods"](checkpoint_paths=["/a/m2/checkpoint.h5", "/b/m0/checkpoint.h5"])
eml_records.append({"checkpoint_paths": ["/a/m2/checkpoint.h5", "/b/m0/checkpoint.h5"], "scores": np.asarray(sc).tolist(),
                    "standard_errors": np.asarray(se).tolist()})
json.dump(eml_records, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_eml.json"), "w"), indent=1)
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_host_helpers.npz"), **out)
print("segment lengths:", segs, "update_hypparams cases:", len(records))
