"""HDF5 tree layout of the reference's checkpoints, produced by RUNNING the reference's own
`save_hdf5` / `_savetree_hdf5` / `load_hdf5` / `_loadtree_hdf5` (/root/reference/keypoint_moseq/io.py:1297-1424).

`keypoint_moseq.io` cannot be imported here (jax, h5py absent); the four functions are cut out with `ast`
and executed with `h5py` bound to tests/fake_h5py.py (an in-memory h5py with h5py's member ordering and
string conventions) and `jax.device_get` bound to the identity.  Writes tests/golden/reference_hdf5_layout.json:
for each case the stored groups / datasets and what the reference's loader returns for them.
Run in the build container: `python tests/golden/make_io_golden.py`."""
import ast
import json
import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import fake_h5py  # noqa: E402

SRC = open("/root/reference/keypoint_moseq/io.py").read()
WANT = ["save_hdf5", "load_hdf5", "_savetree_hdf5", "_loadtree_hdf5", "_get_path", "load_checkpoint",
        "reindex_syllables_in_checkpoint"]


def run_frequencies(z, mask, num_states, runlength=True):
    """Stand-in for jax_moseq.utils.get_frequencies (not vendored), restating its published rule: the masked
    frames of all rows are concatenated, then run onsets (or frames) are counted per label.  Only reached
    by the index=None branch, whose result is recorded for the control flow, not as a pin of this function."""
    z = np.asarray(z)
    flat = z[np.asarray(mask)[:, -z.shape[1]:] > 0].astype(int)
    if runlength:
        flat = flat[np.pad(np.diff(flat).nonzero()[0] + 1, (1, 0))]
    counts = np.bincount(flat, minlength=num_states)
    return counts / counts.sum()


import tqdm  # noqa: E402
from textwrap import fill  # noqa: E402
ns = {"np": np, "os": os, "h5py": fake_h5py, "jax": types.SimpleNamespace(device_get=lambda x: x), "tqdm": tqdm,
      "fill": fill, "get_frequencies": run_frequencies}
for node in ast.parse(SRC).body:
    if isinstance(node, ast.FunctionDef) and node.name in WANT:
        exec(compile(ast.Module(body=[node], type_ignores=[]), "reference/io.py", "exec"), ns)


def cases():
    """Trees shaped like the ones the sweep's callers write (fitting.py:228-236, 270-275; io.py:697-727)."""
    rng = np.random.default_rng(0)
    model = {
        "seed": np.array([0, 7], dtype=np.uint32),
        "states": {"x": rng.standard_normal((2, 5, 3)).astype(np.float32), "z": rng.integers(0, 4, (2, 3)),
                   "h": rng.standard_normal((2, 5))},
        "params": {"pi": rng.dirichlet(np.ones(3), 3), "sigmasq": np.ones(4, dtype=np.float32)},
        "hypparams": {"ar_hypparams": {"nlags": 3, "S_0_scale": 0.01, "latent_dim": 2},
                      "trans_hypparams": {"kappa": 1e6, "num_states": 3}},
        "noise_prior": 1.5,
    }
    data = {"Y": rng.standard_normal((2, 5, 4, 2)), "mask": np.ones((2, 5)), "conf": rng.uniform(size=(2, 5, 4))}
    metadata = (["rec_a", "rec_b", "rec_a"], np.array([[0, 5], [0, 5], [5, 9]]))
    misc = {"names": np.array(["nose", "tail_base", "paw"]), "label": "kappa scan", "levels": [1, 2.5, "x"],
            "nested": [{"a": np.arange(3)}, (np.float64(2.0), [np.zeros(2)])], "empty": np.zeros((0, 3)),
            "flag": True}
    results = {"rec_a": {"syllable": np.arange(5), "latent_state": rng.standard_normal((5, 2)),
                         "centroid": rng.standard_normal((5, 2)), "heading": rng.standard_normal(5)}}
    return {"checkpoint": {"model_snapshots": {"0": model}, "metadata": metadata, "data": data},
            "misc": misc, "results": results}


def reindex_checkpoint():
    """A checkpoint with the state-indexed parameters `reindex_syllables_in_checkpoint` permutes."""
    rng = np.random.default_rng(1)
    K, d, n = 4, 2, 5

    def snap(seed):
        r = np.random.default_rng(seed)
        z = np.repeat(r.integers(0, K, (2, 6)), 3, axis=1)                 # runs of three frames
        return {"seed": np.array([0, seed], dtype=np.uint32),
                "states": {"z": z, "x": r.standard_normal((2, 20, d))},
                "params": {"betas": r.dirichlet(np.ones(K)), "pi": r.dirichlet(np.ones(K), K),
                           "Ab": r.standard_normal((K, d, n)), "Q": r.standard_normal((K, d, d)),
                           "sigmasq": np.ones(3)},
                "hypparams": {"trans_hypparams": {"num_states": K}}, "noise_prior": 1.0}

    mask = np.ones((2, 20))
    mask[1, 14:] = 0
    return {"model_snapshots": {"0": snap(3), "10": snap(4), "5": snap(5)},
            "metadata": (["a", "b"], np.array([[0, 20], [0, 14]])),
            "data": {"Y": rng.standard_normal((2, 20, 3, 2)), "mask": mask}}


def tag(tree):
    """JSON form of a loaded tree that keeps container and leaf types."""
    if isinstance(tree, dict):
        return {"dict": [[k, tag(v)] for k, v in tree.items()]}
    if isinstance(tree, list):
        return {"list": [tag(v) for v in tree]}
    if isinstance(tree, tuple):
        return {"tuple": [tag(v) for v in tree]}
    if isinstance(tree, np.ndarray):
        return {"ndarray": tree.tolist(), "dtype": tree.dtype.str if tree.dtype.kind != "U" else "U", "shape": list(tree.shape)}
    if isinstance(tree, (bool, np.bool_)):
        return {"bool": bool(tree)}
    if isinstance(tree, str):
        return {"str": tree}
    if isinstance(tree, int):
        return {"int": tree}
    if isinstance(tree, float):
        return {"float": tree}
    raise TypeError(type(tree))


if __name__ == "__main__":
    out = {}
    tmp = tempfile.mkdtemp()
    for name, tree in cases().items():
        fake_h5py.reset()
        path = os.path.join(tmp, name + ".h5")
        ns["save_hdf5"](path, tree)
        entry = {"stored": fake_h5py.File(path, "r").describe(), "loaded": tag(ns["load_hdf5"](path))}
        if name == "checkpoint":
            # a later snapshot goes in by datapath, as fit_model does (fitting.py:270-275)
            snap = tree["model_snapshots"]["0"]
            ns["save_hdf5"](path, snap, "model_snapshots/25", exist_ok=True)
            entry["stored_after_snapshot"] = fake_h5py.File(path, "r").describe()
            entry["loaded_datapath"] = tag(ns["load_hdf5"](path, "model_snapshots/25"))
            errors = {}
            for label, kw in [("exists", {}), ("no_overwrite", {"exist_ok": True}),
                              ("overwrite", {"exist_ok": True, "overwrite": True})]:
                try:
                    ns["save_hdf5"](path, snap, "model_snapshots/25", **kw)
                    errors[label] = None
                except AssertionError as e:
                    errors[label] = "AssertionError"
            entry["errors"] = errors
        out[name] = entry
    # load_checkpoint / reindex_syllables_in_checkpoint (io.py:492-619) on a three-snapshot checkpoint
    entry = {}
    for label, kw in [("explicit", {"index": np.array([2, 0, 3, 1])}), ("by_runs", {}), ("by_frames", {"runlength": False})]:
        fake_h5py.reset()
        path = os.path.join(tmp, "re_" + label + ".h5")
        ns["save_hdf5"](path, reindex_checkpoint())
        if label == "explicit":
            model, data, metadata, it = ns["load_checkpoint"](path=path)
            entry["latest"] = {"iteration": int(it), "model": tag(model), "metadata": tag(metadata), "data": tag(data)}
            model, _, _, it = ns["load_checkpoint"](path=path, iteration=5)
            entry["at_5"] = {"iteration": int(it), "seed": tag(model["seed"])}
            try:
                ns["load_checkpoint"](path=path, iteration=7)
                entry["missing"] = None
            except AssertionError:
                entry["missing"] = "AssertionError"
        index = ns["reindex_syllables_in_checkpoint"](path=path, **kw)
        entry[label] = {"index": np.asarray(index).tolist(), "stored": fake_h5py.File(path, "r").describe()}
    out["reindex"] = entry
    # the reference's loader relies on the group's member order: alphabetical, so arr10 sorts before arr2
    fake_h5py.reset()
    path = os.path.join(tmp, "long.h5")
    ns["save_hdf5"](path, {"seq": [int(i) for i in range(12)]})
    out["long_list"] = {"stored": fake_h5py.File(path, "r").describe(), "loaded": tag(ns["load_hdf5"](path))}
    # what the reference refuses
    refused = {}
    for label, tree in [("float32_scalar", {"a": np.float32(1.0)}), ("none", {"a": None}), ("set", {"a": {1, 2}})]:
        fake_h5py.reset()
        try:
            ns["save_hdf5"](os.path.join(tmp, label + ".h5"), tree)
            refused[label] = None
        except Exception as e:  # noqa: BLE001
            refused[label] = type(e).__name__
    out["refused"] = refused
    with open(os.path.join(HERE, "reference_hdf5_layout.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("wrote reference_hdf5_layout.json:", {k: len(v.get("stored", {})) for k, v in out.items()})
