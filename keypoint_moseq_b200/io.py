"""Checkpoint / results layout of the reference, re-provided for the B200 sweep.

Mirrors /root/reference/keypoint_moseq/io.py: `save_hdf5` / `load_hdf5` tree format (:1297-1424),
`load_checkpoint` (:492-549) and `extract_results` (:622-727).  Files are real HDF5 when `h5py`
is importable.  When it is not (this build container has no h5py), the very same tree - same
group paths, same `type` attributes for dict / list / tuple nodes, same `arr{k}` child names - is
written as a NumPy zip archive under the requested file name, so the sweep's callers and tests run
unchanged; a one-line notice says which container format was used.
"""
import json
import os
import warnings
import zipfile

import numpy as np

from .util import to_numpy_tree, unbatch

try:  # pragma: no cover - depends on the environment
    import h5py
    HAVE_H5PY = True
except Exception:  # pragma: no cover
    h5py = None
    HAVE_H5PY = False

__all__ = ["save_hdf5", "load_hdf5", "load_checkpoint", "reindex_syllables_in_checkpoint", "extract_results", "load_results",
           "delete_snapshots_after", "SnapshotWriter", "HAVE_H5PY"]

_TYPES_KEY = "__tree_types__"
_noticed = False


def _notice():
    global _noticed
    if not _noticed and not HAVE_H5PY:
        print("keypoint_moseq_b200.io: h5py is not installed; writing the checkpoint tree as a NumPy zip "
              "archive with the same group layout")
        _noticed = True


def _get_path(project_dir, model_name, path, filename):
    if path is None:
        assert project_dir is not None and model_name is not None, (
            "`model_name` and `project_dir` are required if `path` is None.")
        path = os.path.join(project_dir, model_name, filename)
    return path


# ----------------------------------------------------------------------------
# flat representation shared by both containers: {"a/b/c": leaf}, {"a/b": "dict"|"list"|"tuple"}
# ----------------------------------------------------------------------------
def _flatten(tree, name, leaves, types):
    if isinstance(tree, np.ndarray):
        leaves[name] = tree
    elif isinstance(tree, (float, int, str, np.floating, np.integer)):
        leaves[name] = np.asarray(tree)
    elif isinstance(tree, (tuple, list)):
        types[name] = type(tree).__name__
        for k, sub in enumerate(tree):
            _flatten(sub, f"{name}/arr{k}", leaves, types)
    elif isinstance(tree, dict):
        types[name] = "dict"
        for k, sub in tree.items():
            _flatten(sub, f"{name}/{k}", leaves, types)
    else:
        raise ValueError(f"Unrecognized type {type(tree)}")


def _unflatten(prefix, leaves, types):
    if prefix in leaves:
        data = leaves[prefix]
        if data.dtype.kind in ("U", "S", "O"):
            return np.array([str(i) for i in data]) if data.shape != () else str(data.item())
        return data.item() if data.shape == () else data
    kind = types[prefix]
    children = []
    seen = set()
    for key in list(leaves) + list(types):
        if key.startswith(prefix + "/"):
            child = key[len(prefix) + 1:].split("/")[0]
            if child not in seen:
                seen.add(child)
                children.append(child)
    if kind == "dict":
        return {c: _unflatten(f"{prefix}/{c}", leaves, types) for c in children}
    ordered = sorted(children, key=lambda c: int(c[3:]))
    vals = [_unflatten(f"{prefix}/{c}", leaves, types) for c in ordered]
    return vals if kind == "list" else tuple(vals)


def _is_types_key(key):
    return key == _TYPES_KEY or key.startswith(_TYPES_KEY + "#")


def _npz_read(filepath, only=None):
    """Leaves and group types of an archive; `only` = datapath prefix to load (the rest is not decompressed).
    Group types live in one member per write (`__types__`, `__types__#1`, ...), merged in write order."""
    with np.load(filepath, allow_pickle=False) as f:
        tkeys = sorted((k for k in f.files if _is_types_key(k)), key=lambda k: int(k.partition("#")[2] or 0))
        types = {}
        for k in tkeys:
            types.update(json.loads(str(f[k])))
        want = (lambda k: True) if only is None else (lambda k: k == only or k.startswith(only + "/"))
        leaves = {k: f[k] for k in f.files if not _is_types_key(k) and want(k)}
    return leaves, types


def _npz_write(filepath, leaves, types):
    tmp = filepath + ".tmp"
    with open(tmp, "wb") as fh:
        np.savez(fh, **{_TYPES_KEY: np.asarray(json.dumps(types))}, **leaves)
    os.replace(tmp, filepath)


def _npz_names(filepath):
    """Member names (without .npy) from the central directory alone: no array is read."""
    with zipfile.ZipFile(filepath) as zf:
        return [n[:-4] if n.endswith(".npy") else n for n in zf.namelist()]


def _npz_append(filepath, leaves, types, names):
    """Add new members to an existing archive without rewriting what is there: a snapshot costs its own size,
    not the size of the checkpoint (which holds the data and every earlier snapshot)."""
    serial = 1 + max([int(n.partition("#")[2] or 0) for n in names if _is_types_key(n)], default=0)
    with zipfile.ZipFile(filepath, "a", zipfile.ZIP_STORED, allowZip64=True) as zf:
        members = dict(leaves)
        members[f"{_TYPES_KEY}#{serial}"] = np.asarray(json.dumps(types))
        for name, arr in members.items():
            with zf.open(name + ".npy", "w", force_zip64=True) as fh:
                np.lib.format.write_array(fh, np.asanyarray(arr), allow_pickle=False)


def save_hdf5(filepath, save_dict, datapath=None, exist_ok=False, overwrite=False):
    """Save a dict of pytrees (dicts / lists / tuples of arrays, scalars, strings); io.py:1297-1329."""
    assert not (os.path.exists(filepath) and not exist_ok), (
        f"{filepath} already exists. Set exist_ok to True to allow for editing an existing file.")
    save_dict = to_numpy_tree(save_dict)
    leaves, types = {}, {}
    if datapath is not None:
        _flatten(save_dict, datapath.strip("/"), leaves, types)
        roots = [datapath.strip("/")]
    else:
        for k, tree in save_dict.items():
            _flatten(tree, k, leaves, types)
        roots = list(save_dict.keys())
    if HAVE_H5PY:
        with h5py.File(filepath, "a") as f:
            for root in roots:
                assert not (not overwrite and root in f), (
                    f"{root} already exists in {f}. Set overwrite to True to overwrite data in an existing file.")
                if root in f:
                    del f[root]
            for name, kind in types.items():
                f.require_group(name).attrs["type"] = kind
            for name, arr in leaves.items():
                if arr.dtype.kind == "U":
                    f.create_dataset(name, data=arr.astype(object), dtype=h5py.special_dtype(vlen=str))
                else:
                    f.create_dataset(name, data=arr)
        return
    _notice()
    # parents of a datapath are dict groups
    for root in roots:
        parts = root.split("/")
        for i in range(1, len(parts)):
            types.setdefault("/".join(parts[:i]), "dict")
    if os.path.exists(filepath):
        names = _npz_names(filepath)
        clash = any(k == root or k.startswith(root + "/") for root in roots for k in names)
        if not clash:                                        # the usual snapshot: append, nothing is rewritten
            known = _npz_read(filepath, only="\0")[1]        # group types only
            assert not (not overwrite and any(root in known for root in roots)), (
                f"{roots} already exists in {filepath}. Set overwrite to True to overwrite data in an existing file.")
            if not any(root in known for root in roots):
                _npz_append(filepath, leaves, types, names)
                return
    old_leaves, old_types = _npz_read(filepath) if os.path.exists(filepath) else ({}, {})
    for root in roots:
        exists = any(k == root or k.startswith(root + "/") for k in list(old_leaves) + list(old_types))
        assert not (exists and not overwrite), (
            f"{root} already exists in {filepath}. Set overwrite to True to overwrite data in an existing file.")
        for store in (old_leaves, old_types):
            for k in [k for k in store if k == root or k.startswith(root + "/")]:
                del store[k]
    old_leaves.update(leaves)
    old_types.update(types)
    _npz_write(filepath, old_leaves, old_types)


class SnapshotWriter:
    """Writes checkpoint snapshots on a background thread, in submission order (the per-save
    stall of the reference, fitting.py:266-275, becomes visible once a sweep takes tens of milliseconds).

    `submit(filepath, tree, datapath)` hands over a HOST tree (the caller has already copied it off the device,
    so later sweeps cannot touch it) - or a callable that returns one, e.g. `util.AsyncHostCopy.result`, which waits
    for a device->host copy that is still in flight on a side stream - and returns at once; the thread calls
    `save(filepath, tree, datapath, exist_ok=True)`.  At most `max_pending` snapshots wait in memory (submit blocks beyond that).  `close()`
    waits for the queue to drain and raises the first write error; the error also surfaces at every later
    `submit`, and snapshots submitted after it are dropped rather than written after a hole.  Only this thread touches the file while it is open, which is what h5py's threading model asks."""

    def __init__(self, save=None, max_pending=2):
        import queue
        import threading
        self._save = save if save is not None else save_hdf5
        self._queue = queue.Queue(maxsize=max_pending)
        self._error = None
        self._thread = threading.Thread(target=self._run, name="kpms-snapshot-writer", daemon=True)
        self._thread.start()

    def _run(self):
        while True:
            job = self._queue.get()
            try:
                if job is None:
                    return
                if self._error is None:                      # after a failure later snapshots are dropped, not written
                    filepath, tree, datapath = job
                    if callable(tree):                       # device->host copy still in flight: wait for it here
                        tree = tree()
                    self._save(filepath, tree, datapath, exist_ok=True)
            except BaseException as e:  # noqa: BLE001 - handed to the caller's thread
                self._error = e
            finally:
                self._queue.task_done()

    def _raise(self):
        if self._error is not None:                          # sticky: a writer that failed stays failed
            raise self._error

    def submit(self, filepath, tree, datapath):
        self._raise()
        self._queue.put((filepath, tree, datapath))

    def close(self):
        if self._thread.is_alive():
            self._queue.put(None)
            self._thread.join()
        self._raise()

    def __enter__(self):
        return self

    def __exit__(self, exc_type, exc, tb):
        if exc_type is None:
            self.close()
        else:                                                # do not mask the caller's exception
            try:
                self.close()
            except BaseException:  # noqa: BLE001
                pass
        return False


def _h5_load(node):
    if isinstance(node, h5py.Dataset):
        data = np.array(node[()])
        if h5py.check_dtype(vlen=data.dtype) == str:
            return np.array([item.decode("utf-8") if isinstance(item, bytes) else item for item in data])
        if data.dtype.kind == "S":
            return data.item().decode("utf-8")
        if data.dtype.kind == "O" and data.shape == ():
            item = data.item()
            return item.decode("utf-8") if isinstance(item, bytes) else item
        return data.item() if data.shape == () else data
    kind = node.attrs["type"]
    if kind == "dict":
        return {k: _h5_load(v) for k, v in node.items()}
    vals = [_h5_load(node[k]) for k in sorted(node.keys(), key=lambda c: int(c[3:]))]
    if kind == "list":
        return vals
    if kind == "tuple":
        return tuple(vals)
    raise ValueError(f"Unrecognized type {kind}")


def load_hdf5(filepath, datapath=None):
    """Load a dict of pytrees written by `save_hdf5` (io.py:1332-1354)."""
    if HAVE_H5PY and not zipfile.is_zipfile(filepath):
        with h5py.File(filepath, "r") as f:
            if datapath is None:
                return {k: _h5_load(f[k]) for k in f}
            return _h5_load(f[datapath])
    leaves, types = _npz_read(filepath, only=None if datapath is None else datapath.strip("/"))
    if datapath is not None:
        return _unflatten(datapath.strip("/"), leaves, types)
    roots = []
    for key in list(leaves) + list(types):
        r = key.split("/")[0]
        if r not in roots:
            roots.append(r)
    return {r: _unflatten(r, leaves, types) for r in roots}


def _list_children(filepath, group):
    if HAVE_H5PY and not zipfile.is_zipfile(filepath):
        with h5py.File(filepath, "r") as f:
            return list(f[group].keys()) if group else list(f.keys())
    leaves, types = _npz_read(filepath)
    out = []
    prefix = group + "/" if group else ""
    for key in list(leaves) + list(types):
        if key.startswith(prefix) and key != group:
            c = key[len(prefix):].split("/")[0]
            if c not in out:
                out.append(c)
    return out


def delete_snapshots_after(checkpoint_path, start_iter):
    """Drop model snapshots later than `start_iter` (fitting.py:236-240)."""
    if HAVE_H5PY and not zipfile.is_zipfile(checkpoint_path):
        with h5py.File(checkpoint_path, "a") as f:
            for k in list(f["model_snapshots"].keys()):
                if int(k) > start_iter:
                    del f["model_snapshots"][k]
        return
    leaves, types = _npz_read(checkpoint_path)
    for store in (leaves, types):
        for key in list(store):
            parts = key.split("/")
            if parts[0] == "model_snapshots" and len(parts) > 1 and int(parts[1]) > start_iter:
                del store[key]
    _npz_write(checkpoint_path, leaves, types)


def load_checkpoint(project_dir=None, model_name=None, path=None, iteration=None):
    """(model, data, metadata, iteration) from `{project_dir}/{model_name}/checkpoint.h5`; latest
    snapshot unless `iteration` is given (io.py:492-549)."""
    path = _get_path(project_dir, model_name, path, "checkpoint.h5")
    saved = np.sort([int(i) for i in _list_children(path, "model_snapshots")])
    if iteration is None:
        iteration = int(saved[-1])
    else:
        assert iteration in saved, (f"No snapshot found for iteration {iteration}. "
                                    f"Available iterations are {saved}")
    model = load_hdf5(path, f"model_snapshots/{iteration}")
    metadata = load_hdf5(path, "metadata")
    data = load_hdf5(path, "data")
    return model, data, metadata, iteration


def reindex_syllables_in_checkpoint(project_dir=None, model_name=None, path=None, index=None, runlength=True):
    """Relabel the syllables of every snapshot of a checkpoint, in place, by their frequency in the latest
    snapshot (most frequent -> 0) or by the permutation `index` (`index[i]` is relabelled `i`); the
    state-indexed parameters betas, pi, Ab, Q are permuted with the labels (io.py:552-619).  Returns the
    permutation used."""
    from .util import get_frequencies
    path = _get_path(project_dir, model_name, path, "checkpoint.h5")
    saved = sorted(int(i) for i in _list_children(path, "model_snapshots"))
    if index is None:
        last = load_hdf5(path, f"model_snapshots/{saved[-1]}")
        mask = load_hdf5(path, "data")["mask"]
        num_states = np.asarray(last["params"]["pi"]).shape[0]
        index = np.argsort(get_frequencies(np.asarray(last["states"]["z"]), np.asarray(mask), num_states, runlength))[::-1]
    index = np.asarray(index)
    inverse = np.argsort(index)
    for iteration in saved:
        model = load_hdf5(path, f"model_snapshots/{iteration}")
        pr = model["params"]
        pr["betas"] = np.asarray(pr["betas"])[index]
        pr["pi"] = np.asarray(pr["pi"])[index, :][:, index]
        pr["Ab"] = np.asarray(pr["Ab"])[index]
        pr["Q"] = np.asarray(pr["Q"])[index]
        model["states"]["z"] = inverse[np.asarray(model["states"]["z"])]
        save_hdf5(path, model, f"model_snapshots/{iteration}", exist_ok=True, overwrite=True)
    return index


def extract_results(model, metadata, project_dir=None, model_name=None, save_results=True, path=None,
                    overwrite=False):
    """Per-recording syllables / latent state / centroid / heading, stitched with `unbatch`, optionally
    saved to `results.h5` (io.py:622-727)."""
    if save_results:
        path = _get_path(project_dir, model_name, path, "results.h5")
        if not overwrite and os.path.exists(path):
            overlap = set(metadata[0]) & set(_list_children(path, ""))
            if overlap:
                raise RuntimeError(
                    f"{path} already contains results for {len(overlap)} recording(s), including "
                    f"'{next(iter(overlap))}'. To overwrite existing results, set overwrite=True in apply_model.")
    states = to_numpy_tree(model["states"])
    keys, bounds = list(metadata[0]), np.asarray(metadata[1])
    nlags = states["x"].shape[1] - states["z"].shape[1]
    z = np.pad(np.asarray(states["z"]).astype(np.int64), ((0, 0), (nlags, 0)), mode="edge")
    syllables = unbatch(z, keys, bounds)
    latent = unbatch(states["x"], keys, bounds)
    centroid = unbatch(states["v"], keys, bounds)
    heading = unbatch(states["h"], keys, bounds)
    results = {name: {"syllable": syllables[name], "latent_state": latent[name], "centroid": centroid[name],
                      "heading": heading[name]} for name in syllables}
    if save_results:
        os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
        save_hdf5(path, results, exist_ok=True, overwrite=overwrite)
        print(f"Saved results to {path}")
    return results


def load_results(project_dir=None, model_name=None, path=None):
    """Load `results.h5` (io.py:730-748)."""
    return load_hdf5(_get_path(project_dir, model_name, path, "results.h5"))
