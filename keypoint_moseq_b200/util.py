"""Host-side helpers the sweep's callers expect (NumPy semantics).

Mirrors the `jax_moseq.utils` helpers the reference imports at
/root/reference/keypoint_moseq/fitting.py:15-16, io.py:20, util.py:15 and
viz.py:21, and the part of `format_data` (/root/reference/keypoint_moseq/util.py:929-1089)
that defines the (N, T, ...) layout the kernels consume.
"""
import warnings

import numpy as np
import torch

__all__ = [
    "batch", "unbatch", "get_nlags", "get_durations", "get_frequencies",
    "check_for_nans", "find_optimal_segment_length", "format_data", "to_numpy_tree",
    "reindex_by_bodyparts", "interpolate_keypoints", "NanGuard",
]


def _concat_stateseqs(stateseqs, mask=None):
    if isinstance(stateseqs, dict):
        return np.hstack([np.asarray(v) for v in stateseqs.values()])
    stateseqs = np.asarray(stateseqs)
    if mask is not None:
        mask = np.asarray(mask)
        return stateseqs[mask[:, -stateseqs.shape[1]:] > 0]
    return stateseqs.reshape(-1)


def get_nlags(Ab):
    """Number of AR lags from the shape of Ab (..., d, d*L+1) (fitting.py:543)."""
    return int(Ab.shape[-1] // Ab.shape[-2])


def get_durations(stateseqs, mask=None):
    """Run lengths of the (masked, concatenated) state sequence (viz.py:473, :575)."""
    z = _concat_stateseqs(stateseqs, mask)
    if z.size == 0:
        return np.zeros(0, dtype=int)
    changes = np.flatnonzero(np.diff(z) != 0) + 1
    edges = np.concatenate([[0], changes, [z.size]])
    return np.diff(edges)


def get_frequencies(stateseqs, mask=None, num_states=None, runlength=True):
    """State frequencies, by run onsets when `runlength` (io.py:598, viz.py:576)."""
    z = _concat_stateseqs(stateseqs, mask).astype(int)
    if runlength and z.size:
        keep = np.concatenate([[True], np.diff(z) != 0])
        z = z[keep]
    n = 0 if num_states is None else int(num_states)
    counts = np.bincount(z, minlength=n) if z.size else np.zeros(n)
    tot = counts.sum()
    return counts / tot if tot > 0 else counts.astype(float)


def batch(data_dict, keys=None, seg_length=None, seg_overlap=30):
    """Stack recordings into fixed-length rows (reference call: util.py:1071, :1078).

    Each recording is cut at multiples of `seg_length`; every row carries
    `seg_overlap` extra look-ahead frames and is padded to `seg_length+seg_overlap`
    by repeating its last frame with mask 0.  Returns (stack, mask, (keys, bounds)).
    """
    if keys is None:
        keys = sorted(data_dict.keys())
    lengths = [len(data_dict[k]) for k in keys]
    if seg_length is None:
        seg_length = max(lengths)
    width = seg_length + seg_overlap
    rows, masks, okeys, bounds = [], [], [], []
    for key, n in zip(keys, lengths):
        arr = np.asarray(data_dict[key])
        for start in range(0, n, seg_length):
            end = min(start + width, n)
            seg = arr[start:end]
            pad = width - (end - start)
            if pad:
                seg = np.concatenate([seg, np.repeat(seg[-1:], pad, axis=0)], axis=0)
            rows.append(seg)
            m = np.zeros(width, dtype=int)
            m[:end - start] = 1
            masks.append(m)
            okeys.append(key)
            bounds.append((start, end))
    return np.stack(rows), np.stack(masks), (okeys, np.array(bounds, dtype=int))


def unbatch(data, keys, bounds):
    """Inverse of `batch` (io.py:706-711): later rows overwrite the overlap."""
    data = np.asarray(data)
    bounds = np.asarray(bounds)
    out = {}
    for key in dict.fromkeys(keys):
        idx = [i for i, kk in enumerate(keys) if kk == key]
        length = int(bounds[idx, 1].max())
        seq = np.zeros((length,) + data.shape[2:], dtype=data.dtype)
        for i in idx:
            s, e = int(bounds[i, 0]), int(bounds[i, 1])
            if e > s:
                seq[s:e] = data[i, :e - s]
        out[key] = seq
    return out


def to_numpy_tree(tree):
    """Device -> host for a nested dict/list/tuple of tensors (the `device_get` role)."""
    if isinstance(tree, torch.Tensor):
        return tree.detach().cpu().numpy()
    if isinstance(tree, dict):
        return {k: to_numpy_tree(v) for k, v in tree.items()}
    if isinstance(tree, (list, tuple)):
        return type(tree)(to_numpy_tree(v) for v in tree)
    return tree


class AsyncHostCopy:
    """Device -> host copy of a tree of tensors that does not stall the sweeps (the reference's
    blocking `device_get` before every checkpoint write is keypoint_moseq/fitting.py:266-275).

    `AsyncHostCopy(tree)` enqueues, on a side stream that first waits for the work queued so far on the current
    stream, one non-blocking copy per CUDA leaf into pinned host memory and records an event; the caller's stream
    is never blocked and the host thread returns at once.  `result()` - typically called on the snapshot writer's
    thread - waits for that event and returns the tree as NumPy arrays (copies, so the pinned buffers go back to
    the pool).  The device tensors are kept alive and marked in use on the side stream until then; sweeps return
    fresh tensors and never write into a state that was handed out, so the copy reads a consistent snapshot.
    Host leaves pass through unchanged."""

    _pool = {}          # (shape, dtype) -> [free pinned tensors]
    _streams = {}

    def __init__(self, tree, postprocess=None):
        self._post = postprocess
        self._pairs = []            # (pinned host tensor, device tensor kept alive)
        dev = self._first_device(tree)
        self._event = None
        if dev is None:
            self._tree = tree
            return
        side = self._streams.get(dev)
        if side is None:
            side = self._streams[dev] = torch.cuda.Stream(device=dev)
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(dev))
        side.wait_event(ready)
        with torch.cuda.stream(side):
            self._tree = self._enqueue(tree, side)
            self._event = torch.cuda.Event()
            self._event.record(side)

    @classmethod
    def _first_device(cls, node):
        if isinstance(node, torch.Tensor):
            return node.device if node.is_cuda else None
        if isinstance(node, dict):
            node = list(node.values())
        if isinstance(node, (list, tuple)):
            for v in node:
                d = cls._first_device(v)
                if d is not None:
                    return d
        return None

    def _enqueue(self, node, side):
        if isinstance(node, torch.Tensor):
            if not node.is_cuda:
                return node
            key = (tuple(node.shape), node.dtype)
            free = self._pool.setdefault(key, [])
            host = free.pop() if free else torch.empty(node.shape, dtype=node.dtype, pin_memory=True)
            src = node.detach()
            host.copy_(src, non_blocking=True)
            src.record_stream(side)
            self._pairs.append((key, host, src))
            return host
        if isinstance(node, dict):
            return {k: self._enqueue(v, side) for k, v in node.items()}
        if isinstance(node, (list, tuple)):
            return type(node)(self._enqueue(v, side) for v in node)
        return node

    def result(self):
        if self._event is not None:
            self._event.synchronize()
            self._event = None
        if self._pairs:
            pinned = {id(h) for _, h, _ in self._pairs}

            def settle(node):
                if isinstance(node, torch.Tensor):
                    arr = node.numpy()
                    return arr.copy() if id(node) in pinned else arr
                if isinstance(node, dict):
                    return {k: settle(v) for k, v in node.items()}
                if isinstance(node, (list, tuple)):
                    return type(node)(settle(v) for v in node)
                return node

            self._tree = settle(self._tree)
            for key, host, _ in self._pairs:
                self._pool[key].append(host)
            self._pairs = []
        else:
            self._tree = to_numpy_tree(self._tree)
        out = self._tree
        return self._post(out) if self._post is not None else out


def check_for_nans(model):
    """(any_nans, nan_info, messages) over the leaves of a model dict (fitting.py:30).

    Device leaves are screened with one flag per leaf and ONE device->host read for the whole tree;
    the per-leaf counts are only computed when something is wrong."""
    nan_info, messages = [], []
    cuda_leaves = []

    def collect(node):
        if isinstance(node, dict):
            for v in node.values():
                collect(v)
        elif isinstance(node, (list, tuple)):
            for v in node:
                collect(v)
        elif isinstance(node, torch.Tensor) and node.is_cuda and node.is_floating_point():
            cuda_leaves.append(node)

    collect(model)
    clean_cuda = False
    if cuda_leaves:
        flags = torch.stack([torch.isnan(t).any() for t in cuda_leaves])
        clean_cuda = not bool(flags.any().item())

    def walk(node, path):
        if isinstance(node, dict):
            for k, v in node.items():
                walk(v, path + (k,))
        elif isinstance(node, (list, tuple)):
            for i, v in enumerate(node):
                walk(v, path + (i,))
        elif isinstance(node, torch.Tensor):
            if node.is_floating_point() and not (clean_cuda and node.is_cuda):
                n = int(torch.isnan(node).sum().item())
                if n:
                    nan_info.append((path, n))
                    messages.append(f"{n} NaNs found in {'/'.join(map(str, path))}")
        elif isinstance(node, np.ndarray):
            if node.dtype.kind == "f":
                n = int(np.isnan(node).sum())
                if n:
                    nan_info.append((path, n))
                    messages.append(f"{n} NaNs found in {'/'.join(map(str, path))}")
        elif isinstance(node, float) and node != node:
            nan_info.append((path, 1))
            messages.append(f"NaN found in {'/'.join(map(str, path))}")

    walk(model, ())
    return len(nan_info) > 0, nan_info, messages


class NanGuard:
    """The per-sweep NaN check of `fit_model` (fitting.py:30), pipelined: every sweep's flag is reduced on
    the device and copied to pinned host memory asynchronously, and is read `lag` sweeps later, so the
    host can queue a whole sweep ahead of the GPU instead of draining it after every sweep.  A failed
    check is reported at most `lag` sweeps late; the caller keeps the last model known to be clean.
    `lag = 0` is the synchronous check.  `group`: ranks of a sharded fit agree on every verdict."""

    def __init__(self, lag=1, group=None):
        import collections
        self.lag, self.pending, self.pool, self.group = int(lag), collections.deque(), [], group

    def _any_rank(self, flag):
        """With a process group the verdict is shared (MAX over ranks): a rank that stopped alone would leave
        the others waiting in the next sweep's all-reduce."""
        if self.group is None:
            return flag
        import torch.distributed as dist
        word = flag.to(torch.int32).reshape(1)
        dist.all_reduce(word, op=dist.ReduceOp.MAX, group=self.group)
        return word[0] > 0

    def submit(self, model):
        leaves = []

        def collect(node):
            if isinstance(node, dict):
                for v in node.values():
                    collect(v)
            elif isinstance(node, (list, tuple)):
                for v in node:
                    collect(v)
            elif isinstance(node, torch.Tensor) and node.is_cuda and node.is_floating_point():
                leaves.append(node)

        ready = getattr(model, "nan_flag", None)        # reduced on the device as part of the sweep (gibbs.SweepResult)
        if ready is None:
            collect(model)
            if not leaves:
                self.pending.append((None, None, model))
                return
            ready = torch.stack([torch.isnan(t).any() for t in leaves]).any()
        flag = self._any_rank(ready)
        host = self.pool.pop() if self.pool else torch.empty((), dtype=torch.bool).pin_memory()
        host.copy_(flag, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self.pending.append((host, ev, model))

    def collect(self, keep=None):
        """Reads every pending flag except the newest `keep` (default: lag).  Returns (failed_model,
        last_clean_model): the first model whose check failed (else None) and the newest model checked
        clean in this call (else None)."""
        keep = self.lag if keep is None else keep
        clean = None
        while len(self.pending) > keep:
            host, ev, model = self.pending.popleft()
            bad = False
            if host is None:
                bad = bool(self._any_rank(torch.tensor(bool(check_for_nans(model)[0]))))
            else:
                ev.synchronize()
                bad = bool(host.item())
                self.pool.append(host)
            if bad:
                self.pending.clear()
                return model, clean
            clean = model
        return None, clean


def find_optimal_segment_length(sequence_lengths, max_seg_length=10_000,
                                max_percent_padding=50, min_fragment_length=4):
    """Segment length rule of util.py:865-926 (longest candidate within the padding budget,
    then grown until no remainder is shorter than `min_fragment_length`)."""
    lens = np.asarray(sequence_lengths)
    if not np.all(lens > min_fragment_length):
        raise AssertionError(f"All sequences must have at least {min_fragment_length + 1} elements")
    seg = None
    for cand in sorted(set(np.minimum(lens, max_seg_length).tolist()), reverse=True):
        pad = (-lens % cand).sum() / lens.sum() * 100
        if pad <= max_percent_padding:
            seg = int(cand)
            break
    if seg is None:
        warnings.warn("No segment length found that satisfies the padding constraint. "
                      f"Using maximum value of {max_seg_length}.")
        seg = int(max_seg_length)
    while True:
        rem = lens % seg
        rem = rem[rem != 0]
        if rem.size == 0 or np.all(rem >= min_fragment_length):
            return seg
        seg += int(rem.min())


def reindex_by_bodyparts(data, bodyparts, use_bodyparts, axis=1):
    """Select / reorder keypoints by label (util.py:404-433); `data` is an array or a dict of arrays."""
    ix = np.array([list(bodyparts).index(bp) for bp in use_bodyparts])
    if isinstance(data, np.ndarray):
        return np.take(data, ix, axis)
    return {k: np.take(v, ix, axis) for k, v in data.items()}


def interpolate_keypoints(coordinates, outliers):
    """Impute outlier points by linear interpolation in time, keypoint by keypoint, holding the first /
    last good value beyond the ends; a keypoint with no good frame becomes 0 (util.py:669-690)."""
    coordinates = np.asarray(coordinates, dtype=float)
    out = np.zeros_like(coordinates)
    frames = np.arange(coordinates.shape[0])
    for i in range(coordinates.shape[1]):
        xp = np.nonzero(~outliers[:, i])[0]
        if len(xp) > 0:
            for c in range(coordinates.shape[2]):
                out[:, i, c] = np.interp(frames, xp, coordinates[xp, i, c])
    return out


def format_data(coordinates, confidences=None, keys=None, bodyparts=None, use_bodyparts=None,
                conf_pseudocount=1e-3, added_noise_level=0.1, seg_length=None, max_seg_length=10_000,
                max_percent_padding=50, min_fragment_length=4, device=None, **kwargs):
    """Batch recordings into the `data` dict and `metadata` tuple (util.py:929-1089): keypoints selected
    and ordered by `use_bodyparts`, NaN points imputed by interpolation with their confidence set to 0,
    fixed-length segments (`find_optimal_segment_length`, `batch`), confidences clamped at 0 plus
    `conf_pseudocount`, and Uniform(+-added_noise_level) noise from `default_rng(42)`.

    Arrays are returned as torch tensors on `device` (CUDA when available) in float64, the reference's
    x64 default, where the reference calls `jax.device_put`.
    """
    if keys is None:
        keys = sorted(coordinates.keys())
    else:
        bad_keys = set(keys) - set(coordinates.keys())
        assert len(bad_keys) == 0, f"Keys {bad_keys} not found in coordinates"
    assert len(keys) > 0, "No recordings found"
    num_keypoints = [coordinates[k].shape[-2] for k in keys]
    assert len(set(num_keypoints)) == 1, (f"All recordings must have the same number of keypoints, but "
                                          f"found {set(num_keypoints)} keypoints across recordings.")
    if bodyparts is not None:
        assert len(bodyparts) == num_keypoints[0], (
            f"The number of keypoints in `coordinates` ({num_keypoints[0]}) does not match the number of "
            f"labels in `bodyparts` ({len(bodyparts)})")
    if any("/" in k for k in keys):
        warnings.warn('WARNING: Recording names should not contain "/", this will cause problems with '
                      "saving/loading hdf5 files.")
    if confidences is None:
        confidences = {k: np.ones_like(coordinates[k][..., 0]) for k in keys}
    coordinates = {k: np.asarray(coordinates[k], dtype=float) for k in keys}
    confidences = {k: np.asarray(confidences[k], dtype=float) for k in keys}
    if bodyparts is not None and use_bodyparts is not None:
        coordinates = reindex_by_bodyparts(coordinates, bodyparts, use_bodyparts)
        confidences = reindex_by_bodyparts(confidences, bodyparts, use_bodyparts)
    for k in keys:
        outliers = np.isnan(coordinates[k]).any(-1)
        if outliers.any():
            coordinates[k] = interpolate_keypoints(coordinates[k], outliers)
        confidences[k] = np.where(outliers, 0, np.nan_to_num(confidences[k]))
        if not np.isfinite(coordinates[k]).all():
            raise ValueError(f"non-finite coordinates in {k!r} (infinite values are not interpolated)")
    if not seg_length:
        seg_length = find_optimal_segment_length(
            [coordinates[k].shape[0] for k in keys], max_seg_length, max_percent_padding,
            min_fragment_length)
    Y, mask, metadata = batch(coordinates, seg_length=seg_length, keys=keys)
    if not np.all(mask.sum(1) >= min_fragment_length):
        raise AssertionError(f"All segments must contain at least {min_fragment_length} frames")
    Y = Y.astype(float)
    conf = batch(confidences, seg_length=seg_length, keys=keys)[0].astype(float)
    if conf.min() < 0:
        conf = np.maximum(conf, 0)
        warnings.warn("Negative confidence values are not allowed and will be set to 0.")
    conf = conf + conf_pseudocount
    if added_noise_level > 0:
        rng = np.random.default_rng(42)
        Y += rng.uniform(-added_noise_level, added_noise_level, Y.shape)
    if device is None:
        device = "cuda" if torch.cuda.is_available() else "cpu"
    data = {"mask": torch.as_tensor(mask, device=device),
            "Y": torch.as_tensor(Y, device=device),
            "conf": torch.as_tensor(conf, device=device)}
    return data, metadata
