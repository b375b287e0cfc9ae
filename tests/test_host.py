"""Host-side logic that needs no GPU: batching helpers, checkpoint tree I/O, sharding and the
world_size-2 all-reduce of packed statistics (gloo)."""
import os
import sys

import numpy as np
import pytest
import torch

import oracle as orc
from keypoint_moseq_b200 import io as kio
from keypoint_moseq_b200 import util
from keypoint_moseq_b200.dist import shard_rows, shard_tree
from keypoint_moseq_b200.fitting import update_hypparams
from keypoint_moseq_b200.gibbs import advance_seed, seed_to_u64
from keypoint_moseq_b200.synth import default_hypparams, sample_dataset


def test_batch_unbatch_round_trip_and_overlap():
    rng = np.random.default_rng(0)
    recs = {"a": rng.standard_normal((95, 3)), "b": rng.standard_normal((40, 3)), "c": rng.standard_normal((41, 3))}
    stack, mask, (keys, bounds) = util.batch(recs, seg_length=40, seg_overlap=5)
    assert stack.shape == (3 + 1 + 2, 45, 3)
    assert keys == ["a", "a", "a", "b", "c", "c"]
    assert bounds.tolist() == [[0, 45], [40, 85], [80, 95], [0, 40], [0, 41], [40, 41]]
    assert mask.sum(1).tolist() == [45, 45, 15, 40, 41, 1]
    # padding repeats the last real frame
    np.testing.assert_array_equal(stack[2, 15:], np.repeat(recs["a"][-1:], 30, axis=0))
    back = util.unbatch(stack, keys, bounds)
    for k in recs:
        np.testing.assert_array_equal(back[k], recs[k])
    # time-shortened arrays (z has T - nlags frames): shift the start bound
    z = stack[:, 3:, 0]
    back_z = util.unbatch(z, keys, bounds + np.array([3, 0]))
    np.testing.assert_array_equal(back_z["a"][3:], recs["a"][3:, 0])


def test_durations_frequencies_nlags():
    z = np.array([[0, 0, 1, 1, 1, 2], [2, 2, 2, 0, 0, 0]])
    mask = np.ones((2, 8), dtype=int)
    mask[1, 6:] = 0      # last two frames of row 1 are padding -> z[1, 4:] dropped
    assert util.get_durations(z, mask).tolist() == [2, 3, 4, 1]
    f = util.get_frequencies(z, mask, num_states=4, runlength=True)
    np.testing.assert_allclose(f, [2 / 4, 1 / 4, 1 / 4, 0])
    f2 = util.get_frequencies({"a": z[0], "b": z[1]}, num_states=3, runlength=False)
    np.testing.assert_allclose(f2, [5 / 12, 3 / 12, 4 / 12])
    assert util.get_nlags(np.zeros((7, 4, 13))) == 3


def test_segment_length_rule():
    assert util.find_optimal_segment_length([36000] * 20) == 10000
    assert util.find_optimal_segment_length([500, 800, 1200]) == 1200        # 45.6% padding <= 50%
    seg = util.find_optimal_segment_length([10003, 20000])
    assert all(r == 0 or r >= 4 for r in np.array([10003, 20000]) % seg)
    with pytest.raises(AssertionError):
        util.find_optimal_segment_length([3, 100])


def test_format_data_layout():
    rng = np.random.default_rng(1)
    coords = {"r1": rng.standard_normal((120, 5, 2)), "r0": rng.standard_normal((70, 5, 2))}
    conf = {k: rng.uniform(0, 1, v.shape[:2]) for k, v in coords.items()}
    data, (keys, bounds) = util.format_data(coords, conf, seg_length=50, device="cpu")
    assert keys == ["r0", "r0", "r1", "r1", "r1"]
    assert tuple(data["Y"].shape) == (5, 80, 5, 2) and tuple(data["conf"].shape) == (5, 80, 5)
    assert data["mask"].sum().item() == 70 + 20 + 80 + 70 + 20
    noise = data["Y"][0, :70].numpy() - coords["r0"]
    assert np.abs(noise).max() <= 0.1 and np.abs(noise).max() > 0.05
    assert float(data["conf"].min()) >= 1e-3


def test_check_for_nans():
    model = {"states": {"x": torch.zeros(3, 4), "z": torch.zeros(3, 2, dtype=torch.int32)},
             "params": {"Ab": np.zeros((2, 2))}, "seed": np.array([0, 1], dtype=np.uint32)}
    assert util.check_for_nans(model)[0] is False
    model["states"]["x"][1, 2] = float("nan")
    any_nans, info, msgs = util.check_for_nans(model)
    assert any_nans and "states/x" in msgs[0]


def test_hdf5_tree_round_trip_and_guards(tmp_path):
    _, meta, model = sample_dataset(recordings=2, frames=60, k=4, D=2, d=2, L=2, K=3, seg_length=40)
    path = str(tmp_path / "checkpoint.h5")
    tree = {"model_snapshots": {"0": model}, "metadata": (np.asarray(meta[0]), meta[1]), "list": [1, 2.5, "s"]}
    kio.save_hdf5(path, tree)
    with pytest.raises(AssertionError):
        kio.save_hdf5(path, tree)                                   # file exists, exist_ok False
    kio.save_hdf5(path, model, "model_snapshots/25", exist_ok=True)
    with pytest.raises(AssertionError):
        kio.save_hdf5(path, model, "model_snapshots/25", exist_ok=True)   # group exists, overwrite False
    kio.save_hdf5(path, model, "model_snapshots/50", exist_ok=True)
    m, _, md, it = None, None, None, None
    loaded = kio.load_hdf5(path)
    assert sorted(loaded["model_snapshots"]) == ["0", "25", "50"]
    assert isinstance(loaded["metadata"], tuple) and list(loaded["metadata"][0]) == list(meta[0])
    assert loaded["list"] == [1, 2.5, "s"]
    snap = kio.load_hdf5(path, "model_snapshots/25")
    np.testing.assert_array_equal(snap["states"]["x"], model["states"]["x"])
    assert snap["hypparams"]["trans_hypparams"]["num_states"] == 3
    assert isinstance(snap["hypparams"]["obs_hypparams"]["nu_s"], int)
    kio.delete_snapshots_after(path, 25)
    assert sorted(kio.load_hdf5(path)["model_snapshots"]) == ["0", "25"]


def test_load_checkpoint_and_extract_results(tmp_path):
    data, meta, model = sample_dataset(recordings=2, frames=70, k=4, D=2, d=2, L=2, K=3, seg_length=40)
    d = tmp_path / "proj" / "m"
    d.mkdir(parents=True)
    kio.save_hdf5(str(d / "checkpoint.h5"), {"model_snapshots": {"0": model, "10": model},
                                              "metadata": (np.asarray(meta[0]), meta[1]), "data": data})
    m, dat, md, it = kio.load_checkpoint(str(tmp_path / "proj"), "m")
    assert it == 10 and dat["Y"].shape == data["Y"].shape
    res = kio.extract_results(m, md, str(tmp_path / "proj"), "m")
    for name in ("rec0000", "rec0001"):
        assert res[name]["syllable"].shape == (70,) and res[name]["latent_state"].shape == (70, 2)
        assert res[name]["centroid"].shape == (70, 2) and res[name]["heading"].shape == (70,)
        # the first nlags labels repeat the first sampled label
        assert (res[name]["syllable"][:2] == res[name]["syllable"][2]).all()
    with pytest.raises(RuntimeError):
        kio.extract_results(m, md, str(tmp_path / "proj"), "m")     # recordings already present
    kio.extract_results(m, md, str(tmp_path / "proj"), "m", overwrite=True)
    again = kio.load_results(str(tmp_path / "proj"), "m")
    np.testing.assert_array_equal(again["rec0000"]["syllable"], res["rec0000"]["syllable"])


def test_update_hypparams_and_seed():
    model = {"hypparams": default_hypparams(4, 3, 10)}
    with pytest.warns(UserWarning):
        update_hypparams(model, kappa=5, not_there=1)
    assert model["hypparams"]["trans_hypparams"]["kappa"] == 5.0
    assert isinstance(model["hypparams"]["trans_hypparams"]["kappa"], float)
    s0 = np.array([3, 7], dtype=np.uint32)
    assert seed_to_u64(s0) == (3 << 32) | 7
    s1 = advance_seed(s0)
    assert s1.dtype == np.uint32 and s1.shape == (2,) and not np.array_equal(s0, s1)
    assert np.array_equal(advance_seed(s0), s1)


def test_shard_rows_balances_and_keeps_recordings_together():
    mask = np.zeros((7, 100), dtype=int)
    lens = [100, 100, 30, 100, 60, 100, 10]
    for i, n in enumerate(lens):
        mask[i, :n] = 1
    keys = ["a", "a", "a", "b", "b", "c", "c"]
    shards = shard_rows(mask, 2, keys)
    assert sorted(np.concatenate(shards).tolist()) == list(range(7))
    loads = [mask[s].sum() for s in shards]
    assert abs(loads[0] - loads[1]) <= 100
    for key in "abc":
        owners = {r for r, s in enumerate(shards) for i in s if keys[i] == key}
        assert len(owners) == 1
    sub = shard_tree({"Y": np.arange(7)[:, None] * np.ones((7, 3)), "m": torch.arange(7)}, shards[0])
    assert sub["Y"].shape[0] == len(shards[0]) and sub["m"].tolist() == shards[0].tolist()
    # one long recording on four ranks: whole recordings cannot balance, single rows are dealt out instead
    shards = shard_rows(np.ones((8, 50), dtype=int), 4, ["a"] * 8)
    assert sorted(len(s) for s in shards) == [2, 2, 2, 2]
    assert sorted(np.concatenate(shards).tolist()) == list(range(8))


def _gloo_worker(rank, world, port, tmp):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from keypoint_moseq_b200.dist import allreduce_statistics
    data, meta, model = sample_dataset(recordings=3, frames=90, k=4, D=2, d=2, L=2, K=4, seed=5, seg_length=50)
    K, d, L = 4, 2, 2
    shards = shard_rows(data["mask"], world, meta[0])
    rows = shards[rank]
    x, z, mask = model["states"]["x"][rows], model["states"]["z"][rows], data["mask"][rows]
    gram = orc.ar_suffstats(x, z, mask, K)        # the statistics the CUDA kernels produce, from the checker
    counts = orc.count_transitions(z, mask, K)
    packed = torch.tensor(np.concatenate([gram.reshape(-1), counts.reshape(-1).astype(float)]))
    allreduce_statistics(packed)
    np.save(os.path.join(tmp, f"packed{rank}.npy"), packed.numpy())
    dist.destroy_process_group()


def test_two_rank_statistics_allreduce_matches_single_rank(tmp_path):
    import torch.multiprocessing as mp
    port = 29500 + os.getpid() % 2000
    mp.spawn(_gloo_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    p0, p1 = (np.load(tmp_path / f"packed{r}.npy") for r in range(2))
    np.testing.assert_array_equal(p0, p1)          # every rank holds the same reduced buffer
    data, meta, model = sample_dataset(recordings=3, frames=90, k=4, D=2, d=2, L=2, K=4, seed=5, seg_length=50)
    gram = orc.ar_suffstats(model["states"]["x"], model["states"]["z"], data["mask"], 4)
    counts = orc.count_transitions(model["states"]["z"], data["mask"], 4)
    F = gram.shape[-1]
    np.testing.assert_allclose(p0[:4 * F * F].reshape(4, F, F), gram, rtol=1e-12, atol=1e-12)
    np.testing.assert_array_equal(p0[4 * F * F:].reshape(4, 4), counts)


def test_fit_pca_and_init_params_recover_the_pose_subspace():
    """fit_pca + init_params on data from the generative model: the PCA has (k-1)*D features, the
    whitened Cd reproduces the aligned coordinates from unit-covariance latents, prior draws have the
    checkpoint shapes."""
    from keypoint_moseq_b200 import initialize
    from keypoint_moseq_b200.synth import default_hypparams, sample_dataset
    data, _, model = sample_dataset(recordings=2, frames=600, k=6, D=2, d=4, L=3, K=12, seed=3, seg_length=300)
    pca = initialize.fit_pca(data["Y"], data["mask"], conf=data["conf"], conf_threshold=0.0,
                             anterior_idxs=[0, 1], posterior_idxs=[4, 5])
    assert pca.mean_.shape == (10,) and pca.components_.shape[1] == 10
    assert pca.explained_variance_ratio_[:4].sum() > 0.8          # 4 latent dims generate the poses (plus noise)
    cfgh = default_hypparams(4, 3, 12)
    hyp = initialize.init_hyperparams(cfgh["trans_hypparams"], {k_: cfgh["ar_hypparams"][k_] for k_ in
                                      ("latent_dim", "nlags", "S_0_scale", "K_0_scale")},
                                      cfgh["obs_hypparams"], cfgh["cen_hypparams"])
    assert hyp["ar_hypparams"]["M_0"].shape == (4, 13) and hyp["ar_hypparams"]["nu_0"] == 6
    flat, v, h = initialize.preprocess_for_pca(data["Y"], [0, 1], [4, 5])
    rows = flat[data["mask"] > 0]
    pr = initialize.init_params(pca, hyp, 6, flat=rows, whiten=True, seed=1)
    assert pr["Cd"].shape == (10, 5) and pr["Ab"].shape == (12, 4, 13) and pr["Q"].shape == (12, 4, 4)
    assert np.allclose(pr["pi"].sum(1), 1.0) and np.all(np.linalg.eigvalsh(pr["Q"]) > 0)
    x = (rows - pr["Cd"][:, -1]) @ np.linalg.pinv(pr["Cd"][:, :-1]).T
    assert np.allclose(np.cov(x, rowvar=False), np.eye(4), atol=1e-6)          # whitened latents
    recon = x @ pr["Cd"][:, :-1].T + pr["Cd"][:, -1]
    assert np.sqrt(((recon - rows) ** 2).mean()) < 0.5 * rows.std()
    prior = initialize.noise_prior_from_confidence(data["conf"], {"slope": -0.5, "intercept": 0.25})
    assert prior.shape == data["conf"].shape and np.all(prior > 0)


def test_nan_guard_pipelines_the_check_and_keeps_the_last_clean_model():
    """NanGuard on host leaves (the device path copies one flag per sweep asynchronously): a failed
    check is reported `lag` sweeps late and the newest clean model is handed back."""
    from keypoint_moseq_b200.util import NanGuard
    good = [{"states": {"x": np.ones(3) * i}} for i in range(4)]
    bad = {"states": {"x": np.array([1.0, np.nan])}}
    guard = NanGuard(lag=1)
    seen = []
    for model in (good[0], good[1], bad, good[2], good[3]):
        guard.submit(model)
        failed, clean = guard.collect()
        seen.append((failed is not None, None if clean is None else float(clean["states"]["x"][0])))
        if failed is not None:
            assert failed is bad
            break
    # sweep 0: nothing old enough; sweep 1: model 0 clean; sweep 2 (the bad one): model 1 clean; sweep 3: failure
    assert seen == [(False, None), (False, 0.0), (False, 1.0), (True, None)]
    guard = NanGuard(lag=0)
    guard.submit(bad)
    assert guard.collect()[0] is bad
    guard = NanGuard(lag=2)
    for model in good[:3]:
        guard.submit(model)
    failed, clean = guard.collect(keep=0)
    assert failed is None and clean is good[2]


def test_reindex_syllables_in_checkpoint(tmp_path):
    """Relabelling by frequency permutes z and the state-indexed parameters of every snapshot consistently
    (io.py:552-619): transition structure and likelihood parameters follow their syllable."""
    from keypoint_moseq_b200.util import get_frequencies
    data, meta, model = sample_dataset(recordings=2, frames=200, k=4, D=2, d=2, L=2, K=5, seg_length=120)
    d = tmp_path / "proj" / "m"
    d.mkdir(parents=True)
    path = str(d / "checkpoint.h5")
    kio.save_hdf5(path, {"model_snapshots": {"0": model, "7": model}, "metadata": (np.asarray(meta[0]), meta[1]),
                         "data": data})
    index = kio.reindex_syllables_in_checkpoint(str(tmp_path / "proj"), "m")
    assert sorted(index.tolist()) == list(range(5))
    for it in (0, 7):
        new = kio.load_hdf5(path, f"model_snapshots/{it}")
        old = model
        # label i now means the syllable formerly labelled index[i]
        np.testing.assert_array_equal(index[new["states"]["z"]], old["states"]["z"])
        np.testing.assert_allclose(new["params"]["Ab"], np.asarray(old["params"]["Ab"])[index])
        np.testing.assert_allclose(new["params"]["pi"], np.asarray(old["params"]["pi"])[index][:, index])
        np.testing.assert_allclose(new["params"]["betas"], np.asarray(old["params"]["betas"])[index])
    freq = get_frequencies(kio.load_hdf5(path, "model_snapshots/7")["states"]["z"], data["mask"], 5, True)
    assert np.all(np.diff(freq) <= 1e-12)                       # most frequent syllable first
    # an explicit permutation is applied as given
    perm = np.array([4, 3, 2, 1, 0])
    kio.reindex_syllables_in_checkpoint(path=path, index=perm)
    twice = kio.load_hdf5(path, "model_snapshots/7")
    np.testing.assert_array_equal(index[perm[twice["states"]["z"]]], model["states"]["z"])


def test_format_data_reindexes_bodyparts_and_interpolates_missing_points():
    """format_data (util.py:929-1089): keypoints selected / ordered by use_bodyparts, NaN points imputed by
    linear interpolation in time with confidence 0 (+ pseudocount), everything else as before."""
    from keypoint_moseq_b200.util import format_data, interpolate_keypoints, reindex_by_bodyparts
    rng = np.random.default_rng(0)
    T, parts = 50, ["nose", "ear", "tail", "paw"]
    coords = {"a": rng.standard_normal((T, 4, 2)).cumsum(0), "b": rng.standard_normal((T - 7, 4, 2)).cumsum(0)}
    conf = {k: rng.uniform(0.5, 1, v.shape[:2]) for k, v in coords.items()}
    truth = coords["a"].copy()
    coords["a"][10:13, 2] = np.nan                                     # tail missing for 3 frames
    coords["a"][0, 0] = np.nan                                         # nose missing at the start
    use = ["tail", "nose", "paw"]
    data, (keys, bounds) = format_data(coords, conf, bodyparts=parts, use_bodyparts=use, added_noise_level=0.0,
                                       seg_length=30, device="cpu")
    Y, cf, mask = data["Y"].numpy(), data["conf"].numpy(), data["mask"].numpy()
    assert Y.shape[2:] == (3, 2) and list(keys)[:2] == ["a", "a"]
    # reindexing: column 0 is the tail, 1 the nose
    np.testing.assert_allclose(Y[0, 5, 1], truth[5, 0])
    np.testing.assert_allclose(Y[0, 5, 0], truth[5, 2])
    # interpolation between frames 9 and 13 of the tail, confidence pseudocount only
    for t, w in ((10, 0.25), (11, 0.5), (12, 0.75)):
        np.testing.assert_allclose(Y[0, t, 0], (1 - w) * truth[9, 2] + w * truth[13, 2])
        assert abs(cf[0, t, 0] - 1e-3) < 1e-12
    np.testing.assert_allclose(Y[0, 0, 1], truth[1, 0])                # held from the first good frame
    assert np.isfinite(Y).all() and mask.sum() == 50 + 20 + 43 + 13      # windows of 30 + 30 frames of look-ahead
    out = interpolate_keypoints(np.full((4, 1, 2), np.nan), np.ones((4, 1), bool))
    assert (out == 0).all()
    assert reindex_by_bodyparts(np.arange(8.0).reshape(1, 4, 2), parts, ["paw"]).tolist() == [[[6.0, 7.0]]]


def test_snapshot_writer_orders_bounds_and_reports():
    """io.SnapshotWriter: snapshots reach the file in submission order while the caller keeps going, at most
    `max_pending` wait in memory, everything is on disk after close(), and a failed write surfaces in the
    caller's thread (later snapshots are dropped rather than written after a hole)."""
    import threading
    import time
    from keypoint_moseq_b200.io import SnapshotWriter
    written, gate = [], threading.Event()

    def slow_save(path, tree, datapath, exist_ok=False):
        gate.wait(5)
        assert exist_ok is True
        written.append((path, datapath, tree["k"]))

    w = SnapshotWriter(save=slow_save, max_pending=2)
    t0 = time.perf_counter()
    for i in range(3):                                   # one in flight + two queued: none of these block
        w.submit("f.h5", {"k": i}, f"model_snapshots/{i}")
    assert time.perf_counter() - t0 < 1.0 and written == []
    blocked = threading.Thread(target=lambda: w.submit("f.h5", {"k": 3}, "model_snapshots/3"))
    blocked.start()
    blocked.join(0.3)
    assert blocked.is_alive()                            # the fourth waits for room
    gate.set()
    blocked.join(5)
    w.close()
    assert written == [("f.h5", f"model_snapshots/{i}", i) for i in range(4)]

    def failing_save(path, tree, datapath, exist_ok=False):
        if tree["k"] == 1:
            raise OSError("disk full")
        written.append(tree["k"])

    written.clear()
    w = SnapshotWriter(save=failing_save)
    for i in range(3):
        try:
            w.submit("f.h5", {"k": i}, str(i))
        except OSError:
            break
        time.sleep(0.05)
    else:
        raise AssertionError("the write error never reached the caller")
    with pytest.raises(OSError):
        w.submit("f.h5", {"k": 9}, "9")
    with pytest.raises(OSError):
        w.close()
    assert written == [0]


def test_snapshot_writer_writes_a_loadable_checkpoint(tmp_path):
    from keypoint_moseq_b200.io import SnapshotWriter, load_hdf5, save_hdf5
    path = str(tmp_path / "checkpoint.h5")
    save_hdf5(path, {"model_snapshots": {"0": {"x": np.zeros(3)}}, "data": {"mask": np.ones(2)}})
    with SnapshotWriter() as w:
        for it in (25, 50):
            w.submit(path, {"x": np.full(3, float(it))}, f"model_snapshots/{it}")
    back = load_hdf5(path)
    assert sorted(back["model_snapshots"], key=int) == ["0", "25", "50"]
    assert np.array_equal(back["model_snapshots"]["50"]["x"], np.full(3, 50.0))


def test_fit_model_async_checkpoint_failure_reaches_the_caller(tmp_path, monkeypatch):
    """A background write that fails must not be lost: fit_model raises it (at the next snapshot or on return)."""
    from keypoint_moseq_b200 import fitting
    monkeypatch.setattr(fitting.gibbs, "resample_model", lambda data, count=0, **kw: {"count": count + 1, "x": np.zeros(1)})
    monkeypatch.setattr(fitting.gibbs, "to_device_data", lambda d, *a, **k: d)
    monkeypatch.setattr(fitting.gibbs, "to_device_model", lambda m, *a, **k: m)
    monkeypatch.setattr(fitting, "_host_model", lambda m: m)
    monkeypatch.setattr(fitting, "to_numpy_tree", lambda t: t)

    def save(path, tree, datapath=None, **kw):
        if datapath == "model_snapshots/4":
            raise OSError("disk full")

    monkeypatch.setattr(fitting, "save_hdf5", save)
    with pytest.raises(OSError, match="disk full"):
        fitting.fit_model({"count": 0}, {}, ([], []), str(tmp_path), "m", num_iters=6, save_every_n_iters=2,
                          generate_progress_plots=False, async_checkpoints=True)
    with pytest.raises(OSError, match="disk full"):                       # and the synchronous path as before
        fitting.fit_model({"count": 0}, {}, ([], []), str(tmp_path), "m2", num_iters=6, save_every_n_iters=2,
                          generate_progress_plots=False)


def test_bench_reference_arm_prints_one_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs next to the CUDA arm): exactly one JSON line on
    stdout with the contract's keys, the oracle port labelled as such, rows split over the requested processes."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, KPMS_BENCH_CPU_PROCS="2")
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--variant", "ar_only"], capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, out.stdout
    line = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["metric"] == "frame_sweeps_per_sec" and line["unit"] == "frame-sweeps/s"
    assert line["higher_is_better"] is True and line["vs_baseline"] is None and line["dtype"] == "f64"
    assert line["value"] > 0 and "ar_only" in line["config"]["workload"]
    cpu = line["cpu_baseline"]
    assert cpu["kind"] == "port" and cpu["cores"] == 2 and cpu["value"] == line["value"] and "2 processes" in cpu["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # the line reports what was run: one timed sweep of the sample, its own wall time; the extrapolation is separate
    assert line["steps"] == 1 and len(line["step_ms"]) == 1 and abs(line["ms_per_step"] - line["step_ms"][0]) < 1.0
    assert abs(line["value"] - line["config"]["sample_valid_frames"] / (line["ms_per_step"] * 1e-3)) < 1e-6 * line["value"]
    assert line["extrapolated"]["full_workload_ms_per_sweep"] > line["ms_per_step"]


def _sharded_fit_worker(rank, world, port, tmp):
    """Two ranks run the SAME fit_model / apply_model call on the full cohort with group=WORLD; the sweep is a
    stub (no GPU here) that adds one to x on the rows it is given and all-reduces the valid-frame count."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from keypoint_moseq_b200 import fitting, io
    N, T = 5, 6
    mask = np.ones((N, T))
    mask[1, 4:] = 0
    mask[4, 2:] = 0
    keys = ["a", "a", "b", "c", "c"]
    metadata = (keys, np.array([[0, 6], [6, 10], [0, 6], [0, 6], [6, 8]]))
    x0 = np.arange(N, dtype=np.float64)[:, None, None] * np.ones((N, T, 2))

    def fresh():
        return {"seed": 0, "states": {"x": torch.tensor(x0), "z": torch.zeros(N, T - 1, dtype=torch.int64)},
                "params": {"count": torch.zeros(1, dtype=torch.float64)}, "hypparams": {},
                "noise_prior": torch.ones(N, T, 3, dtype=torch.float64)}

    data = {"Y": torch.zeros(N, T, 3, 2, dtype=torch.float64), "mask": torch.tensor(mask), "row": torch.arange(N, dtype=torch.float64)}

    def shard_x(d):       # x = 10 * global row index, so the joined marginals show which row went where
        return (10.0 * d["row"])[:, None, None] * torch.ones(d["mask"].shape[0], T, 2, dtype=torch.float64)

    poison = {"at": None}
    sweeps = {"n": 0}

    def stub_sweep(d, seed, states, params, hypparams, noise_prior, group=None, **kw):
        assert group is not None and states["x"].shape[0] == d["mask"].shape[0] == noise_prior.shape[0] < N
        sweeps["n"] += 1
        total = d["mask"].sum().reshape(1).clone()
        dist.all_reduce(total, group=group)                 # stands for the statistics all-reduce of the real sweep
        x = states["x"] + 1.0
        if poison["at"] == sweeps["n"] and rank == 1:
            x = x.clone()
            x[0, 0, 0] = float("nan")
        return {"seed": seed + 1, "states": dict(states, x=x), "params": dict(params, count=total), "hypparams": hypparams,
                "noise_prior": noise_prior}

    fitting.gibbs.resample_model = stub_sweep
    fitting.gibbs.to_device_data = lambda d, *a, **k: d
    fitting.gibbs.to_device_model = lambda m, *a, **k: m
    out = {}
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model, name = fitting.fit_model(fresh(), data, metadata, tmp, None, num_iters=4, save_every_n_iters=2,
                                        generate_progress_plots=False, group=dist.group.WORLD, device="cpu")
        out["name"] = name
        out["x"] = np.asarray(model["states"]["x"])
        out["count"] = float(np.asarray(model["params"]["count"])[0])
        out["prior_rows"] = int(np.asarray(model["noise_prior"]).shape[0])
        # a NaN on rank 1 alone stops every rank at the same sweep (synchronous check for a deterministic count)
        sweeps["n"], poison["at"] = 0, 3
        model2, _ = fitting.fit_model(fresh(), data, metadata, tmp, "nan_run", num_iters=6, save_every_n_iters=None,
                                      generate_progress_plots=False, group=dist.group.WORLD, device="cpu", nan_check_lag=0)
        out["x_nan"] = np.asarray(model2["states"]["x"])
        out["sweeps_nan"] = sweeps["n"]
        poison["at"] = None
        # apply_model: rows meet again for extract_results, rank 0 alone saves
        fitting.init_model = lambda data=None, **kw: {"seed": 0, "states": {"x": torch.zeros(data["mask"].shape[0], T, 2, dtype=torch.float64),
                                                                           "z": torch.zeros(data["mask"].shape[0], T - 1, dtype=torch.int64),
                                                                           "v": torch.zeros(data["mask"].shape[0], T, 2, dtype=torch.float64),
                                                                           "h": torch.zeros(data["mask"].shape[0], T, dtype=torch.float64)},
                                                      "params": {"count": torch.zeros(1)}, "hypparams": {},
                                                      "noise_prior": torch.ones(data["mask"].shape[0], T, 3, dtype=torch.float64)}
        res = fitting.apply_model({"seed": 0, "params": {}, "hypparams": {}}, data, metadata, tmp, name, num_iters=3,
                                  group=dist.group.WORLD, device="cpu")
        out["results_keys"] = sorted(res)
        out["latent_a"] = np.asarray(res["a"]["latent_state"])
        # estimate_syllable_marginals: local smoother marginals, rows joined before unbatching
        fitting.gibbs.stateseq_marginals = lambda x, mask, Ab, Q, pi: x[:, 1:, :1].repeat(1, 1, 3) * 0 + x[:, 1:, :1]
        fitting.get_nlags = lambda Ab: 1
        fitting.init_model = lambda data=None, **kw: {
            "seed": 0, "states": {"x": shard_x(data), "z": torch.zeros(data["mask"].shape[0], T - 1, dtype=torch.int64)},
            "params": {"Ab": None, "Q": None, "pi": None}, "hypparams": {}, "noise_prior": torch.ones(data["mask"].shape[0], T, 3, dtype=torch.float64)}
        marg, smp = fitting.estimate_syllable_marginals({"seed": 0, "params": {}, "hypparams": {}}, data, metadata,
                                                        burn_in_iters=1, num_samples=2, steps_per_sample=1,
                                                        return_samples=True, group=dist.group.WORLD, device="cpu")
        out["marg_c"] = np.asarray(marg["c"])
        out["smp_a"] = np.asarray(smp["a"]).shape
    np.save(os.path.join(tmp, f"sharded_fit_{rank}.npy"), np.array([out], dtype=object), allow_pickle=True)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_fit_and_apply_write_on_rank_zero_and_return_whole_models(tmp_path):
    """rows are sharded inside fit_model / apply_model, snapshots are gathered and written by rank 0 only,
    every rank returns the whole model, and a NaN seen by one rank stops all of them together."""
    import torch.multiprocessing as mp
    from keypoint_moseq_b200 import io
    port = 31500 + os.getpid() % 2000
    mp.spawn(_sharded_fit_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    outs = [np.load(tmp_path / f"sharded_fit_{r}.npy", allow_pickle=True)[0] for r in range(2)]
    N, T = 5, 6
    x0 = np.arange(N, dtype=np.float64)[:, None, None] * np.ones((N, T, 2))
    assert outs[0]["name"] == outs[1]["name"]                          # the timestamped name is rank 0's
    for o in outs:
        np.testing.assert_array_equal(o["x"], x0 + 5.0)               # sweeps 0..4, every row, original order
        assert o["count"] == 24.0 and o["prior_rows"] == N            # 30 frames - 6 masked, summed over ranks
        np.testing.assert_array_equal(o["x_nan"], x0 + 2.0)           # last clean model, whole, on every rank
        assert o["sweeps_nan"] == 3
        assert o["results_keys"] == ["a", "b", "c"]
        assert o["latent_a"].shape == (10, 2) and np.all(o["latent_a"] == 3.0)
        # recording c = rows 3 and 4 (6 + 2 frames): sweeps add 1 per iteration, samples after sweeps 2 and 3
        # (frame 6 is the first frame of row 4, cut by the nlags shift; real rows overlap their predecessor there)
        np.testing.assert_allclose(o["marg_c"][:6, 0], 30.0 + 2.5)
        np.testing.assert_allclose(o["marg_c"][7, 0], 40.0 + 2.5)
        assert o["marg_c"].shape == (8, 3) and o["smp_a"] == (10, 2)
    ckpt = os.path.join(str(tmp_path), outs[0]["name"], "checkpoint.h5")
    saved = io.load_hdf5(ckpt)
    assert sorted(saved["model_snapshots"], key=int) == ["0", "2", "4"]
    np.testing.assert_array_equal(saved["model_snapshots"]["2"]["states"]["x"], x0 + 3.0)
    np.testing.assert_array_equal(saved["model_snapshots"]["4"]["states"]["x"], x0 + 5.0)
    assert saved["model_snapshots"]["4"]["noise_prior"].shape == (N, T, 3)
    assert saved["data"]["mask"].shape == (N, T)
    results = io.load_results(str(tmp_path), outs[0]["name"])
    assert sorted(results) == ["a", "b", "c"] and results["c"]["latent_state"].shape == (8, 2)


def test_batch_unbatch_and_sharding_properties_hold_for_ragged_cohorts():
    """Property test over ragged cohorts (empty-ish, tiny and long recordings, any segment length): unbatch
    inverts batch exactly, the mask counts every frame once plus the overlap, and shard_rows is a partition whose
    shards reassemble (as gather_rows does) to the original rows."""
    from hypothesis import given, settings, strategies as st
    from keypoint_moseq_b200.dist import shard_rows, shard_tree
    from keypoint_moseq_b200.util import batch, unbatch

    @settings(max_examples=60, deadline=None)
    @given(lengths=st.lists(st.integers(min_value=1, max_value=90), min_size=1, max_size=6),
           seg=st.integers(min_value=5, max_value=60), overlap=st.integers(min_value=0, max_value=8),
           world=st.integers(min_value=1, max_value=4), seed=st.integers(min_value=0, max_value=10 ** 6))
    def check(lengths, seg, overlap, world, seed):
        rng = np.random.default_rng(seed)
        recs = {f"r{i}": rng.standard_normal((n, 2)) for i, n in enumerate(lengths)}
        stack, mask, (keys, bounds) = batch(recs, seg_length=seg, seg_overlap=overlap)
        assert stack.shape[:2] == mask.shape and stack.shape[1] == seg + overlap
        back = unbatch(stack, keys, bounds)
        assert set(back) == set(recs)
        for name, arr in recs.items():
            np.testing.assert_array_equal(back[name], arr)
        bounds = np.asarray(bounds)
        np.testing.assert_array_equal(mask.sum(1), bounds[:, 1] - bounds[:, 0])      # valid frames per row
        shards = shard_rows(mask, world, keys)
        flat = np.sort(np.concatenate(shards))
        np.testing.assert_array_equal(flat, np.arange(mask.shape[0]))                # a partition of the rows
        joined = np.empty_like(stack)
        for rows in shards:
            if len(rows):
                joined[rows] = shard_tree({"Y": stack}, rows)["Y"]
        np.testing.assert_array_equal(joined, stack)

    check()


def test_npz_checkpoint_snapshots_are_appended_not_rewritten(tmp_path, monkeypatch):
    """Without h5py the checkpoint is a NumPy zip archive with the same group layout: a new snapshot is appended
    as new members (cost = its own size, as with h5py's append mode, keypoint_moseq/io.py:1297-1329), collisions
    are refused unless `overwrite`, deletes / overwrites fall back to a rewrite, and readers see one tree."""
    from keypoint_moseq_b200 import io as kio
    monkeypatch.setattr(kio, "HAVE_H5PY", False)
    f = str(tmp_path / "checkpoint.h5")
    snap = lambda it: {"states": {"x": np.full((3, 5, 2), float(it)), "z": np.arange(6).reshape(3, 2) + it},
                       "params": {"pi": np.eye(2) * it}, "seed": np.array([1, it], np.uint32), "empty": {}}
    kio.save_hdf5(f, {"model_snapshots": {"0": snap(0)}, "data": {"Y": np.ones((3, 5, 4, 2))},
                      "metadata": (np.array(["a", "b"]), np.array([[0, 1], [2, 3]]))})
    rewrites = []
    real_write = kio._npz_write
    monkeypatch.setattr(kio, "_npz_write", lambda *a, **k: (rewrites.append(1), real_write(*a, **k))[1])
    for it in (5, 10):
        kio.save_hdf5(f, snap(it), f"model_snapshots/{it}", exist_ok=True)
    assert rewrites == []                                        # appended
    ck = kio.load_hdf5(f)
    assert sorted(ck["model_snapshots"], key=int) == ["0", "5", "10"]
    assert isinstance(ck["metadata"], tuple) and list(ck["metadata"][0]) == ["a", "b"]
    for it in (0, 5, 10):
        got = kio.load_hdf5(f, f"model_snapshots/{it}")
        assert got["empty"] == {} and got["seed"].dtype == np.uint32
        np.testing.assert_array_equal(got["states"]["x"], snap(it)["states"]["x"])
        np.testing.assert_array_equal(got["params"]["pi"], np.eye(2) * it)
        np.testing.assert_array_equal(ck["model_snapshots"][str(it)]["states"]["z"], snap(it)["states"]["z"])
    with pytest.raises(AssertionError, match="already exists"):
        kio.save_hdf5(f, snap(11), "model_snapshots/10", exist_ok=True)
    kio.save_hdf5(f, snap(11), "model_snapshots/10", exist_ok=True, overwrite=True)     # rewrite path
    assert rewrites == [1]
    np.testing.assert_array_equal(kio.load_hdf5(f, "model_snapshots/10")["params"]["pi"], np.eye(2) * 11)
    kio.delete_snapshots_after(f, 5)
    assert sorted(kio.load_hdf5(f)["model_snapshots"], key=int) == ["0", "5"]
    kio.save_hdf5(f, snap(7), "model_snapshots/7", exist_ok=True)
    assert sorted(kio.load_hdf5(f)["model_snapshots"], key=int) == ["0", "5", "7"]


def test_async_host_copy_passes_host_trees_through():
    """util.AsyncHostCopy without a device: host leaves pass through as NumPy arrays, the post-processing hook runs
    on the settled tree, non-tensor leaves are untouched (the CUDA path is covered by
    tests/test_gpu_fitting.py::test_async_checkpoints_write_the_same_snapshots)."""
    import torch
    from keypoint_moseq_b200.util import AsyncHostCopy
    tree = {"states": {"x": torch.arange(6.0).reshape(2, 3), "z": torch.arange(4, dtype=torch.int32)},
            "params": {"pi": np.eye(2)}, "seed": np.array([1, 2], np.uint32), "hypparams": {"kappa": 1e4, "name": "a"},
            "list": [torch.ones(2), 3]}
    seen = []

    def post(host):
        seen.append(True)
        host["states"]["z"] = np.asarray(host["states"]["z"]).astype(np.int64)
        return host

    out = AsyncHostCopy(tree, post).result()
    assert seen == [True]
    assert isinstance(out["states"]["x"], np.ndarray) and out["states"]["z"].dtype == np.int64
    np.testing.assert_array_equal(out["states"]["x"], np.arange(6.0).reshape(2, 3))
    np.testing.assert_array_equal(out["params"]["pi"], np.eye(2))
    assert out["hypparams"] == {"kappa": 1e4, "name": "a"} and out["list"][1] == 3
    np.testing.assert_array_equal(out["list"][0], np.ones(2))
