"""what a checkpoint costs the sweep loop with and without `async_checkpoints`.

  python tools/checkpoint_overlap.py [--config C2] [--iters 30] [--every 5] > gpurun_out/checkpoint_overlap.jsonl

Runs `fit_model` (float32 states, device-resident, graphs on) on a synthetic cohort three times - no checkpoints,
blocking checkpoints (the reference's behaviour, keypoint_moseq/fitting.py:266-275), background checkpoints
(util.AsyncHostCopy + io.SnapshotWriter) - and prints one JSON line per mode: wall seconds for the loop, per sweep,
and the extra time per snapshot over the no-checkpoint run.  Wall clock on purpose: the cost being measured is host
time during which no sweep is queued."""
import argparse
import json
import os
import sys
import tempfile
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from keypoint_moseq_b200 import fitting  # noqa: E402
from keypoint_moseq_b200.synth import sample_dataset  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="C2", choices=["C1", "C2"])
ap.add_argument("--iters", type=int, default=30)
ap.add_argument("--every", type=int, default=5)
a = ap.parse_args()
shape = dict(C1=dict(recordings=4, frames=10_000, k=10, D=2, d=4, L=3, K=100),
             C2=dict(recordings=20, frames=36_000, k=12, D=2, d=10, L=3, K=100))[a.config]
data, meta, model = sample_dataset(seed=7, **shape)
tmp = tempfile.mkdtemp(prefix="kpms_ckpt_")
# one short fit first: builds the graphs, pins nothing yet, writes nothing
fitting.fit_model(model, data, meta, num_iters=3, save_every_n_iters=None, dtype=torch.float32)
base = None
for mode, kw in (("none", dict(save_every_n_iters=None)),
                 ("blocking", dict(save_every_n_iters=a.every, async_checkpoints=False)),
                 ("background", dict(save_every_n_iters=a.every, async_checkpoints=True)),
                 ("background_again", dict(save_every_n_iters=a.every, async_checkpoints=True))):   # pinned pool warm
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out, _ = fitting.fit_model(model, data, meta, tmp, mode, num_iters=a.iters, generate_progress_plots=False,
                               dtype=torch.float32, **kw)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    snaps = 0 if kw["save_every_n_iters"] is None else a.iters // a.every
    if mode == "none":
        base = wall
    line = dict(mode=mode, config=a.config, sweeps=a.iters + 1, snapshots=snaps, wall_s=round(wall, 3),
                ms_per_sweep=round(1e3 * wall / (a.iters + 1), 2),
                extra_ms_per_snapshot=None if not snaps else round(1e3 * (wall - base) / snaps, 1))
    size = os.path.join(tmp, mode, "checkpoint.h5")
    if os.path.exists(size):
        line["checkpoint_mb"] = round(os.path.getsize(size) / 2**20, 1)
    print(json.dumps(line), flush=True)
