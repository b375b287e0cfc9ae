"""BASELINE config 5: the two FFBS samplers in isolation over chain length, state count and latent
dimension, time-chunked (default) against the sequential recursion (chunks = 1).

  python tools/kernel_sweep.py [--quick] > gpurun_out/kernel_sweep.jsonl

One JSON line per (sampler, T, K, d, mode): median ms over `--reps` calls after one warm-up, frames/s and
the chunk diagnostics.  Inputs are drawn from the generative model (synth.sample_dataset), N chosen so that
N*T is about `--frames` frames.  `--new` runs only the rows added in round 2 (num_states 250 / 500 on the wide-state
kernels, latent_dim 6 / 8 / 12)."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from keypoint_moseq_b200 import _lib, gibbs  # noqa: E402
from keypoint_moseq_b200.synth import sample_dataset  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=800_000)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--quick", action="store_true")
ap.add_argument("--new", action="store_true")
ap.add_argument("--dims", default="", help="comma-separated latent dimensions: only those rows (K = 100, T = 10 000)")
a = ap.parse_args()

base = dict(T=10_000, K=100, d=10)
grid = [dict(base, T=T) for T in (1_000, 10_000, 100_000, 1_000_000)]
grid += [dict(base, K=K) for K in (25, 50, 128, 500)]
grid += [dict(base, d=d) for d in (4, 16)]
if a.quick:
    grid = [dict(base, T=1_000), base, dict(base, K=25), dict(base, d=4)]
if a.new:
    grid = [dict(base, K=K) for K in (250, 500)] + [dict(base, d=d) for d in (6, 8, 12)]
if a.dims:
    grid = [dict(base, d=int(d)) for d in a.dims.split(",")]


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    out = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        out.append(e0.elapsed_time(e1))
    return float(np.median(out))


for g in grid:
    T, K, d = g["T"], g["K"], g["d"]
    N = max(1, a.frames // T)
    t0 = time.time()
    data, _, model = sample_dataset(recordings=N, frames=T, k=12, D=2, d=d, L=3, K=K, seed=5, seg_length=T,
                                    max_seg_length=T)
    gen_s = time.time() - t0
    dd = gibbs.to_device_data(data, "cuda", torch.float32)
    m = gibbs.to_device_model(model, "cuda", torch.float32)
    st, pr = m["states"], m["params"]
    frames = int(data["mask"].sum())
    for mode, chunks in (("chunked", 0), ("sequential", 1)):
        _lib.set_time_chunking(chunks=chunks)
        ms = timed(lambda: gibbs.resample_discrete_stateseqs(st["x"], dd["mask"], pr["Ab"], pr["Q"], pr["pi"], 11), a.reps)
        print(json.dumps(dict(g, N=N, sampler="hmm_ffbs", mode=mode, ms=round(ms, 3), frames_per_s=round(frames / ms * 1e3),
                              diag=gibbs.chunk_diagnostics("hmm_ws"), gen_s=round(gen_s, 1))), flush=True)
        ms = timed(lambda: gibbs.resample_continuous_stateseqs(dd["Y"], dd["mask"], st["v"], st["h"], st["s"], st["z"],
                                                               pr["Cd"], pr["sigmasq"], pr["Ab"], pr["Q"], 1e-3, 11), a.reps)
        print(json.dumps(dict(g, N=N, sampler="kalman_ffbs", mode=mode, ms=round(ms, 3), frames_per_s=round(frames / ms * 1e3),
                              diag=gibbs.chunk_diagnostics("kalman_ws"))), flush=True)
    _lib.set_time_chunking(chunks=0)
    del dd, m, st, pr
    gibbs._SCRATCH.clear()
    torch.cuda.empty_cache()
