"""Behaviour of the verified time chunks away from convergence: fit the C2 cohort from parameters that do
NOT generate it (AR parameters and transitions redrawn from the prior, states re-initialised from the
data) and log, per sweep, the sweep time and the chunk diagnostics (chains re-run sequentially, boundary
mismatches).  Prints one JSON line per sweep.  `python tools/cold_start.py --sweeps 30`"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from keypoint_moseq_b200 import fitting, gibbs  # noqa: E402
from keypoint_moseq_b200.synth import CONFIGS, sample_dataset  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="C2")
ap.add_argument("--sweeps", type=int, default=30)
ap.add_argument("--ar-only-sweeps", type=int, default=0)
a = ap.parse_args()
cfg = dict(CONFIGS[a.config])
kw = dict(recordings=cfg["recordings"], frames=cfg["frames"], k=cfg["k"], D=cfg["D"], d=cfg["d"], L=cfg["L"], K=cfg["K"])
data, _, truth = sample_dataset(seed=1000, kappa=1e4, **kw)
_, _, other = sample_dataset(**dict(kw, recordings=1, frames=64), seed=4321, kappa=1e4)   # unrelated prior draw
params = dict(truth["params"], Ab=other["params"]["Ab"], Q=other["params"]["Q"], pi=other["params"]["pi"],
              betas=other["params"]["betas"])
model = fitting.init_model(data=data, params=params, hypparams=truth["hypparams"], seed=np.array([0, 7], dtype=np.uint32),
                           noise_prior=truth["noise_prior"], dtype=torch.float32)
dd = gibbs.to_device_data(data, "cuda", torch.float32)
m = gibbs.to_device_model(model, "cuda", torch.float32)
for it in range(a.sweeps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    m = gibbs.resample_model(dd, **m, ar_only=it < a.ar_only_sweeps)
    e1.record()
    torch.cuda.synchronize()
    z = m["states"]["z"]
    changes = (z[:, 1:] != z[:, :-1]).float().mean().item()
    print(json.dumps({"sweep": it, "ms": round(e0.elapsed_time(e1), 2), "mean_run": round(1.0 / max(changes, 1e-9), 1),
                      "states_used": int(torch.unique(z).numel()),
                      "finite": bool(torch.isfinite(m["states"]["x"]).all()),
                      "kalman": gibbs.chunk_diagnostics("kalman_ws"), "hmm": gibbs.chunk_diagnostics("hmm_ws")}), flush=True)
