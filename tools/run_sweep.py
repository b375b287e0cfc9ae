"""Runs a few full sweeps on a small cohort (for ncu captures: `ncu ... python tools/run_sweep.py`)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from keypoint_moseq_b200 import gibbs  # noqa: E402
from keypoint_moseq_b200.synth import sample_dataset  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--recordings", type=int, default=16)
ap.add_argument("--frames", type=int, default=3000)
ap.add_argument("--sweeps", type=int, default=2)
ap.add_argument("--dtype", default="float32")
ap.add_argument("--k", type=int, default=12)
ap.add_argument("--D", type=int, default=2)
ap.add_argument("--d", type=int, default=10)
ap.add_argument("--K", type=int, default=100)
a = ap.parse_args()
dt = torch.float32 if a.dtype == "float32" else torch.float64
data, _, model = sample_dataset(recordings=a.recordings, frames=a.frames, k=a.k, D=a.D, d=a.d, L=3, K=a.K, seed=0,
                                seg_length=a.frames)
dd, m = gibbs.to_device_data(data, "cuda", dt), gibbs.to_device_model(model, "cuda", dt)
for _ in range(a.sweeps):
    m = gibbs.resample_model(dd, **m)
torch.cuda.synchronize()
print("ok", float(m["states"]["x"].abs().mean()))
