import sys, json, torch
sys.path.insert(0, "/root/repo")
from keypoint_moseq_b200 import _lib, gibbs
from keypoint_moseq_b200.synth import CONFIGS, sample_dataset
cfg = CONFIGS["C2"]
data, _, model = sample_dataset(seed=1000, kappa=1e4, **cfg)
dd = gibbs.to_device_data(data, "cuda", torch.float32); m = gibbs.to_device_model(model, "cuda", torch.float32)
for _ in range(3): m = gibbs.resample_model(dd, **m)
for W in (64, 48, 32, 16):
    _lib.set_time_chunking(warmup=W)
    gibbs._SCRATCH.clear()
    mm = m
    for _ in range(2): mm = gibbs.resample_model(dd, **mm)
    torch.cuda.synchronize()
    _lib.profile(True)
    for _ in range(3): mm = gibbs.resample_model(dd, **mm)
    rep = _lib.profile_report(); _lib.profile(False)
    print(W, {k: round(v[0] / 3, 3) for k, v in rep.items() if k in ("kalman_forward", "kalman_affine", "hmm_forward", "hmm_backward", "kalman_forward_rerun", "hmm_forward_rerun")},
          gibbs.chunk_diagnostics("kalman_ws"), gibbs.chunk_diagnostics("hmm_ws"))
