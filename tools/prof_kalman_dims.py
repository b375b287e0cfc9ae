"""Per-kernel times of the Kalman FFBS at given latent dimensions (float32, 80 chains x 10 000 frames):
  python tools/prof_kalman_dims.py 10 12 16"""
import sys, torch, json
sys.path.insert(0, "/root/repo")
from keypoint_moseq_b200 import _lib, gibbs
from keypoint_moseq_b200.synth import sample_dataset
dims = [int(a) for a in sys.argv[1:]] or [16, 4]
for d in dims:
    data, _, model = sample_dataset(recordings=80, frames=10000, k=12, D=2, d=d, L=3, K=100, seed=5, seg_length=10000, max_seg_length=10000)
    dd = gibbs.to_device_data(data, "cuda", torch.float32); m = gibbs.to_device_model(model, "cuda", torch.float32)
    st, pr = m["states"], m["params"]
    f = lambda: gibbs.resample_continuous_stateseqs(dd["Y"], dd["mask"], st["v"], st["h"], st["s"], st["z"], pr["Cd"], pr["sigmasq"], pr["Ab"], pr["Q"], 1e-3, 11)
    f(); torch.cuda.synchronize()
    _lib.profile(True); f(); rep = _lib.profile_report(); _lib.profile(False)
    print(json.dumps({"d": d, "kernels_ms": {k: round(v[0], 3) for k, v in sorted(rep.items(), key=lambda kv: -kv[1][0])}, "diag": gibbs.chunk_diagnostics("kalman_ws")}), flush=True)
    del dd, m, st, pr
    gibbs._SCRATCH.clear()
    torch.cuda.empty_cache()
