"""torchrun entry: `fit_model` / `apply_model` called identically on every rank with group=WORLD (NCCL).
Checks that the run completes, that rank 0's checkpoint holds every row at the reference's cadence, that all ranks
return the same whole model, and that `apply_model` returns per-recording results on every rank.
Prints 'dist_fit_check ok' on rank 0.  (Host logic is covered on the CPU by tests/test_host.py with a stub sweep;
this is the same path with the real kernels.)"""
import datetime
import os
import sys
import tempfile

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from keypoint_moseq_b200 import fitting, gibbs, io  # noqa: E402
from keypoint_moseq_b200.synth import sample_dataset  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])),
                        timeout=datetime.timedelta(seconds=120))
data, meta, model = sample_dataset(recordings=5, frames=400, k=5, D=2, d=4, L=3, K=12, seed=7, seg_length=250, kappa=1e2)
N = data["Y"].shape[0]
box = [tempfile.mkdtemp() if rank == 0 else None]
dist.broadcast_object_list(box, src=0)
project = box[0]
ok = True
fit, name = fitting.fit_model(model, data, meta, project, None, num_iters=6, save_every_n_iters=3,
                              generate_progress_plots=False, group=dist.group.WORLD)
for key in ("x", "z", "s", "h", "v"):
    ok &= int(fit["states"][key].shape[0]) == N                      # whole model on every rank
    ok &= bool(torch.isfinite(fit["states"][key].double()).all())
ref = fit["states"]["x"].clone()
dist.broadcast(ref, 0)
ok &= bool(torch.equal(ref, fit["states"]["x"]))                      # ... and the same one
ab = fit["params"]["Ab"].clone()
dist.broadcast(ab, 0)
ok &= bool(torch.equal(ab, fit["params"]["Ab"]))
dist.barrier()
if rank == 0:
    saved = io.load_hdf5(os.path.join(project, name, "checkpoint.h5"))
    ok &= sorted(saved["model_snapshots"], key=int) == ["0", "3", "6"]
    ok &= saved["model_snapshots"]["6"]["states"]["x"].shape[0] == N
    ok &= bool(np.array_equal(saved["model_snapshots"]["6"]["states"]["z"], fit["states"]["z"].cpu().numpy()))
res = fitting.apply_model(fit, data, meta, project, name, num_iters=3, group=dist.group.WORLD)
ok &= sorted(res) == sorted(set(meta[0]))
ok &= all(np.isfinite(r["latent_state"]).all() for r in res.values())
dist.barrier()
if rank == 0:
    ok &= sorted(io.load_results(project, name)) == sorted(set(meta[0]))
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("dist_fit_check ok" if flag.item() == 1 else "dist_fit_check FAILED")
gibbs.release_graphs()
dist.destroy_process_group()
sys.exit(0 if flag.item() == 1 else 1)
