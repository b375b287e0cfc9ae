"""torchrun entry: a sweep sharded over WORLD_SIZE GPUs draws the same parameters as the unsharded
sweep (statistics all-reduced over NCCL), and each rank's chains match the single-GPU result when fed
the same per-chain draws.  Prints 'dist_check ok' on rank 0."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle as orc  # noqa: E402  (checker only)
from keypoint_moseq_b200 import gibbs  # noqa: E402
from keypoint_moseq_b200.dist import shard_rows, shard_tree  # noqa: E402
from keypoint_moseq_b200.synth import sample_dataset  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
import datetime  # noqa: E402
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])),
                        timeout=datetime.timedelta(seconds=120))
data, meta, model = sample_dataset(recordings=4, frames=400, k=5, D=2, d=4, L=3, K=12, seed=7, seg_length=250, kappa=1e2)
N, T, k, D = data["Y"].shape
tape = orc.make_tape(np.random.default_rng(3), N, T, k, D, 4, 3, 12)
rows = shard_rows(data["mask"], world, meta[0])[rank]
per_chain = ("u_z", "w_x", "g_s", "u_h", "w_v")
tape_loc = {key: (val[rows] if key in per_chain else val) for key, val in tape.items()}
dd = gibbs.to_device_data(shard_tree(data, rows), "cuda", torch.float64)
dm = gibbs.to_device_model(dict(model, states=shard_tree(model["states"], rows),
                                noise_prior=model["noise_prior"][rows]), "cuda", torch.float64)
out = gibbs.resample_model(dd, **dm, draws=tape_loc, group=dist.group.WORLD)
st, pr, _ = orc.resample_model(data, model["states"], model["params"], model["hypparams"], model["noise_prior"], tape)
ok = True
for key in ("Ab", "Q", "betas", "pi"):
    err = np.abs(out["params"][key].cpu().numpy() - pr[key]).max() / np.abs(pr[key]).max()
    ok &= err < 1e-8
ok &= np.array_equal(out["states"]["z"].cpu().numpy(), st["z"][rows])
ok &= np.abs(out["states"]["x"].cpu().numpy() - st["x"][rows]).max() < 1e-7 * np.abs(st["x"]).max()
# every rank holds bit-identical parameters
ab = out["params"]["Ab"].clone()
ref = ab.clone()
dist.broadcast(ref, 0)
ok &= bool(torch.equal(ab, ref))
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("dist_check ok" if flag.item() == 1 else "dist_check FAILED")
gibbs.release_graphs()
dist.destroy_process_group()
sys.exit(0 if flag.item() == 1 else 1)
