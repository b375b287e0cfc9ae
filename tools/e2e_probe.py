"""Why is the end-to-end step (host operands in, host states out) slower at N > 1?  Run under torchrun:

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/e2e_probe.py

Per rank, C2-sized shard, float32: (A) host-operand sweep without a process group, (B) with it (statistics all-reduce
in the eager path), (C) device-resident eager sweep with the group, (D) the bare pinned-host <-> device copies of one
step, (E) like B with the NCCL all-reduce on a tiny tensor only.  Each: mean ms over 5 steps after one warm-up."""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from keypoint_moseq_b200 import gibbs  # noqa: E402
from keypoint_moseq_b200.synth import CONFIGS, sample_dataset  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
group = dist.group.WORLD if world > 1 else None
data, _, model = sample_dataset(seed=1000, data_seed=rank, kappa=1e4, **CONFIGS["C2"])
dd = gibbs.to_device_data(data, "cuda", torch.float32)
m = gibbs.to_device_model(model, "cuda", torch.float32)
for _ in range(3):
    m = gibbs.resample_model(dd, **m, group=group)
torch.cuda.synchronize()
pin = lambda t: t.cpu().pin_memory()  # noqa: E731
hd = {k: pin(v) for k, v in dd.items()}
hs = {k: pin(v) for k, v in m["states"].items()}
hp = {k: pin(v) for k, v in m["params"].items()}
hprior = pin(m["noise_prior"])
out_host = {k: torch.empty_like(v).pin_memory() for k, v in hs.items()}


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return round(1e3 * (time.perf_counter() - t0) / reps, 2)


def host_step(grp):
    def f():
        mm = {"seed": m["seed"], "states": hs, "params": hp, "hypparams": m["hypparams"], "noise_prior": hprior}
        gibbs.resample_model(hd, **mm, host_out=out_host, group=grp)
        torch.cuda.synchronize()
    return f


def dev_eager(grp):
    def f():
        gibbs.resample_model(dd, **m, group=grp, graph=False)
        torch.cuda.synchronize()
    return f


dev_bufs = {k: torch.empty_like(v, device="cuda") for k, v in list(hd.items()) + list(hs.items())}


def copies():
    for k, v in list(hd.items()) + list(hs.items()):
        dev_bufs[k].copy_(v, non_blocking=True)
    for k, v in out_host.items():
        v.copy_(dev_bufs[k], non_blocking=True)
    torch.cuda.synchronize()


res = {"rank": rank, "world": world}
res["A_host_nogroup"] = timed(host_step(None))
if world > 1:
    res["B_host_group"] = timed(host_step(group))
    res["C_device_eager_group"] = timed(dev_eager(group))
res["C0_device_eager_nogroup"] = timed(dev_eager(None))
res["D_copies_only"] = timed(copies)
nbytes = sum(v.numel() * v.element_size() for v in list(hd.values()) + list(hs.values()) + list(out_host.values()))
res["D_gb_per_s"] = round(nbytes / res["D_copies_only"] / 1e6, 1)
if world > 1:
    tiny = torch.zeros(16, device="cuda", dtype=torch.float64)
    res["E_allreduce_tiny_ms"] = timed(lambda: (dist.all_reduce(tiny), torch.cuda.synchronize()))
    packed = torch.zeros(150000, device="cuda", dtype=torch.float64)
    res["E_allreduce_1MB_ms"] = timed(lambda: (dist.all_reduce(packed), torch.cuda.synchronize()))
print(json.dumps(res), flush=True)
if world > 1:
    gibbs.release_graphs()
    dist.destroy_process_group()
