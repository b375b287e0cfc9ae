"""Sharding of chains across ranks and the one per-sweep exchange.

The reference's only multi-device mechanism maps independent sequences over GPUs
(`jax_moseq.utils.set_mixed_map_gpus`, docs/source/FAQs.rst:489-506) and has no collective.
Here every rank owns a fixed subset of the (N, T) rows; the per-chain samplers never
communicate, and the packed float64 sufficient statistics (AR Gram matrices, transition
counts, optional observation-variance sums) are summed with ONE all-reduce per sweep
(NCCL over NVLink on GPUs; gloo in the CPU tests).  All ranks then draw identical
parameters from identical statistics and the same Philox key, so no broadcast follows.
"""
import numpy as np
import torch

__all__ = ["shard_rows", "shard_tree", "allreduce_statistics", "gather_rows"]


def shard_rows(mask, world_size, keys=None):
    """Greedy balanced assignment of rows to ranks by valid-frame count.

    Rows of one recording (equal `keys`) are kept on one rank when that does not unbalance the
    load by more than one row; otherwise (few long recordings, many ranks) single rows are dealt out.
    Returns a list of index arrays, one per rank, each sorted.
    """
    mask = np.asarray(mask.cpu() if isinstance(mask, torch.Tensor) else mask)
    load = mask.sum(1).astype(np.int64)
    N = mask.shape[0]
    if keys is None:
        groups = [[i] for i in range(N)]
    else:
        order = {}
        for i, key in enumerate(keys):
            order.setdefault(key, []).append(i)
        groups = list(order.values())

    def deal(groups):
        groups = sorted(groups, key=lambda g: (-int(load[g].sum()), g[0]))
        totals = np.zeros(world_size, dtype=np.int64)
        out = [[] for _ in range(world_size)]
        for g in groups:
            r = int(np.argmin(totals))
            out[r].extend(g)
            totals[r] += int(load[g].sum())
        return out, totals

    out, totals = deal(groups)
    if keys is not None and N and int(totals.max() - totals.min()) > int(load.max()):
        # few long recordings on many ranks: whole recordings cannot be balanced, deal single rows instead
        out, totals = deal([[i] for i in range(N)])
    return [np.array(sorted(rows), dtype=np.int64) for rows in out]


def shard_tree(tree, rows):
    """Select `rows` along the leading axis of every (N, ...) leaf of a data / states dict."""
    if isinstance(tree, dict):
        return {k: shard_tree(v, rows) for k, v in tree.items()}
    if isinstance(tree, torch.Tensor):
        return tree[torch.as_tensor(rows, device=tree.device)]
    return np.asarray(tree)[rows]


def allreduce_statistics(packed, group=None):
    """In-place SUM all-reduce of the packed statistics buffer (a no-op without a process group)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
    return packed


def gather_rows(local, rows_per_rank, group=None):
    """Reassemble an (N, ...) array on every rank from per-rank row blocks (checkpoint time only)."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    parts = [None] * world
    dist.all_gather_object(parts, local.cpu().numpy() if isinstance(local, torch.Tensor) else local, group=group)
    N = sum(len(r) for r in rows_per_rank)
    out = np.empty((N,) + parts[0].shape[1:], dtype=parts[0].dtype)
    for rows, part in zip(rows_per_rank, parts):
        out[rows] = part
    return out
