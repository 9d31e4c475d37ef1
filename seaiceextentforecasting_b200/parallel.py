"""Multi-GPU glue: one process per GPU, `torch.distributed` (NCCL on the box, gloo in the CPU tests).

Two ways the path shards, both without touching `R`:
  * independent (config, year) tasks: `SweepPlan(rank, world)` owns whole tasks, ranks exchange nothing during the
    step; `gather_results` all-gathers the GP records (80-byte structs, tensor collective) at the end;
  * one large network (25 km grid): `sie_corr_tau(shard_rank, shard_count)` computes the 128-row tile rows
    `bi % shard_count == shard_rank` of the upper triangle, so every rank holds a partial (sum, count) of the
    significant correlations; `tau_from_shards` all-reduces those 16 bytes per network.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

TILE_ROWS = 128     # csrc/corr.cu TILE


def shard_tile_rows(n_nodes, rank, count):
    """Tile rows (128 nodes each) of the upper triangle computed by `rank` of `count` (csrc/corr.cu shard_rows)."""
    nb = (int(n_nodes) + TILE_ROWS - 1) // TILE_ROWS
    return list(range(rank, nb, count))


def tau_from_shards(tau_sum, tau_cnt, group=None):
    """tau_sum (float64 [B]), tau_cnt (int64 [B]): this rank's partials -> the global tau [B] on every rank
    (ComplexNetworks.py:45-47: mean of the significant non-negative correlations over both triangles)."""
    s = tau_sum.clone()
    c = tau_cnt.clone()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(s, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(c, op=dist.ReduceOp.SUM, group=group)
    return s / c.to(torch.float64)


def gather_records(plan, raw, group=None, device=None, raw_dev=None):
    """Tensor all-gather of every rank's GP records (80 B each) and their (config, region, year, member) keys: ranks
    own different numbers of problems, so the counts are gathered first and the payloads padded to the largest.
    Returns (raw_all, meta_all [P,3], member_all [P]) in rank order.  `device`: where the collective runs ("cuda" under
    NCCL, "cpu" under gloo); default = CUDA when the backend is NCCL.  `raw_dev`: the records as a uint8 device tensor
    (GpBatch.out) instead of `raw` -- the payload then goes GPU -> NVLink -> GPU without a host round trip."""
    from .forecast import GP_RESULT_DTYPE
    if raw is None:
        raw = np.zeros(plan.P, dtype=GP_RESULT_DTYPE) if raw_dev is not None else np.zeros(0, dtype=GP_RESULT_DTYPE)
        if raw_dev is not None and not (dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1):
            raw = raw_dev.cpu().numpy().view(GP_RESULT_DTYPE)[:plan.P]
    raw = np.ascontiguousarray(raw)
    meta = np.asarray(plan.prob_meta, dtype=np.int32).reshape(-1, 3)
    keys = np.concatenate([meta, np.asarray(plan.prob_member, dtype=np.int32).reshape(-1, 1)], axis=1)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return raw, keys[:, :3], keys[:, 3]
    world = dist.get_world_size(group)
    if device is None:
        device = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    isz = raw.dtype.itemsize
    n = torch.tensor([len(raw)], dtype=torch.int64, device=device)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n, group=group)
    counts = [int(c.item()) for c in counts]
    cap = max(max(counts), 1)
    pay = torch.zeros(cap * isz, dtype=torch.uint8, device=device)
    if raw_dev is not None:
        pay[:len(raw) * isz] = raw_dev[:len(raw) * isz].to(device)
    else:
        pay[:len(raw) * isz] = torch.from_numpy(raw.view(np.uint8).reshape(-1).copy()).to(device)
    key = torch.zeros((cap, 4), dtype=torch.int32, device=device)
    key[:len(raw)] = torch.from_numpy(np.ascontiguousarray(keys)).to(device)
    pays = [torch.empty_like(pay) for _ in range(world)]
    keyl = [torch.empty_like(key) for _ in range(world)]
    dist.all_gather(pays, pay, group=group)
    dist.all_gather(keyl, key, group=group)
    raw_all = np.concatenate([p.cpu().numpy()[:c * isz].view(raw.dtype) for p, c in zip(pays, counts)])
    key_all = np.concatenate([k.cpu().numpy()[:c] for k, c in zip(keyl, counts)])
    return raw_all, key_all[:, :3], key_all[:, 3]


def gather_results(plan, raw, group=None, device=None):
    """Every rank's GP records -> the assembled GPR dict of the whole sweep (SweepPlan.assemble) on every rank."""
    raw_all, meta, member = gather_records(plan, raw, group, device)
    return plan.assemble(raw_all, meta, member)
