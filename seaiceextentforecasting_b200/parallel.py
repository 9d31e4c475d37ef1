"""Multi-GPU glue: one process per GPU, `torch.distributed` (NCCL on the box, gloo in the CPU tests).

Two ways the path shards, both without touching `R`:
  * independent (config, year) tasks: `SweepPlan(rank, world)` owns whole tasks, ranks exchange nothing during the
    step; `gather_results` collects the GP records at the end (a few KB);
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


def gather_results(plan, raw, group=None):
    """Every rank's GP records -> the assembled GPR dict of the whole sweep (SweepPlan.assemble) on every rank."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return plan.assemble(raw)
    world = dist.get_world_size(group)
    gathered = [None] * world
    dist.all_gather_object(gathered, (plan.prob_meta, np.ascontiguousarray(raw).tobytes()), group=group)
    meta = [m for g in gathered for m in g[0]]
    allraw = np.concatenate([np.frombuffer(g[1], dtype=raw.dtype) for g in gathered])
    return plan.assemble(allraw, meta)
