"""Multi-GPU sharding of the hot path (one process per GPU, ``torch.distributed``).

States and edges are independent given the (replicated) model, so a batch is cut into
contiguous per-rank blocks and every rank validates its block with its own engine -- there is no
collective in the data path.  The one exchange step is assembling a full edge-validity mask on
every rank for PRM roadmap construction (BASELINE.json config 4): one ``all_gather_into_tensor``
of the per-rank masks (NCCL over NVLink on the GPU box, gloo in the CPU tests).
"""

import numpy as np


def shard_bounds(n, rank, world):
    """[lo, hi) of the contiguous block of ``n`` items owned by ``rank`` (blocks of ceil(n/world))."""
    per = (n + world - 1) // world
    lo = min(rank * per, n)
    return lo, min(lo + per, n)


def gather_mask(local_mask, n_total, group=None):
    """All-gather per-rank boolean masks (block-sharded by :func:`shard_bounds`) into the full mask.

    ``local_mask``: 1-D torch tensor (bool / uint8) on the backend's device (CUDA for nccl, CPU for
    gloo) holding this rank's block.  Returns a bool tensor of length ``n_total`` on every rank.
    """
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        assert local_mask.numel() == n_total
        return local_mask.to(torch.bool)
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    per = (n_total + world - 1) // world
    lo, hi = shard_bounds(n_total, rank, world)
    assert local_mask.numel() == hi - lo, "local mask does not match this rank's block"
    send = torch.zeros(per, dtype=torch.uint8, device=local_mask.device)
    send[: hi - lo] = local_mask.to(torch.uint8)
    recv = torch.empty(per * world, dtype=torch.uint8, device=local_mask.device)
    dist.all_gather_into_tensor(recv, send, group=group)
    return recv[:n_total].to(torch.bool)  # blocks are contiguous: rank r owns [r*per, (r+1)*per)


def validate_edges_sharded(check_edges, Q1, Q2, max_step_size, padding=0.0, group=None):
    """Edge validity mask (True = collision free) for all edges, computed block-sharded.

    ``check_edges(q1_block, q2_block, step, padding) -> hit`` is the rank-local validator, normally
    ``CollisionEngine.check_edges`` (CUDA tensors in, CUDA bool tensor out).  Every rank passes the
    same full ``Q1``/``Q2`` and receives the same full mask.
    """
    import torch
    import torch.distributed as dist

    n = Q1.shape[0]
    if dist.is_available() and dist.is_initialized():
        rank, world = dist.get_rank(group), dist.get_world_size(group)
    else:
        rank, world = 0, 1
    lo, hi = shard_bounds(n, rank, world)
    hit = check_edges(Q1[lo:hi], Q2[lo:hi], max_step_size, padding)
    if not isinstance(hit, torch.Tensor):
        hit = torch.as_tensor(np.asarray(hit))
    return ~gather_mask(hit, n, group)


def check_states_sharded(check_states, Q, padding=0.0, group=None):
    """Collision verdicts for all states, block-sharded; every rank receives the full mask."""
    import torch
    import torch.distributed as dist

    n = Q.shape[0]
    if dist.is_available() and dist.is_initialized():
        rank, world = dist.get_rank(group), dist.get_world_size(group)
    else:
        rank, world = 0, 1
    lo, hi = shard_bounds(n, rank, world)
    hit = check_states(Q[lo:hi], padding)
    if not isinstance(hit, torch.Tensor):
        hit = torch.as_tensor(np.asarray(hit))
    return gather_mask(hit, n, group)
