"""Multi-GPU sharding of the hot path (one process per GPU, ``torch.distributed``).

States and edges are independent given the (replicated) model, so a batch is cut into
contiguous per-rank blocks and every rank validates its block with its own engine -- there is no
collective in the data path.  The exchange steps belong to PRM roadmap construction
(BASELINE.json config 4), where every rank needs the whole roadmap: one ``all_gather_into_tensor``
of the interleaved k-NN rows, one of the BIT-PACKED edge-validity masks (300 k edges = 37.5 KB),
NCCL over NVLink on the GPU box, gloo in the CPU tests.  Edges are cut where the prefix sum of their
sample counts crosses multiples of total / world, not into equal counts: a long edge
costs many more states than a short one.
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


# ----------------------------------------------------------------------------------------------
# PRM roadmap construction across ranks: balanced splits, bit-packed masks
# ----------------------------------------------------------------------------------------------
def _world(group=None):
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def balanced_bounds(weights, world):
    """Cut points [world + 1] of contiguous blocks with (nearly) equal weight sums: block r = [cut[r], cut[r+1]).

    ``weights``: 1-D torch tensor or numpy array of non-negative work estimates per item (for edges:
    samples per edge, ceil(length / step) + 1).  Identical on every rank for identical weights."""
    import torch

    w = torch.as_tensor(weights).to(torch.float64).reshape(-1)
    n = w.numel()
    if n == 0:
        return [0] * (world + 1)
    csum = torch.cumsum(w, dim=0)
    total = float(csum[-1])
    targets = torch.arange(1, world, dtype=torch.float64, device=w.device) * (total / world)
    inner = torch.searchsorted(csum, targets, right=False).clamp_(0, n).tolist() if world > 1 else []
    cuts = [0] + [int(c) for c in inner] + [n]
    for r in range(1, len(cuts)):      # monotone even when many weights are zero
        cuts[r] = max(cuts[r], cuts[r - 1])
    return cuts


def pack_bits(mask):
    """bool / uint8 tensor [n] -> uint8 tensor [ceil(n / 8)], bit i of byte b = mask[8 b + i]."""
    import torch

    m = mask.to(torch.uint8).reshape(-1)
    pad = (-m.numel()) % 8
    if pad:
        m = torch.cat([m, torch.zeros(pad, dtype=torch.uint8, device=m.device)])
    weights = torch.tensor([1, 2, 4, 8, 16, 32, 64, 128], dtype=torch.uint8, device=m.device)
    return (m.reshape(-1, 8) * weights).sum(dim=1, dtype=torch.int32).to(torch.uint8)


def unpack_bits(packed, n):
    """Inverse of :func:`pack_bits` -> bool tensor [n]."""
    import torch

    shifts = torch.arange(8, dtype=torch.uint8, device=packed.device)
    bits = (packed.reshape(-1, 1) >> shifts) & 1
    return bits.reshape(-1)[:n].to(torch.bool)


def gather_mask_packed(local_mask, cuts, group=None):
    """All-gather per-rank masks of UNEQUAL blocks (block r = [cuts[r], cuts[r+1])), bit-packed on the wire.
    Returns the full bool mask [cuts[-1]] on every rank."""
    import torch
    import torch.distributed as dist

    rank, world = _world(group)
    n = cuts[-1]
    if world == 1:
        assert local_mask.numel() == n
        return local_mask.to(torch.bool)
    assert local_mask.numel() == cuts[rank + 1] - cuts[rank], "local mask does not match this rank's block"
    per = (max(cuts[r + 1] - cuts[r] for r in range(world)) + 7) // 8
    send = torch.zeros(per, dtype=torch.uint8, device=local_mask.device)
    packed = pack_bits(local_mask)
    send[: packed.numel()] = packed
    recv = torch.empty(per * world, dtype=torch.uint8, device=local_mask.device)
    dist.all_gather_into_tensor(recv, send, group=group)
    parts = [unpack_bits(recv[r * per:(r + 1) * per], cuts[r + 1] - cuts[r]) for r in range(world)]
    return torch.cat(parts)


def validate_edges_balanced(check_edges, Q1, Q2, max_step_size, padding=0.0, lengths=None, group=None):
    """Edge validity mask (True = collision free) on every rank; edges cut by prefix-summed sample counts.

    ``lengths``: joint-space edge lengths [E] (computed from Q1 / Q2 when omitted); the per-edge work
    estimate is the reference's sample count ``ceil(length / step) + 1``
    (/root/reference/src/pyroboplan/planning/utils.py:137)."""
    import torch

    rank, world = _world(group)
    n = Q1.shape[0]
    if world == 1:
        hit = check_edges(Q1, Q2, max_step_size, padding)
        return ~(hit if isinstance(hit, torch.Tensor) else torch.as_tensor(np.asarray(hit))).to(torch.bool)
    if lengths is None:
        lengths = torch.linalg.vector_norm(torch.as_tensor(Q2) - torch.as_tensor(Q1), dim=1)
    counts = torch.ceil(torch.as_tensor(lengths).to(torch.float64) / max_step_size) + 1.0
    cuts = balanced_bounds(counts, world)
    lo, hi = cuts[rank], cuts[rank + 1]
    if hi > lo:
        hit = check_edges(Q1[lo:hi], Q2[lo:hi], max_step_size, padding)
        hit = hit if isinstance(hit, torch.Tensor) else torch.as_tensor(np.asarray(hit))
    else:
        hit = torch.zeros(0, dtype=torch.bool, device=torch.as_tensor(Q1).device if isinstance(Q1, torch.Tensor) else "cpu")
    return ~gather_mask_packed(hit, cuts, group)


def knn_interleaved(knn_blocks, n, k, query_block, device, group=None):
    """k-NN rows of all ``n`` nodes on every rank, every rank computing an interleaved share of the query blocks.

    ``knn_blocks(block_first, block_stride) -> (idx [rows, k] int32, dist [rows, k] float32)`` computes query
    blocks block_first, block_first + block_stride, ... of ``query_block`` consecutive nodes each (rows in that
    order, padded with -1 / inf to whole blocks): ``CollisionEngine.knn`` with those arguments.  Interleaving
    balances the predecessor-only search, whose work grows with the node index.  ONE all-gather: the float32
    distances travel bit-cast in the same int32 buffer as the indices."""
    import torch
    import torch.distributed as dist

    rank, world = _world(group)
    if world == 1:
        idx, d = knn_blocks(0, 1)
        return idx[:n], d[:n]
    nblocks = (n + query_block - 1) // query_block
    per_blocks = (nblocks + world - 1) // world
    rows = per_blocks * query_block
    idx, d = knn_blocks(rank, world)
    send = torch.full((rows, 2 * k), -1, dtype=torch.int32, device=device)
    send[:, k:] = torch.full((rows, k), float("inf"), dtype=torch.float32, device=device).view(torch.int32)
    send[: idx.shape[0], :k] = idx
    send[: d.shape[0], k:] = d.contiguous().view(torch.int32)
    recv = torch.empty((world * rows, 2 * k), dtype=torch.int32, device=device)
    dist.all_gather_into_tensor(recv, send, group=group)
    # recv[r, t, j] = query block (r + t * world), row j  ->  node order = (t, r, j)
    full = recv.reshape(world, per_blocks, query_block, 2 * k).permute(1, 0, 2, 3).reshape(-1, 2 * k)[:n]
    return full[:, :k].contiguous(), full[:, k:].contiguous().view(torch.float32)
