"""Batched PRM roadmap construction (BASELINE.json config 4).

The reference grows a roadmap node by node (`PRMPlanner.construct_roadmap` /
`connect_node`, /root/reference/src/pyroboplan/planning/prm.py:115-205): rejection-sample a free
configuration, take its `k` nearest *earlier* nodes within `radius`
(`Graph.get_nearest_neighbors`, planning/graph.py:271-301, sorted by distance) and keep the
edges whose straight segment is collision free (`has_collision_free_path(node.q, new_node.q, ...)`).

Every one of those steps is independent across nodes once the node set is fixed, so here they are
batched: (1) free nodes from the device sampler, (2) k nearest predecessors by brute-force L2 on
the GPU (`pbp_knn`, a hand-written tiled kernel with FP64 distance accumulation so the neighbour
order equals the reference's) -- across ranks every rank searches an interleaved share of the query blocks and one
all-gather assembles the rows --, (3) ALL candidate edges validated in one `check_edges` call per rank, the edges cut
where the prefix sum of their sample counts crosses multiples of total / world, one all-gather of the bit-packed
validity mask.  (The sampler stays replicated: the same Philox stream on every rank, ~33 k draws, one small round --
sharding it would replace 0.1 ms of compute by an all-gather of the same length.)
"""

import numpy as np

from ..distributed import knn_interleaved, validate_edges_balanced
from ..engine import get_engine


def nearest_predecessors(nodes, k, radius=np.inf, engine=None, predecessors_only=True, group=None):
    """For node i: its k nearest nodes j < i with distance <= radius, sorted by distance.

    nodes: CUDA float32 tensor [N, nq].  Returns (idx [N, k] int64 with -1 padding, dist [N, k]).
    Mirrors the reference's incremental semantics: a node only sees the nodes added before it.
    """
    import torch

    if engine is None:
        raise ValueError("nearest_predecessors needs the CollisionEngine that owns the device (engine=...)")
    idx = torch.full((nodes.shape[0], k), -1, dtype=torch.int64, device=nodes.device)
    dist = torch.full((nodes.shape[0], k), float("inf"), dtype=torch.float32, device=nodes.device)
    for k0 in range(0, k, 64):   # the kernel keeps at most 64 neighbours per pass; more is not a PRM use case
        if k0 > 0:
            raise ValueError("k > 64 is not supported")
        # under torch.distributed every rank searches an interleaved share of the query blocks; one all-gather
        i32, d = knn_interleaved(lambda first, stride: engine.knn(nodes, min(k, 64), radius, predecessors_only, first, stride),
                                 nodes.shape[0], min(k, 64), engine.KNN_QUERY_BLOCK, nodes.device, group)
        idx[:, : i32.shape[1]] = i32.to(torch.int64)
        dist[:, : d.shape[1]] = d
    return idx, dist


def construct_roadmap_batched(model, collision_model, n_nodes, k=10, radius=np.inf, max_step_size=0.05,
                              distance_padding=0.0, seed=0, nodes=None, group=None):
    """Builds a PRM roadmap with batched sampling, k-NN and edge validation.

    Returns a dict with
      nodes       float32 [N, nq]   collision-free configurations (device tensor)
      candidates  int64   [M, 2]    (neighbour j, node i) with j < i, in the reference's visiting order
      valid       bool    [M]       True where the segment nodes[j] -> nodes[i] is collision free
      edges       int64   [E, 2]    candidates[valid]
      lengths     float32 [M]       joint-space edge lengths
    Under torch.distributed every rank validates its block of the candidates and receives the
    full mask (one all-gather); without it the single process does everything.
    """
    import torch

    eng = get_engine(model, collision_model)
    if nodes is None:
        nodes = eng.sample_free_states(n_nodes, seed=seed, padding=distance_padding)
        if nodes.shape[0] < n_nodes:
            raise RuntimeError(f"sampler found only {nodes.shape[0]} of {n_nodes} free states")
    else:
        nodes = torch.as_tensor(np.asarray(nodes, dtype=np.float32) if not isinstance(nodes, torch.Tensor) else nodes,
                                device=eng.torch_device).to(torch.float32).contiguous()
    idx, dist = nearest_predecessors(nodes, k, radius, engine=eng, group=group)
    node_ids = torch.arange(nodes.shape[0], device=nodes.device)[:, None].expand_as(idx)
    keep = idx >= 0
    cand = torch.stack([idx[keep], node_ids[keep]], dim=1)  # row-major: node by node, nearest first
    lengths = dist[keep]
    q1 = nodes[cand[:, 0]].contiguous()
    q2 = nodes[cand[:, 1]].contiguous()
    if len(collision_model.collisionPairs) == 0:
        valid = torch.ones(cand.shape[0], dtype=torch.bool, device=nodes.device)
    else:
        valid = validate_edges_balanced(eng.check_edges, q1, q2, max_step_size, distance_padding, lengths=lengths, group=group)
    return {"nodes": nodes, "candidates": cand, "valid": valid, "edges": cand[valid], "lengths": lengths}
