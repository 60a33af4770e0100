"""Batched PRM construction (BASELINE.json config 4) against the reference's incremental semantics."""

import numpy as np
import pytest

from conftest import make_panda

pytestmark = pytest.mark.gpu


def reference_neighbours(nodes, i, k, radius):
    """Graph.get_nearest_neighbors over the nodes added before node i (reference graph.py:271-301)."""
    d = np.linalg.norm(nodes[:i].astype(np.float64) - nodes[i].astype(np.float64), axis=1)
    order = np.argsort(d, kind="stable")
    return [j for j in order if d[j] <= radius][:k]


def test_batched_roadmap_small(panda_full, panda_oracle):
    import torch

    from pyroboplan_b200.planning.prm_batched import construct_roadmap_batched

    model, cmodel, _ = panda_full
    rm = construct_roadmap_batched(model, cmodel, n_nodes=400, k=6, radius=2.5, max_step_size=0.05, seed=11)
    nodes = rm["nodes"].cpu().numpy()
    cand = rm["candidates"].cpu().numpy()
    valid = rm["valid"].cpu().numpy()
    assert nodes.shape == (400, 9)
    assert not panda_oracle.check_states(nodes.astype(np.float64))[0].any()  # every node is free
    # candidate edges = each node's k nearest predecessors within the radius, nearest first
    by_node = {}
    for j, i in cand:
        by_node.setdefault(int(i), []).append(int(j))
    for i in range(0, 400, 17):
        assert by_node.get(i, []) == reference_neighbours(nodes, i, 6, 2.5)
    assert (cand[:, 0] < cand[:, 1]).all()
    # validity mask = has_collision_free_path(node_j.q, node_i.q) for every candidate (oracle)
    hit, _, _, _ = panda_oracle.check_edges(nodes[cand[:, 0]].astype(np.float64), nodes[cand[:, 1]].astype(np.float64), 0.05)
    assert (valid == ~hit.astype(bool)).all()
    assert (rm["edges"].cpu().numpy() == cand[valid]).all()
    assert 0 < valid.sum() < len(valid)


def test_config4_full_size_roadmap(panda_full, panda_oracle):
    """20 k free nodes, 15 nearest predecessors: ~300 k candidate edges validated in one call;
    the mask is checked against the oracle on a 3000-edge random subset and through the
    order-independent property mask(edge) == mask(same edge alone)."""
    import torch

    from pyroboplan_b200.engine import get_engine
    from pyroboplan_b200.planning.prm_batched import construct_roadmap_batched

    model, cmodel, _ = panda_full
    rm = construct_roadmap_batched(model, cmodel, n_nodes=20_000, k=15, radius=np.inf, max_step_size=0.05, seed=4)
    cand, valid = rm["candidates"], rm["valid"]
    assert cand.shape[0] == 20_000 * 15 - 15 * 16 // 2   # node i has min(i, 15) predecessors
    nodes = rm["nodes"]
    pick = torch.randperm(cand.shape[0], device=cand.device)[:3000]
    q1 = nodes[cand[pick, 0]].cpu().numpy().astype(np.float64)
    q2 = nodes[cand[pick, 1]].cpu().numpy().astype(np.float64)
    hit, _, _, _ = panda_oracle.check_edges(q1, q2, 0.05)
    assert (valid[pick].cpu().numpy() == ~hit.astype(bool)).all()
    eng = get_engine(model, cmodel)
    again = ~eng.check_edges(nodes[cand[pick, 0]].contiguous(), nodes[cand[pick, 1]].contiguous(), 0.05)
    assert (again == valid[pick]).all()
