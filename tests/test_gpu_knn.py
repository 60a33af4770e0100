"""pbp_knn (batched Graph.get_nearest_neighbors, /root/reference/src/pyroboplan/planning/graph.py:271-301)
against a float64 numpy brute force: identical neighbour indices, in the same order."""

import numpy as np
import pytest

from conftest import make_panda, make_two_dof, uniform_states

pytestmark = pytest.mark.gpu


def brute_force(nodes, k, radius, predecessors_only):
    n = len(nodes)
    x = nodes.astype(np.float64)
    idx = np.full((n, k), -1, dtype=np.int64)
    dist = np.full((n, k), np.inf)
    for i in range(n):
        lim = i if predecessors_only else n
        d = np.sqrt(((x[:lim] - x[i]) ** 2).sum(axis=1))
        cand = [j for j in np.argsort(d, kind="stable") if j != i and d[j] <= radius][:k]
        idx[i, : len(cand)] = cand
        dist[i, : len(cand)] = d[cand]
    return idx, dist


@pytest.mark.parametrize("predecessors_only", [True, False])
@pytest.mark.parametrize("radius", [np.inf, 2.0])
def test_knn_panda(panda_full, predecessors_only, radius):
    import torch

    from pyroboplan_b200.engine import get_engine

    model, cmodel, _ = panda_full
    eng = get_engine(model, cmodel)
    nodes = uniform_states(model, 1500, 3)
    idx, dist = eng.knn(torch.as_tensor(nodes, device="cuda:0"), 15, radius, predecessors_only)
    ridx, rdist = brute_force(nodes, 15, radius, predecessors_only)
    assert (idx.cpu().numpy() == ridx).all()
    found = ridx >= 0
    np.testing.assert_allclose(dist.cpu().numpy()[found], rdist[found], rtol=1e-6)
    assert np.isinf(dist.cpu().numpy()[~found]).all()


def test_knn_ties_small_and_generic_nq():
    import torch

    from pyroboplan_b200.engine import get_engine

    # 2-DOF model (nq = 2 specialisation): a lattice makes many exactly tied distances
    model, cmodel, _ = make_two_dof()
    eng = get_engine(model, cmodel)
    g = np.stack(np.meshgrid(np.arange(12), np.arange(12)), axis=-1).reshape(-1, 2).astype(np.float32) * 0.25
    for pred in (True, False):
        idx, _ = eng.knn(torch.as_tensor(g, device="cuda:0"), 8, 0.5, pred)
        assert (idx.cpu().numpy() == brute_force(g, 8, 0.5, pred)[0]).all()
    # fewer nodes than k, and a single node
    idx, dist = eng.knn(torch.as_tensor(g[:3], device="cuda:0"), 8, np.inf, False)
    assert (idx.cpu().numpy() == brute_force(g[:3], 8, np.inf, False)[0]).all()
    idx, dist = eng.knn(torch.as_tensor(g[:1], device="cuda:0"), 4, np.inf, False)
    assert (idx.cpu().numpy() == -1).all() and np.isinf(dist.cpu().numpy()).all()
    # Panda without obstacles still has nq = 9; a generic-nq path is exercised with k = 64 (the maximum)
    model9, cmodel9, _ = make_panda(objects=False)
    eng9 = get_engine(model9, cmodel9)
    nodes = uniform_states(model9, 300, 8)
    idx, _ = eng9.knn(torch.as_tensor(nodes, device="cuda:0"), 64, np.inf, False)
    assert (idx.cpu().numpy() == brute_force(nodes, 64, np.inf, False)[0]).all()
    with pytest.raises(RuntimeError):
        eng9.knn(torch.as_tensor(nodes, device="cuda:0"), 65)
