"""Batched RRT* rewire against a line-by-line sequential replay of the reference's rule
(/root/reference/src/pyroboplan/planning/rrt.py:376-402) with the float64 oracle as edge checker."""

import numpy as np
import pytest

from conftest import uniform_states

pytestmark = pytest.mark.gpu


def sequential_rewire(orc, node_qs, node_costs, q_new, parent, max_rewire_dist, step):
    best = parent
    min_cost = node_costs[parent] + np.linalg.norm(node_qs[parent] - q_new)
    for k in range(len(node_qs)):
        if k == parent or np.array_equal(node_qs[k], q_new):
            continue
        d = np.linalg.norm(node_qs[k] - q_new)
        if d > max_rewire_dist:
            continue
        c = node_costs[k] + d
        if c < min_cost:
            hit, _, _, _ = orc.check_edges(q_new[None, :], node_qs[k][None, :], step)
            if not hit[0]:
                best, min_cost = k, c
    return best, min_cost


def test_rewire_matches_sequential_rule(panda_full, panda_oracle):
    from pyroboplan_b200.planning.rrt_batched import rewire_choice, validate_extensions

    model, cmodel, _ = panda_full
    rng = np.random.default_rng(5)
    changed = 0
    for trial in range(12):
        centre = uniform_states(model, 1, 100 + trial)[0].astype(np.float64)
        nodes = (centre + rng.normal(scale=0.35, size=(60, 9))).astype(np.float32).astype(np.float64)
        nodes = np.clip(nodes, model.lowerPositionLimit, model.upperPositionLimit)
        costs = rng.uniform(0.0, 3.0, size=60)
        q_new = nodes[0] + rng.normal(scale=0.1, size=9).astype(np.float32)
        q_new = np.clip(q_new, model.lowerPositionLimit, model.upperPositionLimit).astype(np.float32).astype(np.float64)
        parent = int(rng.integers(0, 60))
        got = rewire_choice(model, cmodel, nodes, costs, q_new, parent, max_rewire_dist=2.0, max_step_size=0.05)
        ref = sequential_rewire(panda_oracle, nodes, costs, q_new, parent, 2.0, 0.05)
        assert got[0] == ref[0] and got[1] == pytest.approx(ref[1], abs=1e-12)
        changed += int(got[0] != parent)
    assert changed > 0
    a, b = uniform_states(model, 500, 1).astype(np.float64), uniform_states(model, 500, 2).astype(np.float64)
    b = a + 0.1 * (b - a)
    ref, _, _, _ = panda_oracle.check_edges(a, b, 0.05)
    assert (validate_extensions(model, cmodel, a, b, 0.05) == ~ref.astype(bool)).all()
