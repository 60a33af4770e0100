"""The reference planners' call order on the GPU drop-in functions against the identical loop on the oracle
(tests/planner_loops.py): same seeded samples => the same trees / roadmaps, node for node.

Scenarios follow the reference's planner tests (/root/reference/test/planning/test_rrt.py:12-111 -- the Panda
start / goal pair, RRT-Connect, bidirectional, RRT* --, test/planning/test_prm.py:22-75) plus the obstacle
versions of both example models, where the collision checks actually decide."""

import time

import numpy as np
import pytest

from conftest import make_panda, make_two_dof
from planner_loops import DropInBackend, OracleBackend, prm_build, rrt_plan, same_trees

pytestmark = pytest.mark.gpu

Q_START = np.array([0.0, 1.57, 0.0, 0.0, 1.57, 1.57, 0.0, 0.0, 0.0])
Q_GOAL = np.array([1.57, 1.57, 0.0, 0.0, 1.57, 1.57, 0.0, 0.0, 0.0])
BASE = dict(max_step_size=0.05, max_connection_dist=np.inf, rrt_connect=False, bidirectional_rrt=False, rrt_star=False,
            max_rewire_dist=np.inf, goal_biasing_probability=0.0)


def _backends(model, cmodel, batched=True):
    from oracle.oracle import Oracle

    return DropInBackend(model, cmodel, batched), OracleBackend(model, cmodel, Oracle(model, cmodel))


@pytest.mark.parametrize("name,over", [
    ("reference_defaults", dict()),
    ("vanilla", dict(max_connection_dist=1.0, goal_biasing_probability=0.15)),
    ("connect_bidirectional", dict(max_connection_dist=0.5, rrt_connect=True, bidirectional_rrt=True)),
    ("star", dict(max_connection_dist=0.5, rrt_connect=True, bidirectional_rrt=True, rrt_star=True, max_rewire_dist=3.0)),
])
def test_rrt_on_panda_with_obstacles_grows_the_oracles_tree(name, over):
    model, cmodel, _ = make_panda()
    gpu, cpu = _backends(model, cmodel)
    opt = dict(BASE, **over)
    a = rrt_plan(gpu, model, Q_START, Q_GOAL, opt, seed=1234, max_iters=60)
    b = rrt_plan(cpu, model, Q_START, Q_GOAL, opt, seed=1234, max_iters=60)
    assert same_trees(a, b), f"{name}: trees differ ({[len(t.q) for t in a['trees']]} vs {[len(t.q) for t in b['trees']]} nodes)"
    assert sum(len(t.q) for t in a["trees"]) > 4


def test_rrt_star_batched_rewire_equals_the_sequential_calls():
    model, cmodel, _ = make_panda()
    opt = dict(BASE, max_connection_dist=0.5, rrt_connect=True, bidirectional_rrt=True, rrt_star=True, max_rewire_dist=3.0)
    a = rrt_plan(DropInBackend(model, cmodel, batched=True), model, Q_START, Q_GOAL, opt, seed=7, max_iters=40)
    b = rrt_plan(DropInBackend(model, cmodel, batched=False), model, Q_START, Q_GOAL, opt, seed=7, max_iters=40)
    assert same_trees(a, b)


def test_rrt_trivial_direct_connection():
    model, cmodel, _ = make_panda()
    gpu, cpu = _backends(model, cmodel)
    a = rrt_plan(gpu, model, Q_START, Q_START + 0.01, BASE, seed=0, max_iters=5)     # test_rrt.py:12-26
    b = rrt_plan(cpu, model, Q_START, Q_START + 0.01, BASE, seed=0, max_iters=5)
    assert a["found"] and a["iters"] == 0 and same_trees(a, b)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_rrt_on_two_dof_with_obstacles(seed):
    model, cmodel, _ = make_two_dof()
    gpu, cpu = _backends(model, cmodel)
    opt = dict(BASE, max_connection_dist=0.25, rrt_connect=True, bidirectional_rrt=True, rrt_star=True, max_rewire_dist=1.0)
    q0, q1 = np.array([0.0, 0.0]), np.array([2.0, 1.0])
    a = rrt_plan(gpu, model, q0, q1, opt, seed=seed, max_iters=150)
    b = rrt_plan(cpu, model, q0, q1, opt, seed=seed, max_iters=150)
    assert same_trees(a, b)


def test_prm_roadmaps_equal(capsys):
    for make, n, radius, k in ((make_two_dof, 120, 1.0, 8), (make_panda, 80, 3.14, 10)):
        model, cmodel, _ = make()
        gpu, cpu = _backends(model, cmodel)
        t0 = time.perf_counter()
        nodes_a, edges_a = prm_build(gpu, model, n, radius, k, 0.05, seed=3)
        t1 = time.perf_counter()
        nodes_b, edges_b = prm_build(cpu, model, n, radius, k, 0.05, seed=3)
        t2 = time.perf_counter()
        assert np.array_equal(nodes_a, nodes_b) and edges_a == edges_b and len(edges_a) > 0
        with capsys.disabled():
            print(f"\n[prm loop] nq={model.nq}: {n} nodes, {len(edges_a)} edges; drop-in {t1 - t0:.3f} s, oracle {t2 - t1:.3f} s")
