"""Sequential planner loops on the GPU drop-in functions: plans/s and per-call latency (one JSON line).

    python tools/gpu_planner_bench.py > gpurun_out/planner_bench.json

The loops are tests/planner_loops.py (the reference planners' call order, rrt.py:125-318,350-402, prm.py:115-205);
the numbers that matter for a sequential planner are the single-call latencies, not the batch throughput."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import make_panda, make_two_dof
from planner_loops import DropInBackend, OracleBackend, prm_build, rrt_plan, same_trees
from oracle.oracle import Oracle

BASE = dict(max_step_size=0.05, max_connection_dist=0.5, rrt_connect=True, bidirectional_rrt=True, rrt_star=False, max_rewire_dist=3.0, goal_biasing_probability=0.0)
Q_START = np.array([0.0, 1.57, 0.0, 0.0, 1.57, 1.57, 0.0, 0.0, 0.0]); Q_GOAL = np.array([1.57, 1.57, 0.0, 0.0, 1.57, 1.57, 0.0, 0.0, 0.0])
out = {}
for label, make, q0, q1 in (("panda", make_panda, Q_START, Q_GOAL), ("two_dof", make_two_dof, np.array([0.0, 0.0]), np.array([2.0, 1.0]))):
    model, cmodel, _ = make()
    gpu, cpu = DropInBackend(model, cmodel), OracleBackend(model, cmodel, Oracle(model, cmodel))
    for star in (False, True):
        opt = dict(BASE, rrt_star=star)
        rrt_plan(gpu, model, q0, q1, opt, seed=0, max_iters=20)   # warm-up (engine creation, JIT)
        res = {"plans": 0, "found": 0, "same_as_oracle": 0, "nodes": 0}
        t_gpu = t_cpu = 0.0
        for seed in range(20):
            t0 = time.perf_counter(); a = rrt_plan(gpu, model, q0, q1, opt, seed=seed, max_iters=200); t_gpu += time.perf_counter() - t0
            t0 = time.perf_counter(); b = rrt_plan(cpu, model, q0, q1, opt, seed=seed, max_iters=200); t_cpu += time.perf_counter() - t0
            res["plans"] += 1; res["found"] += int(a["found"]); res["same_as_oracle"] += int(same_trees(a, b)); res["nodes"] += sum(len(t.q) for t in a["trees"])
        res.update(plans_per_s_gpu=res["plans"] / t_gpu, plans_per_s_oracle=res["plans"] / t_cpu, ms_per_plan_gpu=1e3 * t_gpu / res["plans"])
        out[f"{label}_rrt_connect{'_star' if star else ''}"] = res
    t0 = time.perf_counter(); nodes, edges = prm_build(gpu, model, 300, 3.14 if model.nq > 2 else 1.0, 10, 0.05, seed=0); t = time.perf_counter() - t0
    out[f"{label}_prm_incremental_300"] = {"s": t, "edges": len(edges), "nodes_per_s": 300 / t}
    from pyroboplan_b200.core.utils import check_collisions_at_state
    from pyroboplan_b200.planning.utils import has_collision_free_path
    qs = np.random.default_rng(0).uniform(model.lowerPositionLimit, model.upperPositionLimit, size=(2000, model.nq))
    t0 = time.perf_counter()
    for q in qs: check_collisions_at_state(model, cmodel, q)
    out[f"{label}_check_collisions_at_state_us"] = 1e6 * (time.perf_counter() - t0) / len(qs)
    t0 = time.perf_counter()
    for q in qs[:1000]: has_collision_free_path(q, q + 0.3 * (qs[0] - q) / max(np.linalg.norm(qs[0] - q), 1e-9), 0.05, model, cmodel)
    out[f"{label}_has_collision_free_path_0.3rad_us"] = 1e6 * (time.perf_counter() - t0) / 1000
print(json.dumps(out))
