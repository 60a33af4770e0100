#!/usr/bin/env python
"""Throughput of BASELINE.json configs 1, 3, 4, 5 on one B200 (config 2 is bench.py's line).
Prints one JSON object; run through gpurun and keep the output under profiles/."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bench import make_batch, panda_models  # noqa: E402
from pyroboplan_b200.engine import CollisionEngine  # noqa: E402
from pyroboplan_b200.ik.differential_ik import DifferentialIk, DifferentialIkOptions  # noqa: E402
from pyroboplan_b200.models import two_dof  # noqa: E402
from pyroboplan_b200.planning.prm_batched import construct_roadmap_batched  # noqa: E402


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(reps):
        out = fn()
    ev1.record()
    torch.cuda.synchronize()
    return ev0.elapsed_time(ev1) / reps, out


dev = torch.device("cuda:0")
res = {}
# config 1: 2-DOF, 10 k states (latency case)
m2, c2, v2 = two_dof.load_models()
two_dof.add_object_collisions(m2, c2, v2)
e2 = CollisionEngine(m2, c2)
q2 = torch.as_tensor(np.random.default_rng(0).uniform(-3.14, 3.14, size=(10000, 2)).astype(np.float32), device=dev)
ms, hit = timed(lambda: e2.check_states(q2), reps=20)
res["config1_two_dof_10k_states"] = {"ms": ms, "states_per_s": 10000 / ms * 1e3, "p_collision": float(hit.float().mean())}

model, cmodel = panda_models()
eng = CollisionEngine(model, cmodel)
# config 3: 100 k random edges at 0.05 rad
rng = np.random.default_rng(1)
A = torch.as_tensor(rng.uniform(model.lowerPositionLimit, model.upperPositionLimit, size=(100000, 9)).astype(np.float32), device=dev)
B = torch.as_tensor(rng.uniform(model.lowerPositionLimit, model.upperPositionLimit, size=(100000, 9)).astype(np.float32), device=dev)
counts, _, _ = eng.discretize(A[:1], B[:1], 0.05)
n_states = int((torch.ceil((A.double() - B.double()).norm(dim=1) / 0.05) + 1).sum().item())
eng.stats(reset=True)
ms, hit = timed(lambda: eng.check_edges(A, B, 0.05), reps=5)
st = eng.stats(reset=True)
res["config3_100k_edges"] = {"ms": ms, "edges_per_s": 100000 / ms * 1e3, "states_covered": n_states,
                             "states_covered_per_s": n_states / ms * 1e3, "states_evaluated_per_call": st["states"] / 7,
                             "p_edge_collision": float(hit.float().mean())}
# short edges (RRT-like extensions, mostly free): every sample is evaluated
Bs = A + 0.1 * (B - A)
n_short = int((torch.ceil((A.double() - Bs.double()).norm(dim=1) / 0.05) + 1).sum().item())
ms, hit = timed(lambda: eng.check_edges(A, Bs, 0.05), reps=5)
res["config3b_100k_short_edges"] = {"ms": ms, "edges_per_s": 100000 / ms * 1e3, "states_covered_per_s": n_short / ms * 1e3,
                                    "p_edge_collision": float(hit.float().mean())}
# config 4: PRM 20 k nodes, k = 15
t0 = time.perf_counter()
rm = construct_roadmap_batched(model, cmodel, n_nodes=20000, k=15, radius=np.inf, max_step_size=0.05, seed=4)
torch.cuda.synchronize()
t_first = time.perf_counter() - t0
walls = []
for rep in range(5):
    t0 = time.perf_counter()
    rm = construct_roadmap_batched(model, cmodel, n_nodes=20000, k=15, radius=np.inf, max_step_size=0.05, seed=5 + rep)
    torch.cuda.synchronize()
    walls.append(time.perf_counter() - t0)
t_prm = sorted(walls)[len(walls) // 2]
nodes = rm["nodes"]
ms_knn, _ = timed(lambda: eng.knn(nodes, 15, np.inf, True), reps=5)
res["config4_prm_20k_nodes_k15"] = {"wall_s_median_of_5": t_prm, "wall_s_all": walls, "first_call_s": t_first, "candidate_edges": int(rm["candidates"].shape[0]),
                                    "valid_edges": int(rm["valid"].sum().item()), "knn_ms": ms_knn}
# config 5: 64 k IK targets
fid = model.getFrameId("panda_hand")
goals = eng.sample_free_states(65536, seed=2)
_, _, oMf = eng.fk(goals, joints=False, geoms=False, frames=True)
targets = oMf[:, fid].contiguous()
opts = DifferentialIkOptions(ignore_joint_indices=[7, 8])
ik = DifferentialIk(model, collision_model=cmodel, options=opts)
init = eng.sample_free_states(65536, seed=3).double()
ms_try, (Q1, status, iters, _) = timed(lambda: eng.ik_iterate(fid, targets, init, active_mask=0x7F), reps=3, warm=1)
res["config5_ik_one_try_64k"] = {"ms": ms_try, "tries_per_s": 65536 / ms_try * 1e3, "iterations": int(iters.sum().item()),
                                 "iterations_per_s": float(iters.sum().item()) / ms_try * 1e3,
                                 "converged_first_try": float((status & 1).float().mean())}
ik_walls = []
for rep in range(3):   # untimed: sizes the device buffers for exactly these restart draws (as tools/gpu_ik_sweep.py)
    ik.solve_batch("panda_hand", targets, seed=5 + rep)
for rep in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    Q, solved, tries = ik.solve_batch("panda_hand", targets, seed=5 + rep)
    torch.cuda.synchronize()
    ik_walls.append(time.perf_counter() - t0)
t_ik = sorted(ik_walls)[1]
res["config5_ik_solve_batch_64k"] = {"wall_s_median_of_3": t_ik, "wall_s_all": ik_walls, "solves_per_s": 65536 / t_ik, "success_rate": float(solved.float().mean()),
                                     "mean_tries": float(tries.float().mean())}
print(json.dumps(res, indent=1))
