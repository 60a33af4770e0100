"""UR5 / UR10 (non-convex meshes): states/s of the boolean check on 200 k uniform states."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_gpu_ur import _ur, _states
from pyroboplan_b200.engine import CollisionEngine
for which in (5, 10):
    m, c = _ur(which)
    eng = CollisionEngine(m, c, device=0)
    qd = torch.as_tensor(_states(m, 200000, 3), device="cuda:0")
    for _ in range(2): eng.check_states(qd)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): h = eng.check_states(qd)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
    eng.set_profiling(True); eng.profile(reset=True); eng.check_states(qd); torch.cuda.synchronize(); pr = eng.profile(reset=True)
    print(f"UR{which}: {len(qd) / dt:.3e} states/s ({dt * 1e3:.2f} ms per 200 k states), P(hit) {h.float().mean():.3f}", {k: round(v * 1e3) for k, v in pr.items() if k.endswith('_ms')}, "us")
