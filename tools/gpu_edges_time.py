"""Config 3 (100 k random Panda edges, step 0.05) timing, median of 7."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import panda_models
from pyroboplan_b200.engine import CollisionEngine
model, cmodel = panda_models()
eng = CollisionEngine(model, cmodel, device=0)
rng = np.random.default_rng(1)
a = torch.as_tensor(rng.uniform(model.lowerPositionLimit, model.upperPositionLimit, size=(100000, 9)).astype(np.float32), device="cuda:0")
b = torch.as_tensor(rng.uniform(model.lowerPositionLimit, model.upperPositionLimit, size=(100000, 9)).astype(np.float32), device="cuda:0")
ts = []
for _ in range(9):
    torch.cuda.synchronize(); t0 = time.perf_counter(); eng.check_edges(a, b, 0.05); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
print(f"config 3: {sorted(ts)[len(ts)//2]*1e3:.3f} ms per 100 k edges")
