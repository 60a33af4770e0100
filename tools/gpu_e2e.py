"""Host-buffer path (pbp_check_states_host / _host_f64) of a 1 Mi-state call: slice size and host-thread sweeps.
    python tools/gpu_e2e.py"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import make_batch, panda_models
from pyroboplan_b200.engine import CollisionEngine
model, cmodel = panda_models()
n = 1 << 20
host = [torch.from_numpy(make_batch(model, n, b)).pin_memory() for b in range(4)]
host_np = [h.numpy() for h in host]
f64 = [h.numpy().astype(np.float64) for h in host[:2]]
def rate(eng, batches, reps=30):
    for w in range(3): eng.check_states(batches[w % len(batches)])
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for k in range(reps): eng.check_states(batches[k % len(batches)])
    torch.cuda.synchronize(); return n * reps / (time.perf_counter() - t0)
for s in [int(x) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else "131072,196608,262144,349526,524288,1048576".split(","))]:
    os.environ["PBP_HOST_SLICE"] = str(s)
    eng = CollisionEngine(model, cmodel, device=0)
    print(f"slice {s:8d}: pinned f32 {rate(eng, host_np):.3e} states/s", flush=True)
    eng.close()
os.environ["PBP_HOST_SLICE"] = "262144"
for th in (4, 8, 12, 16, 24):
    os.environ["PBP_HOST_THREADS"] = str(th)
    eng = CollisionEngine(model, cmodel, device=0)
    print(f"host threads {th:2d}: pageable f64 {rate(eng, f64, 15):.3e} states/s", flush=True)
    eng.close()
