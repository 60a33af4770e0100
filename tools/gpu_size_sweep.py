#!/usr/bin/env python
"""Device time of check_states vs batch size, with the per-kernel split (run through gpurun)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bench import make_batch, panda_models  # noqa: E402
from pyroboplan_b200.engine import CollisionEngine  # noqa: E402

model, cmodel = panda_models()
eng = CollisionEngine(model, cmodel)
dev = torch.device("cuda:0")
for n in [1 << 14, 1 << 16, 1 << 17, 1 << 18, 1 << 20]:
    Q = torch.as_tensor(make_batch(model, n, 5), device=dev)
    for _ in range(5):
        eng.check_states(Q)
    torch.cuda.synchronize()
    eng.set_profiling(True)
    eng.profile(reset=True)
    reps = 50
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    ev0.record()
    for _ in range(reps):
        eng.check_states(Q)
    ev1.record()
    t_launch = time.perf_counter() - t0
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / reps
    p = eng.profile(reset=True)
    eng.set_profiling(False)
    r = max(p["rounds"], 1)
    print(f"N={n:8d}  {ms*1e3:8.1f} us/call  ({n/ms*1e3:.3e}/s)  host enqueue {t_launch/reps*1e6:6.1f} us/call  "
          f"bp {p['broadphase_ms']/r*1e3:7.1f}  setup {p['item_setup_ms']/r*1e3:7.1f}  np {p['narrowphase_ms']/r*1e3:7.1f} us")
    # host path with pinned input
    Qh = torch.from_numpy(make_batch(model, n, 6)).pin_memory().numpy()
    for _ in range(3):
        eng.check_states(Qh)
    t0 = time.perf_counter()
    for _ in range(20):
        eng.check_states(Qh)
    dt = (time.perf_counter() - t0) / 20
    print(f"           host path {dt*1e6:8.1f} us/call ({n/dt:.3e}/s)")
