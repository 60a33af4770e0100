#!/usr/bin/env python
"""Wall-clock latency of the single-query drop-in calls (what a sequential RRT loop pays)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bench import panda_models  # noqa: E402
from pyroboplan_b200.core.utils import check_collisions_at_state, get_random_collision_free_state  # noqa: E402
from pyroboplan_b200.engine import get_engine  # noqa: E402
from pyroboplan_b200.planning.utils import has_collision_free_path  # noqa: E402

model, cmodel = panda_models()
eng = get_engine(model, cmodel)
rng = np.random.default_rng(0)
qs = rng.uniform(model.lowerPositionLimit, model.upperPositionLimit, size=(2000, 9))


def timeit(fn, n):
    for i in range(20):
        fn(i)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(n):
        fn(i)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e6


print("check_collisions_at_state        %.1f us/call" % timeit(lambda i: check_collisions_at_state(model, cmodel, qs[i % 2000]), 2000))
print("has_collision_free_path (0.5 rad) %.1f us/call" % timeit(
    lambda i: has_collision_free_path(qs[i % 2000], qs[i % 2000] + 0.5 / 3.0, 0.05, model, cmodel), 1000))
print("has_collision_free_path (random)  %.1f us/call" % timeit(
    lambda i: has_collision_free_path(qs[i % 2000], qs[(i + 1) % 2000], 0.05, model, cmodel), 1000))
np.random.seed(0)
print("get_random_collision_free_state  %.1f us/call" % timeit(lambda i: get_random_collision_free_state(model, cmodel), 500))
for n in (1, 16, 128, 1024):
    Q = qs[:n].astype(np.float32)
    print(f"check_states(numpy, N={n:5d})       %.1f us/call" % timeit(lambda i: eng.check_states(Q), 500))
    Qd = torch.as_tensor(Q, device="cuda:0")
    print(f"check_states(cuda,  N={n:5d})       %.1f us/call (async enqueue + final sync amortised)" % timeit(lambda i: eng.check_states(Qd), 500))
