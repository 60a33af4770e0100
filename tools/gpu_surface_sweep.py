#!/usr/bin/env python
"""Verdict parity of the surface semantics on many seeds: engine (boolean pipeline) against the oracle's
exact triangle tests, config-2 distribution.  Mismatching states are listed with the oracle's min distance.

    python tools/gpu_surface_sweep.py [first_seed] [n_seeds] [states_per_seed]
"""

import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle.oracle import Oracle  # noqa: E402
from pyroboplan_b200.engine import CollisionEngine  # noqa: E402
from pyroboplan_b200.models import panda  # noqa: E402


def main():
    first = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    count = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    n = int(sys.argv[3]) if len(sys.argv) > 3 else 1000000
    model, cmodel, vmodel = panda.load_models()
    panda.add_self_collisions(model, cmodel)
    panda.add_object_collisions(model, cmodel, vmodel)
    eng = CollisionEngine(model, cmodel, device=0)
    orc = Oracle(model, cmodel)
    total = bad_total = far_total = 0
    for seed in range(first, first + count):
        q = np.random.default_rng(seed).uniform(model.lowerPositionLimit, model.upperPositionLimit, size=(n, model.nq)).astype(np.float32)
        hit = eng.check_states(torch.as_tensor(q, device="cuda:0")).cpu().numpy().astype(bool)
        ref = orc.check_states(q.astype(np.float64))[0].astype(bool)
        bad = np.nonzero(hit != ref)[0]
        d = orc.check_states(q[bad].astype(np.float64), want_all=True, use_cull=False)[1] if len(bad) else np.zeros(0)
        far = [(int(i), bool(hit[i]), float(x)) for i, x in zip(bad, d) if abs(x) > 1e-6]
        total += n
        bad_total += len(bad)
        far_total += len(far)
        print(f"seed {seed}: {n} states, P(hit) {ref.mean():.4f}, {len(bad)} verdict mismatches, {len(far)} beyond 1e-6 m {far}", flush=True)
    print(f"total {total} states: {bad_total} mismatches, {far_total} beyond 1e-6 m of contact")


if __name__ == "__main__":
    main()
