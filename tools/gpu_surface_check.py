#!/usr/bin/env python
"""Surface (BVHModel) semantics on the GPU against the oracle's exact triangle tests: verdicts of the
boolean pipeline, the distance mode and the fused small-batch kernel on config-2 states, with the
mismatches classified.  Run through gpurun."""

import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle.oracle import Oracle  # noqa: E402
from pyroboplan_b200.engine import CollisionEngine  # noqa: E402
from pyroboplan_b200.models import panda  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    pad = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
    model, cmodel, vmodel = panda.load_models()
    panda.add_self_collisions(model, cmodel)
    panda.add_object_collisions(model, cmodel, vmodel)
    rng = np.random.default_rng(seed)
    q = rng.uniform(model.lowerPositionLimit, model.upperPositionLimit, size=(n, model.nq)).astype(np.float32)
    qd = torch.as_tensor(q, device="cuda:0")
    for sem in ("hull", "surface"):
        eng = CollisionEngine(model, cmodel, device=0, mesh_semantics=sem)
        orc = Oracle(model, cmodel, mesh_semantics=sem)
        hit = eng.check_states(qd, pad).cpu().numpy().astype(bool)
        hit_d, dist = eng.check_states(qd, pad, return_distance=True)
        hit_d, dist = hit_d.cpu().numpy().astype(bool), dist.cpu().numpy()
        hit_s = np.concatenate([eng.check_states(qd[i:i + 512], pad).cpu().numpy().astype(bool) for i in range(0, 65536, 512)])
        for _ in range(3):
            eng.check_states(qd)
        torch.cuda.synchronize()
        t0 = time.time()
        for _ in range(20):
            eng.check_states(qd)
        torch.cuda.synchronize()
        dt = (time.time() - t0) / 20
        t0 = time.time()
        ref, dref, _, _ = orc.check_states(q.astype(np.float64), padding=pad, want_all=True, use_cull=True)
        t_or = time.time() - t0
        ref = ref.astype(bool)
        print(f"[{sem}] seed {seed} padding {pad} {n} states: {n / dt:.3e} states/s on the GPU ({dt * 1e3:.3f} ms), oracle {t_or:.1f} s; P(hit) {ref.mean():.4f}")
        for name, h in (("pipeline", hit), ("distance mode", hit_d), ("fused small", hit_s)):
            bad = np.nonzero(h != ref[: len(h)])[0]
            far = [i for i in bad if abs(dref[i] - pad) > 1e-6]
            print(f"   {name}: {len(bad)} verdict mismatches, {len(far)} beyond 1e-6 m:",
                  [(int(i), bool(h[i]), float(dref[i])) for i in far[:8]])
        free = (~ref) & (~hit_d)
        err = np.abs(dist[free] - dref[free])
        print(f"   min distance over {free.sum()} free states: max err {err.max():.3e}, > 1e-5: {(err > 1e-5).sum()}, > 5e-4: {(err > 5e-4).sum()}")
        print("   stats", eng.stats())
        eng.close()


if __name__ == "__main__":
    main()
