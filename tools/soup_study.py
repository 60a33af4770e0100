#!/usr/bin/env python
"""How often does the convex-hull semantics of this engine differ from the surface (triangle-soup)
semantics of Pinocchio's BVHModel meshes?  CPU study on the benchmark's own inputs (float64 oracle,
exact triangle tests; DESIGN.md 2 "Mesh semantics").

    python tools/soup_study.py [N] [padding]

Prints, for N uniformly sampled Panda configurations (config 2's distribution, seed 0): the number
of states whose verdict differs, and the pairs responsible.
"""

import sys
import time
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle.oracle import Oracle  # noqa: E402
from pyroboplan_b200.models import panda  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
    padding = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
    model, cmodel, vmodel = panda.load_models()
    panda.add_self_collisions(model, cmodel)
    panda.add_object_collisions(model, cmodel, vmodel)
    orc = Oracle(model, cmodel)
    orc.set_mesh_triangles(cmodel)
    rng = np.random.default_rng(0)
    q = rng.uniform(model.lowerPositionLimit, model.upperPositionLimit, size=(n, model.nq)).astype(np.float32).astype(np.float64)
    t0 = time.time()
    hh, sh, fl, tests = orc.soup_check_states(q, padding)
    dt = time.time() - t0
    diff = np.nonzero(hh != sh)[0]
    print(f"{n} states, padding {padding}: hull verdict hit {hh.mean():.4f}, surface verdict hit {sh.mean():.4f}; "
          f"{tests} pair triangle tests in {dt:.1f} s")
    print(f"states whose verdict differs: {len(diff)} ({len(diff) / n:.2e}); states with at least one hull-only pair: {(fl > 0).sum()}")
    geoms = cmodel.geometryObjects
    for i in diff[:40]:
        _, oMg = orc.fk(q[i])
        for p in cmodel.collisionPairs:
            ds, _, _, _, exact = orc.pair_signed_distance_placed(oMg, p.first, p.second)
            if ds <= padding:
                d_s = orc.soup_pair_distance(oMg, p.first, p.second)
                print(f"  state {i}: {geoms[p.first].name} - {geoms[p.second].name}: hull signed distance {ds * 1e3:+.4f} mm, "
                      f"surface distance {d_s * 1e3:.4f} mm")


if __name__ == "__main__":
    main()
