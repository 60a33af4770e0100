#!/usr/bin/env python
"""Generate tests/golden/*.npz: seeded inputs + float64-oracle outputs for the GPU parity tests.

    python tools/make_golden.py

The reference itself (Pinocchio/coal) cannot run in the authoring container, so these vectors come
from the oracle (oracle/pbp_oracle.c), whose pair results are certificate-checked and which
reproduces every known-answer vector of the reference's own tests (tests/test_oracle_golden.py).
The three known-answer states of /root/reference/test/core/test_utils.py:129-212 are included
verbatim with the verdicts the reference asserts.  The oracle runs with the reference's mesh
semantics (surfaces: exact triangle tests wherever the hulls are within the padding, min distances
measured to the triangles).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from conftest import COLLISION_STATE, FREE_STATE_1, FREE_STATE_2, make_panda, make_two_dof, uniform_states  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def states_fixture(name, model, cmodel, n, seed, padding=0.0):
    orc = Oracle(model, cmodel)
    q = uniform_states(model, n, seed)
    hit, md, fp, _ = orc.check_states(q.astype(np.float64), padding=padding, want_all=True, use_cull=False)
    np.savez_compressed(os.path.join(OUT, f"{name}.npz"), q=q, hit=hit, min_dist=md, first_pair=fp, padding=padding)
    print(name, "N", n, "P(hit)", hit.mean())


def main():
    os.makedirs(OUT, exist_ok=True)
    m, c, _ = make_panda()
    states_fixture("panda_states", m, c, 3000, 101)
    states_fixture("panda_states_padding", m, c, 2000, 102, padding=0.025)
    m2, c2, _ = make_two_dof()
    states_fixture("two_dof_states", m2, c2, 3000, 103)
    ms, cs, _ = make_panda(objects=False)
    orc = Oracle(ms, cs)
    q = np.array([COLLISION_STATE, FREE_STATE_1, FREE_STATE_2])
    hit, md, fp, _ = orc.check_states(q, want_all=True, use_cull=False)
    assert list(hit) == [1, 0, 0]  # what the reference's test asserts
    np.savez_compressed(os.path.join(OUT, "panda_self_reference_states.npz"), q=q, hit=hit, min_dist=md, first_pair=fp)
    # edges: half long (mostly colliding), half short (mostly free)
    orc = Oracle(m, c)
    a = uniform_states(m, 400, 104)
    b = uniform_states(m, 400, 105)
    b[:250] = a[:250] + np.float32(0.12) * (b[:250] - a[:250])
    hit, fb, evals, _ = orc.check_edges(a.astype(np.float64), b.astype(np.float64), 0.05)
    counts = np.array([orc.segment_samples(a[i].astype(np.float64), b[i].astype(np.float64), 0.05) for i in range(len(a))])
    np.savez_compressed(os.path.join(OUT, "panda_edges.npz"), q1=a, q2=b, hit=hit, first_bad=fb, counts=counts, step=0.05)
    print("edges P(hit)", hit.mean(), "states evaluated by the sequential reference order", evals)


if __name__ == "__main__":
    main()
