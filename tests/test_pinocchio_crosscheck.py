"""Cross-check against the REAL reference arithmetic (Pinocchio + coal) -- runs wherever ``import pinocchio``
works, skips cleanly everywhere else (the authoring container and the project's GPU box have no Pinocchio).

What it pins when it runs (SURVEY.md 8c "weakly pinned": no reference test holds an obstacle pair, a
padding > 0 verdict or a min distance):
  * the oracle's verdicts == ``pinocchio.computeCollisions`` through the reference's own
    ``check_collisions_at_state`` loop (/root/reference/src/pyroboplan/core/utils.py:8-51) on the committed
    golden states (self + obstacle pairs, padding 0 and > 0) for every state farther than 1e-6 m from contact;
  * the oracle's min distances == ``computeDistances`` (core/utils.py:54-87) within 1e-5 m;
  * ``from_pinocchio`` on real Pinocchio objects yields the same device tables as the native containers;
  * (GPU) the drop-in functions called WITH Pinocchio objects answer like Pinocchio.
"""

import os

import numpy as np
import pytest

from conftest import COLLISION_STATE, FREE_STATE_1, FREE_STATE_2, make_panda
from oracle import pinocchio_ref

pytestmark = pytest.mark.skipif(not pinocchio_ref.available(), reason="pinocchio is not importable here (Tier-A cross-check needs it)")

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def models():
    model, cmodel, _ = make_panda()
    pm, pcm = pinocchio_ref.to_pinocchio(model, cmodel)
    return model, cmodel, pm, pcm


def _states(n=2000):
    g = np.load(os.path.join(GOLDEN, "panda_states.npz"))
    return np.concatenate([np.stack([COLLISION_STATE, FREE_STATE_1, FREE_STATE_2]), np.asarray(g["q"], dtype=np.float64)[:n]])


def test_adapter_on_real_pinocchio_objects(models):
    from pyroboplan_b200.engine import flatten_model
    from pyroboplan_b200.model.pinocchio_adapter import from_pinocchio

    model, cmodel, pm, pcm = models
    m2, c2 = from_pinocchio(pm, pcm)
    a, sa = flatten_model(model, cmodel)
    b, sb = flatten_model(m2, c2)
    assert sa == sb
    for k in a:
        np.testing.assert_allclose(a[k], b[k], rtol=0, atol=1e-7, err_msg=k)


@pytest.mark.parametrize("padding", [0.0, 0.02])
def test_oracle_verdicts_equal_pinocchio(models, padding):
    from oracle.oracle import Oracle

    model, cmodel, pm, pcm = models
    Q = _states()
    ref = pinocchio_ref.check_states(pm, pcm, Q, padding)
    hit, dist, _, _ = Oracle(model, cmodel).check_states(Q, padding=padding, want_all=True, use_cull=False)
    bad = np.nonzero(ref != hit.astype(bool))[0]
    assert all(abs(dist[i] - padding) <= 1e-6 for i in bad), f"oracle and Pinocchio disagree away from the boundary at {bad[:5]}"


def test_oracle_min_distances_equal_pinocchio(models):
    from oracle.oracle import Oracle

    model, cmodel, pm, pcm = models
    Q = _states(500)
    ref = pinocchio_ref.min_distances(pm, pcm, Q)
    _, dist, _, _ = Oracle(model, cmodel).check_states(Q, want_all=True, use_cull=False)
    free = ref > 1e-6
    assert np.abs(ref[free] - dist[free]).max() < 1e-5


@pytest.mark.gpu
def test_dropin_functions_accept_pinocchio_objects(models):
    from pyroboplan_b200.core.utils import check_collisions_at_state, check_collisions_at_states, get_minimum_distance_at_state

    _, _, pm, pcm = models
    Q = _states(1000)
    ref = pinocchio_ref.check_states(pm, pcm, Q)
    refd = pinocchio_ref.min_distances(pm, pcm, Q)
    hit = np.asarray(check_collisions_at_states(pm, pcm, Q))
    bad = np.nonzero(ref != hit)[0]
    assert all(abs(refd[i]) <= 1e-6 for i in bad)
    assert bool(check_collisions_at_state(pm, pcm, COLLISION_STATE)) and not bool(check_collisions_at_state(pm, pcm, FREE_STATE_1))
    assert abs(get_minimum_distance_at_state(pm, pcm, FREE_STATE_1) - refd[1]) < 1e-5
