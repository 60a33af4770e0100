"""The other bundled models through the same engine: the marble labyrinth
(/root/reference/src/pyroboplan/models/marble.py: sphere vs ten boxes, two prismatic joints) and
the sphere-approximated Panda (panda_spheres.urdf, models/panda.py:13-27: 17 link spheres).
Verdicts, distances and edges against the float64 oracle."""

import numpy as np
import pytest

from conftest import uniform_states

pytestmark = pytest.mark.gpu


def make_marble():
    from pyroboplan_b200.models import marble

    m, c, v = marble.load_models()
    marble.add_object_collisions(m, c, v)
    return m, c


def make_sphere_panda():
    from pyroboplan_b200.models import panda

    m, c, v = panda.load_models(use_sphere_collisions=True)
    panda.add_self_collisions(m, c)
    panda.add_object_collisions(m, c, v)
    return m, c


@pytest.mark.parametrize("which,n", [("marble", 20000), ("sphere_panda", 100000)])
def test_states_and_distances(which, n):
    from oracle.oracle import Oracle
    from pyroboplan_b200.engine import CollisionEngine

    model, cmodel = make_marble() if which == "marble" else make_sphere_panda()
    eng = CollisionEngine(model, cmodel, device=0)
    orc = Oracle(model, cmodel)
    q = uniform_states(model, n, 7)
    hit_ref, md_ref, _, _ = orc.check_states(q.astype(np.float64), want_all=True, use_cull=False)
    hit = eng.check_states(q)
    bad = np.nonzero(hit != hit_ref.astype(bool))[0]
    assert all(abs(md_ref[i]) <= 1e-6 for i in bad)
    assert 0.02 < hit_ref.mean() < 0.98
    hit2, dist = eng.check_states(q, return_distance=True)
    free = md_ref > 1e-6
    assert np.abs(dist[free] - md_ref[free]).max() < 1e-5
    # padding moves the verdict boundary, not the arithmetic
    hp_ref, _, _, _ = orc.check_states(q[:20000].astype(np.float64), padding=0.02)
    hp = eng.check_states(q[:20000], padding=0.02)
    badp = np.nonzero(hp != hp_ref.astype(bool))[0]
    assert all(abs(md_ref[i] - 0.02) <= 2e-6 for i in badp)


def test_marble_edges_and_sampler():
    from oracle.oracle import Oracle
    from pyroboplan_b200.core.utils import check_collisions_at_state, get_random_collision_free_state
    from pyroboplan_b200.engine import get_engine

    model, cmodel = make_marble()
    eng = get_engine(model, cmodel)
    orc = Oracle(model, cmodel)
    a, b = uniform_states(model, 4000, 1), uniform_states(model, 4000, 2)
    b[:2000] = a[:2000] + 0.1 * (b[:2000] - a[:2000])
    ref, _, _, _ = orc.check_edges(a.astype(np.float64), b.astype(np.float64), 0.01)
    got = eng.check_edges(a, b, 0.01)
    assert (got == ref.astype(bool)).all() and 0.1 < ref.mean() < 0.95
    np.random.seed(0)
    q = get_random_collision_free_state(model, cmodel)
    assert q is not None and not check_collisions_at_state(model, cmodel, q)
    free = eng.sample_free_states(5000, seed=1)
    assert free.shape == (5000, 2) and not orc.check_states(free.cpu().numpy().astype(np.float64))[0].any()
