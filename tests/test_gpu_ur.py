"""UR5 / UR10 on their base (reference models/ur5.py:9-47, ur10.py:9-47): NON-convex collision
meshes.  The engine brackets each between its hull and an inner convex cell of its triangle planes (far apart here:
up to 0.67 m), so almost every near item is settled by the triangle stage -- against the oracle's exact
triangle tests: identical verdicts outside the 1e-6 m band, min distances within 1e-5 m."""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _ur(which, obstacles=True):
    from pyroboplan_b200.core.utils import set_collisions
    from pyroboplan_b200.model import Box, GeometryObject, SE3, Sphere
    from pyroboplan_b200.models import ur5, ur10

    if which == 5:
        m, c, v = ur5.load_ur5_on_base_models()
        ur5.add_ur5_on_base_self_collisions(m, c)
    else:
        m, c, v = ur10.load_ur10_on_base_models()
        ur10.add_ur10_on_base_self_collisions(m, c)
    if obstacles:
        names = [g.name for g in c.geometryObjects if "link" in g.name]
        for name, xyz, shape in (("ball", [0.6, 0.3, 0.7], Sphere(0.2)), ("crate", [-0.4, 0.6, 0.9], Box(0.3, 0.4, 0.3))):
            c.addGeometryObject(GeometryObject(name, 0, SE3(np.eye(3), np.array(xyz)), shape))
            for n in names:
                set_collisions(m, c, name, n, True)
    return m, c


def _states(m, n, seed):
    lo, hi = np.maximum(m.lowerPositionLimit, -3.2), np.minimum(m.upperPositionLimit, 3.2)
    return np.random.default_rng(seed).uniform(lo, hi, size=(n, m.nq)).astype(np.float32)


@pytest.mark.parametrize("which", [5, 10])
def test_ur_states_match_the_triangle_oracle(which):
    import torch

    from oracle.oracle import Oracle
    from pyroboplan_b200.engine import CollisionEngine

    m, c = _ur(which)
    eng, orc = CollisionEngine(m, c, device=0), Oracle(m, c)
    q = _states(m, 20000, which)
    hit_ref, d_ref, _, _ = orc.check_states(q.astype(np.float64), want_all=True, use_cull=False)
    qd = torch.as_tensor(q, device="cuda:0")
    for name, hit in (("pipeline", eng.check_states(qd)), ("small batch", eng.check_states(qd[:300])), ("distance mode", eng.check_states(qd, return_distance=True)[0])):
        h = hit.cpu().numpy()
        bad = np.nonzero(h != hit_ref[: len(h)].astype(bool))[0]
        assert all(abs(d_ref[i]) <= 1e-6 for i in bad), f"UR{which} {name}: verdicts differ at {bad[:5]} (oracle distances {d_ref[bad[:5]]})"
    _, dist = eng.check_states(qd, return_distance=True)
    free = d_ref > 1e-6
    assert np.abs(dist.cpu().numpy()[free] - d_ref[free]).max() < 1e-5
    assert 0.2 < hit_ref.mean() < 0.9
    # with padding
    hp, dp, _, _ = orc.check_states(q[:5000].astype(np.float64), padding=0.02, want_all=True, use_cull=False)
    h = eng.check_states(qd[:5000], padding=0.02).cpu().numpy()
    bad = np.nonzero(h != hp.astype(bool))[0]
    assert all(abs(dp[i] - 0.02) <= 1e-6 for i in bad)


def test_ur5_edges_and_dropin_calls():
    from oracle.oracle import Oracle
    from pyroboplan_b200.core.utils import check_collisions_at_state, get_minimum_distance_at_state
    from pyroboplan_b200.planning.utils import has_collision_free_paths

    m, c = _ur(5)
    orc = Oracle(m, c)
    a, b = _states(m, 400, 1), _states(m, 400, 2)
    b = a + np.float32(0.15) * (b - a)
    ref = orc.check_edges(a.astype(np.float64), b.astype(np.float64), 0.05)[0].astype(bool)
    assert (np.asarray(has_collision_free_paths(a, b, 0.05, m, c)) == ~ref).all()
    hit_ref, d_ref, _, _ = orc.check_states(a[:40].astype(np.float64), want_all=True, use_cull=False)
    for q, h, d in zip(a[:40], hit_ref, d_ref):
        if abs(d) > 1e-6:
            assert bool(check_collisions_at_state(m, c, q)) == bool(h)
        if d > 1e-6:
            assert abs(get_minimum_distance_at_state(m, c, q) - d) < 1e-5
