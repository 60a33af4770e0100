"""Pins the float64 oracle against every known-answer vector the reference's tests hold for the
hot path and verifies its pair results with GJK-independent certificates."""

import numpy as np
import pytest

from conftest import COLLISION_STATE, FREE_STATE_1, FREE_STATE_2, make_panda, make_two_dof, uniform_states
from oracle.oracle import Oracle, discretize_joint_space_path


def test_reference_known_answer_states(panda_self):
    # reference test/core/test_utils.py:125-158 (self-collision pairs only)
    orc = Oracle(panda_self[0], panda_self[1])
    assert orc.check_state(COLLISION_STATE)
    assert not orc.check_state(FREE_STATE_1)
    assert not orc.check_state(FREE_STATE_2)
    # which pair / how far: link1-link5 collide; free-state clearances (SURVEY.md 8c probe values)
    hit, md, fp, _ = orc.check_states(np.array([COLLISION_STATE, FREE_STATE_1, FREE_STATE_2]), want_all=True, use_cull=False)
    pair = panda_self[1].collisionPairs[fp[0]]
    assert (pair.first, pair.second) == (1, 5)
    assert md[1] == pytest.approx(0.1024, abs=1e-4)
    assert md[2] == pytest.approx(0.1232, abs=1e-4)


def test_reference_known_answer_paths(panda_self):
    # reference test/core/test_utils.py:161-212: three explicit states, not interpolated
    orc = Oracle(panda_self[0], panda_self[1])
    assert orc.check_path([FREE_STATE_1, COLLISION_STATE, FREE_STATE_2])
    assert not orc.check_path([FREE_STATE_1, FREE_STATE_2, FREE_STATE_1])


def test_reference_discretize_vectors():
    # reference test/planning/test_planning_utils.py:26-43
    path = discretize_joint_space_path([np.array([0.0, 0.0]), np.array([2.0, 0.0])], 1.0)
    np.testing.assert_array_equal(np.array(path), [[0.0, 0.0], [1.0, 0.0], [2.0, 0.0]])
    same = discretize_joint_space_path([np.array([1.0, 2.0]), np.array([1.0, 2.0])], 0.1)
    assert len(same) == 1
    np.testing.assert_array_equal(same[0], [1.0, 2.0])


def test_discretize_matches_numpy_formula():
    rng = np.random.default_rng(3)
    for _ in range(50):
        a, b = rng.uniform(-3, 3, size=9), rng.uniform(-3, 3, size=9)
        step = rng.choice([0.05, 0.1, 0.37])
        ours = np.array(discretize_joint_space_path([a, b], step))
        n = int(np.ceil(np.linalg.norm(b - a) / step)) + 1  # reference planning/utils.py:137-139
        ref = np.array([a + s * (b - a) for s in np.linspace(0.0, 1.0, n)])
        assert ours.shape == ref.shape
        np.testing.assert_array_equal(ours, ref)


def test_zero_pairs_is_free():
    # reference test/planning/test_prm.py loads the 2-DOF model without obstacles: zero pairs
    model, cmodel, _ = make_two_dof(objects=False)
    orc = Oracle(model, cmodel)
    hit, md, fp, _ = orc.check_states(np.zeros((3, 2)))
    assert not hit.any() and np.all(np.isinf(md)) and np.all(fp == -1)
    assert orc.min_distance(np.zeros(2)) == np.inf


def test_cull_modes_do_not_change_verdicts(panda_full, panda_oracle):
    q = uniform_states(panda_full[0], 3000, 11).astype(np.float64)
    h0, _, fp0, c0 = panda_oracle.check_states(q, use_cull=0)
    h1, _, fp1, c1 = panda_oracle.check_states(q, use_cull=1)
    h2, _, _, c2 = panda_oracle.check_states(q, use_cull=2)
    assert (h0 == h1).all() and (h0 == h2).all() and (fp0 == fp1).all()
    assert c2["pair_tests"] < c1["pair_tests"] < c0["pair_tests"]
    hp0, _, _, _ = panda_oracle.check_states(q, padding=0.02, use_cull=0)
    hp1, _, _, _ = panda_oracle.check_states(q, padding=0.02, use_cull=1)
    hp2, _, _, _ = panda_oracle.check_states(q, padding=0.02, use_cull=2)
    assert (hp0 == hp1).all() and (hp0 == hp2).all() and hp0.sum() > h0.sum()


def _support_extent(geom, T, direction):
    """max over the (solid) shape of <x, direction> in the world frame -- independent of GJK."""
    R, t = T[:9].reshape(3, 3), T[9:]
    d = R.T @ direction
    s = geom.geometry
    name = type(s).__name__
    if name == "Convex":
        local = (s.points @ d).max()
    elif name == "Box":
        local = np.abs(d) @ s.halfSide
    elif name == "Sphere":
        local = s.radius * np.linalg.norm(d)
    elif name == "Cylinder":
        local = s.radius * np.hypot(d[0], d[1]) + s.halfLength * abs(d[2])
    else:
        local = s.radius * np.linalg.norm(d) + s.halfLength * abs(d[2])
    return local + t @ direction


def _contains(geom, T, p, tol=1e-9):
    R, t = T[:9].reshape(3, 3), T[9:]
    x = R.T @ (p - t)
    s = geom.geometry
    name = type(s).__name__
    if name == "Convex":
        return bool(np.all(s.planes[:, :3] @ x + s.planes[:, 3] <= tol))
    if name == "Box":
        return bool(np.all(np.abs(x) <= s.halfSide + tol))
    if name == "Sphere":
        return bool(np.linalg.norm(x) <= s.radius + tol)
    if name == "Cylinder":
        return bool(np.hypot(x[0], x[1]) <= s.radius + tol and abs(x[2]) <= s.halfLength + tol)
    zc = np.clip(x[2], -s.halfLength, s.halfLength)
    return bool(np.linalg.norm(x - [0, 0, zc]) <= s.radius + tol)


@pytest.mark.parametrize("which", ["panda", "two_dof"])
def test_pair_distance_certificates(which):
    """Every oracle pair result is PROVED by data that does not involve GJK:
    separated -> both witnesses lie in their shapes (upper bound) and the plane through them
    separates the shapes by exactly the reported distance (lower bound);
    overlapping -> the reported common point lies in both shapes."""
    model, cmodel, _ = make_panda() if which == "panda" else make_two_dof()
    orc = Oracle(model, cmodel)
    geoms = cmodel.geometryObjects
    q = uniform_states(model, 40 if which == "panda" else 400, 5).astype(np.float64)
    checked = 0
    for qi in q:
        _, oMg = orc.fk(qi)
        for p in cmodel.collisionPairs:
            a, b = p.first, p.second
            d, pa, pb, inter = orc.pair_distance(qi, a, b)
            if inter:
                assert d == 0.0
                assert _contains(geoms[a], oMg[a], pa, 1e-8) and _contains(geoms[b], oMg[b], pb, 1e-8)
                np.testing.assert_allclose(pa, pb, atol=1e-12)
            else:
                assert d > 0.0
                assert _contains(geoms[a], oMg[a], pa, 1e-8) and _contains(geoms[b], oMg[b], pb, 1e-8)
                n = (pb - pa) / np.linalg.norm(pb - pa)
                assert np.linalg.norm(pb - pa) == pytest.approx(d, abs=1e-9)
                gap = -_support_extent(geoms[b], oMg[b], -n) - _support_extent(geoms[a], oMg[a], n)
                assert gap == pytest.approx(d, abs=1e-8)  # separating slab is exactly d wide
            checked += 1
    assert checked > 1000


def test_fk_against_independent_composition(panda_full, panda_oracle):
    """Oracle FK vs a direct product of homogeneous matrices built from the URDF numbers
    (SURVEY.md Appendix A.1), including the fixed hand offset and the prismatic fingers."""
    from pyroboplan_b200.model import rpy_to_matrix

    def H(xyz, rpy):
        m = np.eye(4)
        m[:3, :3] = rpy_to_matrix(*rpy)
        m[:3, 3] = xyz
        return m

    def Rz(a):
        m = np.eye(4)
        m[:2, :2] = [[np.cos(a), -np.sin(a)], [np.sin(a), np.cos(a)]]
        return m

    def Ty(d):
        m = np.eye(4)
        m[1, 3] = d
        return m

    hp = 1.57079632679
    origins = [([0, 0, 0.333], [0, 0, 0]), ([0, 0, 0], [-hp, 0, 0]), ([0, -0.316, 0], [hp, 0, 0]), ([0.0825, 0, 0], [hp, 0, 0]),
               ([-0.0825, 0.384, 0], [-hp, 0, 0]), ([0, 0, 0], [hp, 0, 0]), ([0.088, 0, 0], [hp, 0, 0])]
    rng = np.random.default_rng(2)
    for _ in range(20):
        q = rng.uniform(panda_full[0].lowerPositionLimit, panda_full[0].upperPositionLimit)
        T = np.eye(4)
        link = [T.copy()]
        for (xyz, rpy), qi in zip(origins, q[:7]):
            T = T @ H(xyz, rpy) @ Rz(qi)
            link.append(T.copy())
        hand = link[7] @ H([0, 0, 0.107], [0, 0, 0]) @ H([0, 0, 0], [0, 0, -0.785398163397])
        left = hand @ H([0, 0, 0.0584], [0, 0, 0]) @ Ty(q[7])
        right = hand @ H([0, 0, 0.0584], [0, 0, 0]) @ Ty(-q[8]) @ H([0, 0, 0], [0, 0, 3.14159265359])
        _, oMg = panda_oracle.fk(q)
        expected = link + [hand, left, right]
        for g, E in enumerate(expected):
            got = np.eye(4)
            got[:3, :3] = oMg[g, :9].reshape(3, 3)
            got[:3, 3] = oMg[g, 9:]
            np.testing.assert_allclose(got, E, atol=1e-12)


def test_edge_semantics(panda_full, panda_oracle):
    """has_collision_free_path restated: q2 first, then the discretised samples in order."""
    model = panda_full[0]
    q = uniform_states(model, 400, 21).astype(np.float64)
    a, b = q[:200], q[:200] + 0.2 * (q[200:] - q[:200])
    hit, fb, evals, _ = panda_oracle.check_edges(a, b, 0.05)
    for e in range(0, 200, 7):
        path = np.array(discretize_joint_space_path([a[e], b[e]], 0.05))
        per_state, _, _, _ = panda_oracle.check_states(path)
        assert bool(hit[e]) == bool(per_state.any())
        if per_state[-1]:
            assert fb[e] == -2  # the destination itself collides
        elif per_state.any():
            assert fb[e] == int(np.argmax(per_state))
        else:
            assert fb[e] == -1


@pytest.mark.parametrize("which", ["panda", "two_dof"])
def test_penetration_depth_certificates(which):
    """The oracle's EPA (negative distances, coal's convention) is PROVED without EPA or GJK:
    the support function of A - B along the reported normal equals the depth (so translating B by
    depth * normal makes the shapes touch), no sampled direction gives a smaller value (the depth is
    the minimum translation), the deepest points lie in their shapes and differ by normal * distance,
    and separated pairs agree with the unsigned distance."""
    model, cmodel, _ = make_panda() if which == "panda" else make_two_dof()
    orc = Oracle(model, cmodel)
    geoms = cmodel.geometryObjects
    rng = np.random.default_rng(9)
    q = uniform_states(model, 150 if which == "panda" else 1500, 77).astype(np.float64)
    dirs = rng.normal(size=(400, 3))
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    penetrating = 0
    for qi in q:
        _, oMg = orc.fk(qi)
        for p in cmodel.collisionPairs:
            a, b = p.first, p.second
            d0, _, _, inter = orc.pair_distance(qi, a, b)
            d, pa, pb, n, exact = orc.pair_signed_distance(qi, a, b)
            if not inter:
                assert d == pytest.approx(d0, abs=1e-12) and d > 0
                continue
            if not exact:
                continue
            penetrating += 1
            assert d <= 0.0 and np.linalg.norm(n) == pytest.approx(1.0, abs=1e-9)
            tol = 1e-7 if which == "panda" else 2e-6       # the 2-DOF cylinder is curved: EPA converges asymptotically
            width = _support_extent(geoms[a], oMg[a], n) + _support_extent(geoms[b], oMg[b], -n)
            assert width == pytest.approx(-d, abs=tol)                     # depth along the normal
            others = [_support_extent(geoms[a], oMg[a], u) + _support_extent(geoms[b], oMg[b], -u) for u in dirs[:60]]
            assert min(others) >= -d - tol                                   # and no direction does better
            assert _contains(geoms[a], oMg[a], pa, 1e-6) and _contains(geoms[b], oMg[b], pb, 1e-6)
            np.testing.assert_allclose(pb - pa, n * d, atol=1e-6)            # p2 - p1 = normal * signed distance
    assert penetrating > 30


def test_curved_pairs_are_stable_under_tiny_perturbations():
    """Regression: cylinder-cylinder GJK produces sliver tetrahedra near convergence whose inside test
    is rounding noise; the oracle once reported a pair 4.3 cm apart as intersecting for one state and
    not for the same state perturbed by 1e-8 (found by tests/test_gpu_random_models.py).  A claimed
    enclosure that contradicts a positive separating-plane bound is now discarded."""
    from test_gpu_random_models import random_model

    model, cm = random_model(3)
    orc = Oracle(model, cm)
    q = np.array([-0.08408726, 0.8007079, 0.00525507, -1.2401325, -0.89210567, -0.01261778, 1.05439098])
    d0 = orc.pair_distance(q, 3, 5)[0]
    assert d0 == pytest.approx(0.0431242, abs=1e-6)
    rng = np.random.default_rng(0)
    for _ in range(200):
        d = orc.pair_distance(q + rng.normal(scale=1e-8, size=q.shape), 3, 5)[0]
        assert d == pytest.approx(d0, abs=1e-6)
    # and broadly: distances of curved pairs vary continuously along a fine path
    qs = q + np.linspace(0.0, 1e-3, 400)[:, None] * rng.normal(size=q.shape)
    for a, b in [(p.first, p.second) for p in cm.collisionPairs][:40]:
        ds = np.array([orc.pair_distance(x, a, b)[0] for x in qs])
        assert np.abs(np.diff(ds)).max() < 1e-4


def test_penetration_depth_from_a_planar_gjk_simplex():
    """Regression: a sphere centre inside a cylinder keeps every GJK direction in the axial plane through
    the centre, so GJK ends on a TRIANGLE through the origin, not on a tetrahedron.  The oracle once
    read that as touching cores (depth = sphere radius only; found by tests/test_gpu_random_models.py,
    where the device's bipyramid start was right).  The analytic depth of a point inside a cylinder is
    min(R - rho, h/2 - |z|); the sphere radius adds to it."""
    from test_gpu_random_models import random_model

    model, cm = random_model(22)
    orc = Oracle(model, cm)
    geoms = cm.geometryObjects
    pairs = [(p.first, p.second) for p in cm.collisionPairs
             if {type(geoms[p.first].geometry).__name__, type(geoms[p.second].geometry).__name__} == {"Sphere", "Cylinder"}]
    rng = np.random.default_rng(222)
    q = rng.uniform(model.lowerPositionLimit, model.upperPositionLimit, size=(300, model.nq))
    placements = [orc.fk(x)[1] for x in q]
    inside = 0
    for pair in pairs:
        a, b = pair if type(geoms[pair[0]].geometry).__name__ == "Sphere" else pair[::-1]
        r, R, hl = geoms[a].geometry.radius, geoms[b].geometry.radius, geoms[b].geometry.halfLength
        for oMg in placements:
            Rb, tb = oMg[b][:9].reshape(3, 3), oMg[b][9:]
            c = Rb.T @ (oMg[a][9:] - tb)            # sphere centre in the cylinder's frame
            rho = np.hypot(c[0], c[1])
            if rho >= R or abs(c[2]) >= hl:
                continue
            ds, _, _, _, exact = orc.pair_signed_distance_placed(oMg, *pair)
            assert exact
            assert ds == pytest.approx(-(min(R - rho, hl - abs(c[2])) + r), abs=1e-4)   # 320-vertex polytope on a curved support
            inside += 1
    assert inside > 20


# ------------------------------------------------------------------ surface (triangle-soup) semantics
def _tri_dist_sampled(a, b, n=60):
    """Brute-force lower-resolution check: minimum distance between dense barycentric samples."""
    u = np.linspace(0.0, 1.0, n)
    w = np.array([(x, y, 1.0 - x - y) for x in u for y in u if x + y <= 1.0 + 1e-12])
    pa, pb = w @ a, w @ b
    return np.sqrt(((pa[:, None, :] - pb[None, :, :]) ** 2).sum(axis=2).min())


def test_triangle_distance_known_answers(panda_full):
    """The oracle's exact triangle-triangle distance (the kernel of the surface semantics): analytic cases
    and a sampled cross-check on random triangles."""
    model, cmodel = panda_full[0], panda_full[1]
    orc = Oracle(model, cmodel)
    eye = np.tile(np.concatenate([np.eye(3).reshape(9), np.zeros(3)]), (len(cmodel.geometryObjects), 1))

    def dist(a, b):
        orc._tris = np.ascontiguousarray(np.stack([np.asarray(a, float).reshape(9), np.asarray(b, float).reshape(9)]))
        orc._tri_off, orc._tri_cnt = np.zeros(len(eye), np.int32), np.zeros(len(eye), np.int32)
        orc._tri_off[1], orc._tri_cnt[0], orc._tri_cnt[1] = 1, 1, 1
        return orc.soup_pair_distance(eye, 0, 1, brute=True)

    base = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], float)
    assert dist(base, base + [0, 0, 0.25]) == pytest.approx(0.25)                       # parallel faces
    assert dist(base, base + [2.0, 0, 0]) == pytest.approx(1.0)                         # vertex - vertex
    assert dist(base, np.array([[0.2, 0.2, -1], [0.2, 0.2, 1], [5, 5, 1]], float)) == 0.0   # an edge pierces the face
    assert dist(base, np.array([[0.2, 0.2, 0.5], [0.3, 0.2, 1], [0.2, 0.3, 1]], float)) == pytest.approx(0.5)   # vertex - face
    assert dist(base, np.array([[0.5, -1, 0.3], [0.5, 2, 0.3], [0.5, 0.5, 4]], float)) == pytest.approx(0.3)    # edge - face interior / edge
    assert dist(base, base * 0.5 + [0.1, 0.1, 0.0]) == 0.0                              # coplanar, overlapping
    rng = np.random.default_rng(5)
    for _ in range(60):
        a, b = rng.normal(size=(3, 3)), rng.normal(size=(3, 3)) + rng.normal(size=3)
        d = dist(a, b)
        s = _tri_dist_sampled(a, b)
        assert d <= s + 1e-12 and s - d < 0.06          # the samples can only overestimate, by about one sample spacing


def test_surface_semantics_differ_from_hulls_by_enclosure(panda_full):
    """Pinocchio's meshes are surfaces (coal BVHModel): a finger wholly inside link 5 / 0 / 1 / 2 -- the hand
    pair is disabled by the SRDF -- is free for the reference and a collision for convex hulls.  The
    oracle's surface mode must (1) never report a hit the hulls do not, (2) agree with brute force over
    all triangle pairs, (3) flip about 1 in 1000 uniform states, all of them through deeply overlapping
    hulls (enclosure), not through the sub-millimetre dents."""
    model, cmodel = panda_full[0], panda_full[1]
    surf, hull = Oracle(model, cmodel), Oracle(model, cmodel, mesh_semantics="hull")
    assert surf.mesh_semantics == "surface" and hull.mesh_semantics == "hull"
    q = uniform_states(model, 60000, 0).astype(np.float64)
    hs = surf.check_states(q)[0].astype(bool)
    hh = hull.check_states(q)[0].astype(bool)
    assert not (hs & ~hh).any()
    flipped = np.nonzero(hh & ~hs)[0]
    assert 20 <= len(flipped) <= 120
    geoms = cmodel.geometryObjects
    for i in flipped[:25]:
        _, oMg = surf.fk(q[i])
        deep = 0
        for p in cmodel.collisionPairs:
            if hull.pair_distance_placed(oMg, p.first, p.second)[0] > 0.0:
                continue
            exact = surf.soup_pair_distance(oMg, p.first, p.second, brute=True)
            assert exact > 0.0                                        # no surface contact on any pair
            assert surf.soup_pair_distance(oMg, p.first, p.second) == pytest.approx(exact, abs=1e-12)      # box rejection is exact
            if hull.pair_signed_distance_placed(oMg, p.first, p.second)[0] < -5e-3:
                deep += 1
                assert "finger" in geoms[p.second].name or "link7" in geoms[p.second].name
        assert deep >= 1
    # min distance in surface mode: the clearance of the enclosed shape, positive
    md = surf.check_states(q[flipped[:25]], want_all=True, use_cull=False)[1]
    assert (md > 0).all()
