"""Random kinematic trees with random primitive / hull geometry: the engine is model-agnostic (the
broadphase is GENERATED from the model), so its verdicts and distances are compared with the
float64 oracle on models nobody tuned it for -- generic joint axes, branching trees, prismatic
joints, moving boxes / cylinders / capsules, static-vs-static pairs, random convex hulls."""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def random_rotation(rng):
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def random_model(seed):
    from pyroboplan_b200.model import (Box, Capsule, CollisionPair, Convex, Cylinder, GeometryModel, GeometryObject, Model, SE3, Sphere,
                                       JOINT_PRISMATIC, JOINT_REVOLUTE)

    rng = np.random.default_rng(seed)
    model = Model(f"random{seed}")
    cm = GeometryModel()
    njoints = int(rng.integers(3, 8))
    for j in range(njoints):
        parent = 0 if j == 0 else int(rng.integers(max(0, j - 2), j + 1))   # chains with the odd branch
        prismatic = rng.random() < 0.25
        axis = np.eye(3)[rng.integers(0, 3)] * rng.choice([-1.0, 1.0]) if rng.random() < 0.5 else rng.normal(size=3)
        placement = SE3(random_rotation(rng) if rng.random() < 0.7 else np.eye(3), rng.uniform(-0.3, 0.3, size=3))
        lo, hi = (-0.3, 0.3) if prismatic else (-2.5, 2.5)
        model.addJoint(parent, JOINT_PRISMATIC if prismatic else JOINT_REVOLUTE, axis, placement, f"j{j}", lo, hi)

    def random_shape():
        kind = rng.integers(0, 5)
        if kind == 0:
            return Sphere(rng.uniform(0.04, 0.12))
        if kind == 1:
            return Box(*rng.uniform(0.06, 0.3, size=3))
        if kind == 2:
            return Cylinder(rng.uniform(0.03, 0.1), rng.uniform(0.1, 0.4))
        if kind == 3:
            return Capsule(rng.uniform(0.03, 0.08), rng.uniform(0.1, 0.3))
        pts = rng.normal(size=(int(rng.integers(8, 40)), 3)) * rng.uniform(0.03, 0.15, size=3)
        return Convex(pts)

    moving = []
    for j in range(1, njoints + 1):
        for _ in range(int(rng.integers(1, 3))):
            obj = GeometryObject(f"g{len(cm.geometryObjects)}", j, SE3(random_rotation(rng), rng.uniform(-0.1, 0.1, size=3)), random_shape())
            moving.append(cm.addGeometryObject(obj))
    static = []
    for _ in range(int(rng.integers(2, 5))):
        obj = GeometryObject(f"s{len(cm.geometryObjects)}", 0, SE3(random_rotation(rng), rng.uniform(-0.6, 0.6, size=3)), random_shape())
        static.append(cm.addGeometryObject(obj))
    geoms = cm.geometryObjects
    for a in static:
        for b in moving:
            if rng.random() < 0.8:
                cm.addCollisionPair(CollisionPair(a, b))
    for i, a in enumerate(moving):
        for b in moving[i + 1:]:
            if geoms[a].parentJoint != geoms[b].parentJoint and rng.random() < 0.4:
                cm.addCollisionPair(CollisionPair(a, b))
    if len(static) >= 2:
        # a static-vs-static pair (the same verdict for every state): kept only when the two are apart,
        # otherwise every state of the model would trivially collide
        from oracle.oracle import Oracle

        pair = CollisionPair(static[0], static[1])
        cm.addCollisionPair(pair)
        if Oracle(model, cm).pair_distance(np.zeros(model.nq), static[0], static[1])[0] < 0.02:
            cm.removeCollisionPair(pair)
    return model, cm


@pytest.mark.parametrize("seed", list(range(1, 15)))
def test_random_model_matches_oracle(seed):
    from oracle.oracle import Oracle
    from pyroboplan_b200.engine import CollisionEngine

    model, cm = random_model(seed)
    eng = CollisionEngine(model, cm, device=0)
    orc = Oracle(model, cm)
    rng = np.random.default_rng(100 + seed)
    q = rng.uniform(model.lowerPositionLimit, model.upperPositionLimit, size=(12000, model.nq)).astype(np.float32)
    hit_ref, md_ref, _, _ = orc.check_states(q.astype(np.float64), want_all=True, use_cull=False)
    for name, got in (("pipeline", eng.check_states(q)), ("fused", eng.check_states(q[:300]))):
        bad = np.nonzero(got != hit_ref[: len(got)].astype(bool))[0]
        assert all(abs(md_ref[i]) <= 1e-6 for i in bad), (name, seed, bad[:5], md_ref[bad[:5]])
    _, dist = eng.check_states(q, return_distance=True)
    free = md_ref > 1e-6
    if free.any():
        assert np.abs(dist[free] - md_ref[free]).max() < 1e-5
    # edges through the same model
    a, b = q[:600], q[600:1200]
    ref, _, _, _ = orc.check_edges(a.astype(np.float64), b.astype(np.float64), 0.05)
    assert (eng.check_edges(a, b, 0.05) == ref.astype(bool)).all()
    eng.close()


@pytest.mark.parametrize("seed", [21, 22, 23, 24])
def test_random_model_signed_distances_and_witnesses(seed):
    """Per-pair signed distances (EPA for overlaps), witness points and normals on random models."""
    from oracle.oracle import Oracle
    from pyroboplan_b200.engine import CollisionEngine

    model, cm = random_model(seed)
    eng = CollisionEngine(model, cm, device=0)
    orc = Oracle(model, cm)
    rng = np.random.default_rng(200 + seed)
    q = rng.uniform(model.lowerPositionLimit, model.upperPositionLimit, size=(150, model.nq)).astype(np.float32)
    dist, p1, p2, nrm = eng.compute_distances(q)
    curved = [type(g.geometry).__name__ in ("Cylinder",) for g in cm.geometryObjects]
    checked = deep = 0
    for i in range(len(q)):
        _, oMg = orc.fk(q[i].astype(np.float64))
        for k, cp in enumerate(cm.collisionPairs):
            ds, pa, pb, n, exact = orc.pair_signed_distance_placed(oMg, cp.first, cp.second)
            if not exact:
                continue
            tol = 2e-3 if (curved[cp.first] or curved[cp.second]) and ds < 0 else 3e-5   # 40-vertex EPA polytope on curved supports
            assert abs(dist[i, k] - ds) < tol, (seed, i, k, dist[i, k], ds)
            if abs(ds) > 1e-4:
                np.testing.assert_allclose(p2[i, k].astype(np.float64) - p1[i, k], nrm[i, k].astype(np.float64) * dist[i, k], atol=2e-5)
            checked += 1
            deep += int(ds < -1e-4)
    assert checked > 1000 and deep > 5
    eng.close()


@pytest.mark.parametrize("seed", [31, 32, 33])
def test_random_tree_ik_iteration(seed):
    """One damped-least-squares iteration on random trees (generic axes, prismatic joints, branches)."""
    import torch

    from oracle import ik_oracle as ik
    from pyroboplan_b200.engine import CollisionEngine
    from pyroboplan_b200.model import Frame, FRAME_OP, SE3

    model, cm = random_model(seed)
    rng = np.random.default_rng(300 + seed)
    fid = model.addFrame(Frame("tool", model.njoints - 1, SE3(random_rotation(rng), rng.uniform(-0.1, 0.1, size=3)), FRAME_OP))
    eng = CollisionEngine(model, cm, device=0)
    n = 100
    goal = rng.uniform(model.lowerPositionLimit, model.upperPositionLimit, size=(n, model.nq))
    init = rng.uniform(model.lowerPositionLimit, model.upperPositionLimit, size=(n, model.nq))
    targets = np.array([np.concatenate([T[:3, :3].reshape(9), T[:3, 3]]) for T in (ik.frame_pose(model, fid, g)[0] for g in goal)])
    Q1, status, iters, e0 = eng.ik_iterate(fid, torch.as_tensor(targets, device="cuda:0"), torch.as_tensor(init, device="cuda:0"), max_iters=1)
    Q1 = Q1.cpu().numpy()
    for i in range(n):
        T = np.eye(4)
        T[:3, :3] = targets[i, :9].reshape(3, 3)
        T[:3, 3] = targets[i, 9:]
        q_ref, conv, _, it, _ = ik.ik_try(model, fid, T, init[i], max_iters=1)
        if conv:
            continue
        step = np.abs(q_ref - init[i]).max()
        np.testing.assert_allclose(Q1[i], q_ref, atol=1e-9 * max(1.0, step / 1e-3))
    eng.close()
