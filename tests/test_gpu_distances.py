"""Per-pair distances / witness points / collision Jacobians and the collision-avoidance null-space
term on the GPU against the float64 oracle and the reference's own tests
(/root/reference/test/ik/test_nullspace_components.py:75-111)."""

import numpy as np
import pytest

from conftest import make_panda, uniform_states  # noqa: F401
from oracle import ik_oracle as ik
from test_oracle_golden import _contains

pytestmark = pytest.mark.gpu


def test_distances_and_witness_points(panda_full, panda_oracle):
    from pyroboplan_b200.engine import get_engine

    model, cmodel, _ = panda_full
    eng = get_engine(model, cmodel)
    q = uniform_states(model, 300, 41)
    dist, p1, p2, nrm = eng.compute_distances(q)
    assert dist.shape == (300, 74) and p1.shape == (300, 74, 3)
    geoms = cmodel.geometryObjects
    checked = penetrating = 0
    for i in range(0, 300, 7):
        _, oMg = panda_oracle.fk(q[i].astype(np.float64))
        for k, cp in enumerate(cmodel.collisionPairs):
            d_ref, pa, pb, inter = panda_oracle.pair_distance(q[i].astype(np.float64), cp.first, cp.second)
            if inter or dist[i, k] <= 0.0:
                # overlapping: negative distance = penetration depth (coal's convention), FP32 EPA vs the oracle's
                ds, pas, pbs, ns, exact = panda_oracle.pair_signed_distance(q[i].astype(np.float64), cp.first, cp.second)
                if exact and abs(ds) > 1e-5:
                    assert abs(dist[i, k] - ds) < 3e-5, (i, k, dist[i, k], ds)
                    np.testing.assert_allclose(p2[i, k].astype(np.float64) - p1[i, k], nrm[i, k].astype(np.float64) * dist[i, k], atol=1e-5)
                    if abs(ds) > 1e-3:
                        assert nrm[i, k] @ ns > 1.0 - 1e-3
                    penetrating += 1
                else:
                    assert abs(dist[i, k]) < 3e-5 or not exact
                continue
            assert abs(dist[i, k] - d_ref) < 1e-5          # north_star: distances within 1e-5 m
            a, b = p1[i, k].astype(np.float64), p2[i, k].astype(np.float64)
            assert _contains(geoms[cp.first], oMg[cp.first], a, 5e-6) and _contains(geoms[cp.second], oMg[cp.second], b, 5e-6)
            assert abs(np.linalg.norm(b - a) - d_ref) < 2e-5
            if d_ref > 1e-3:
                np.testing.assert_allclose(nrm[i, k], (pb - pa) / d_ref, atol=2e-3 + 2e-5 / d_ref)
            checked += 1
    assert checked > 2000 and penetrating > 20
    # check_states(return_distance=True) reports the minimum over the pairs of the SURFACE distance (the mesh
    # triangles, as the reference measures); the per-pair distances here are hull distances: a lower bound,
    # equal except where the nearest feature lies in a dent of a mesh (sub-millimetre)
    _, md = eng.check_states(q, return_distance=True)
    hull_min = np.maximum(dist.min(axis=1), 0.0)
    shallow = dist.min(axis=1) > -1e-3      # (a shape deep inside another hull may be ENCLOSED by its surface: distance > 0)
    assert np.all(md >= hull_min - 2e-6) and np.all(md[shallow] <= hull_min[shallow] + 1e-3)
    assert (np.abs(md - hull_min) <= 2e-6).mean() > 0.95
    # a cutoff only replaces far pairs by lower bounds
    dc = eng.compute_distances(q, cutoff=0.05, witness=False)
    near = dist <= 0.05
    np.testing.assert_array_equal(dc[near], dist[near])
    assert (dc[~near] > 0.05).all() and (dc[~near] <= dist[~near] + 1e-6).all()


def test_collision_vector_and_jacobians(panda_full, panda_oracle):
    from pyroboplan_b200.core.utils import calculate_collision_vector_and_jacobians

    model, cmodel, _ = panda_full
    q = uniform_states(model, 6, 43).astype(np.float64)
    n = 0
    for qi in q:
        for k in range(0, 74, 5):
            vec_ref, J1_ref, J2_ref, d, exact = ik.collision_vector_and_jacobians(model, cmodel, panda_oracle, k, qi.astype(np.float32).astype(np.float64))
            if not exact or abs(d) < 1e-3:
                continue
            vec, J1, J2 = calculate_collision_vector_and_jacobians(model, cmodel, None, None, k, qi)
            assert J1.shape == (3, 9) and J2.shape == (3, 9)
            # witness points of parallel faces are not unique: compare what the avoidance term uses,
            # the distance vector and the Jacobians projected on it
            np.testing.assert_allclose(vec, vec_ref, atol=3e-3 * abs(d) + 2e-5)
            np.testing.assert_allclose(vec_ref @ (J2 - J1), vec_ref @ (J2_ref - J1_ref), atol=5e-3 * abs(d) + 1e-5)
            n += 1
    assert n > 40


def test_collision_avoidance_component(panda_full, panda_oracle):
    from pyroboplan_b200.engine import get_engine

    model, cmodel, _ = panda_full
    eng = get_engine(model, cmodel)
    q = uniform_states(model, 400, 47)
    comp = eng.collision_nullspace(q, dist_padding=0.05, gain=1.0)
    assert comp.shape == (400, 9)
    compared = nonzero = 0
    for i in range(400):
        ref, exact = ik.collision_avoidance_component(model, cmodel, panda_oracle, q[i].astype(np.float64), 0.05, 1.0)
        if not exact:
            continue    # a pair touching in a way the oracle's polytope expansion cannot start from
        np.testing.assert_allclose(comp[i], ref, atol=2e-4 + 2e-3 * np.abs(ref).max())
        compared += 1
        nonzero += int(np.abs(ref).max() > 0)
    assert compared > 300 and nonzero > 100


def test_reference_nullspace_component_tests():
    """/root/reference/test/ik/test_nullspace_components.py:75-111 on the self-collision Panda."""
    from pyroboplan_b200.ik.nullspace_components import (
        collision_avoidance_nullspace_component,
        joint_center_nullspace_component,
        joint_limit_nullspace_component,
        zero_nullspace_component,
    )

    model, cmodel, _ = make_panda(objects=False)
    q_mid = 0.5 * (model.lowerPositionLimit + model.upperPositionLimit)
    np.testing.assert_array_equal(zero_nullspace_component(model, q_mid), np.zeros(9))
    np.testing.assert_array_equal(joint_center_nullspace_component(model, q_mid), np.zeros(9))
    np.testing.assert_array_equal(joint_limit_nullspace_component(model, q_mid), np.zeros(9))
    q_out = q_mid.copy()
    q_out[0] = model.upperPositionLimit[0] + 0.2
    q_out[3] = model.lowerPositionLimit[3] - 0.1
    g = joint_limit_nullspace_component(model, q_out, gain=2.0)
    np.testing.assert_allclose(g[[0, 3]], [-0.4, 0.2])
    assert np.count_nonzero(g) == 2
    # collision free, far from any self collision: exactly zero (reference :75-92)
    comp = collision_avoidance_nullspace_component(model, None, cmodel, None, q_mid, dist_padding=0.05)
    np.testing.assert_array_equal(comp, np.zeros(9))
    # the reference's collision_state: non-zero (reference :95-111)
    q_coll = np.array([1.98103111, 0.27084068, -1.60819541, -3.07282607, -0.56607161, 2.2466373, 2.80244523, 0.01154539, 0.00453211])
    comp = collision_avoidance_nullspace_component(model, None, cmodel, None, q_coll, gain=1.0, dist_padding=0.05)
    assert comp.shape == (9,)
    assert not np.all(comp == 0.0)
    # joint-centre vector of the reference (:58-66)
    q = 0.5 * (model.lowerPositionLimit + model.upperPositionLimit)
    q[2] += 0.25
    q[6] -= 0.5
    c = joint_center_nullspace_component(model, q, gain=0.5)
    assert c[2] == pytest.approx(-0.125) and c[6] == pytest.approx(0.25)


def test_ik_with_nullspace_components(panda_full, panda_oracle):
    """One try with joint-centre + collision-avoidance null-space terms: every iteration of the
    device path equals the float64 oracle's iteration (terms evaluated by the oracle)."""
    import torch

    from pyroboplan_b200.engine import get_engine
    from pyroboplan_b200.ik.nullspace_components import collision_avoidance_nullspace_component_batch, joint_center_nullspace_component_batch

    model, cmodel, _ = panda_full
    eng = get_engine(model, cmodel)
    fid = model.getFrameId("panda_hand")
    n = 40
    goal = eng.sample_free_states(n, seed=3).cpu().numpy().astype(np.float64)
    init = eng.sample_free_states(n, seed=4).cpu().numpy().astype(np.float64)
    targets = np.array([np.concatenate([T[:3, :3].reshape(9), T[:3, 3]]) for T in (ik.frame_pose(model, fid, g)[0] for g in goal)])
    Qd = torch.as_tensor(init, device="cuda:0")
    Td = torch.as_tensor(targets, device="cuda:0")
    ns = joint_center_nullspace_component_batch(model, Qd, gain=0.1) + collision_avoidance_nullspace_component_batch(model, cmodel, Qd, 0.05, 1.0)
    Q1, status, iters, _ = eng.ik_iterate(fid, Td, Qd, max_iters=1, nullspace=ns)
    Q1 = Q1.cpu().numpy()
    ns_h = ns.cpu().numpy()
    for i in range(n):
        T = np.eye(4)
        T[:3, :3] = targets[i, :9].reshape(3, 3)
        T[:3, 3] = targets[i, 9:]
        comps = [lambda m, q, v=ns_h[i]: v]   # the oracle step with the SAME null-space vector: isolates the IK arithmetic
        q_ref, _, _, it, _ = ik.ik_try(model, fid, T, init[i], max_iters=1, nullspace_components=comps)
        step = np.abs(q_ref - init[i]).max()
        np.testing.assert_allclose(Q1[i], q_ref, atol=1e-9 * max(1.0, step / 1e-3))
        assert np.abs(Q1[i, 7:] - init[i, 7:]).max() > 0 or np.abs(ns_h[i, 7:]).max() == 0   # fingers move by the null-space term only


def test_minimum_distance_is_signed(panda_self):
    """get_minimum_distance_at_state returns coal's signed value: negative (the penetration depth) at the
    reference's own colliding state (/root/reference/test/core/test_utils.py:129-158), positive at its free states."""
    from conftest import COLLISION_STATE, FREE_STATE_1, FREE_STATE_2
    from oracle.oracle import Oracle
    from pyroboplan_b200.core.utils import get_minimum_distance_at_state

    model, cmodel, _ = panda_self
    orc = Oracle(model, cmodel)
    for q in (COLLISION_STATE, FREE_STATE_1, FREE_STATE_2):
        q32 = q.astype(np.float32).astype(np.float64)
        ref = min(orc.pair_signed_distance(q32, p.first, p.second)[0] for p in cmodel.collisionPairs)
        got = get_minimum_distance_at_state(model, cmodel, q)
        assert got == pytest.approx(ref, abs=3e-5)
    assert get_minimum_distance_at_state(model, cmodel, COLLISION_STATE) < -1e-3
    assert get_minimum_distance_at_state(model, cmodel, FREE_STATE_1) == pytest.approx(0.1024, abs=2e-4)   # the survey's independent probe
