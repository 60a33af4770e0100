"""Pins the numpy IK oracle (oracle/ik_oracle.py): SE(3) log vs scipy's matrix logarithm, the LOCAL
frame Jacobian vs finite differences, and the reference's own IK tests
(/root/reference/test/ik/test_differential_ik.py:14-81)."""

import numpy as np
from scipy.linalg import logm

from conftest import make_panda
from oracle import ik_oracle as ik


def test_log6_matches_matrix_logarithm():
    model, _, _ = make_panda(False, False)
    fid = model.getFrameId("panda_hand")
    rng = np.random.default_rng(0)
    for _ in range(20):
        q = rng.uniform(model.lowerPositionLimit, model.upperPositionLimit)
        T, _ = ik.frame_pose(model, fid, q)
        L = logm(T).real
        ref = np.array([L[0, 3], L[1, 3], L[2, 3], L[2, 1], L[0, 2], L[1, 0]])
        np.testing.assert_allclose(ik.log6(T), ref, atol=1e-10)


def test_local_jacobian_matches_finite_differences():
    model, _, _ = make_panda(False, False)
    fid = model.getFrameId("panda_hand")
    rng = np.random.default_rng(1)
    for _ in range(10):
        q = rng.uniform(model.lowerPositionLimit, model.upperPositionLimit)
        T, _ = ik.frame_pose(model, fid, q)
        J = ik.frame_jacobian_local(model, fid, q)
        h = 1e-6
        for k in range(model.nq):
            dq = np.zeros(model.nq)
            dq[k] = h
            Tp, _ = ik.frame_pose(model, fid, q + dq)
            np.testing.assert_allclose(J[:, k], ik.log6(np.linalg.inv(T) @ Tp) / h, atol=1e-5)
        assert np.all(J[:, 7:] == 0)  # the fingers do not move the hand frame


def test_reference_trivial_and_offset_ik():
    model, _, _ = make_panda(False, False)
    fid = model.getFrameId("panda_hand")
    q = np.array([0.0, 1.57, 0.0, 0.0, 1.57, 1.57, 0.0, 0.0, 0.0])
    T, _ = ik.frame_pose(model, fid, q)
    qs, conv, within, iters, _ = ik.ik_try(model, fid, T, q)
    assert conv and within and iters == 0
    np.testing.assert_almost_equal(q, qs, decimal=6)
    T2 = T.copy()
    T2[2, 3] += 0.01
    qs, conv, within, iters, _ = ik.ik_try(model, fid, T2, q)
    assert conv and within
    err = np.linalg.inv(T2) @ ik.frame_pose(model, fid, qs)[0]
    np.testing.assert_almost_equal(np.identity(3), err[:3, :3], decimal=3)
    np.testing.assert_almost_equal(np.zeros(3), err[:3, 3], decimal=3)


def test_reference_collision_avoidance_known_answers():
    """/root/reference/test/ik/test_nullspace_components.py:69-111: the avoidance term is exactly zero
    at mid-range (every self pair farther than 5 cm) and non-zero at the given colliding state."""
    from oracle.oracle import Oracle

    model, cmodel, _ = make_panda(objects=False)
    orc = Oracle(model, cmodel)
    q_mid = 0.5 * (model.lowerPositionLimit + model.upperPositionLimit)
    comp, exact = ik.collision_avoidance_component(model, cmodel, orc, q_mid)
    assert comp.shape == (9,) and np.all(comp == 0.0) and exact
    q = np.array([1.98103111, 0.27084068, -1.60819541, -3.07282607, -0.56607161, 2.2466373, 2.80244523, 0.01154539, 0.00453211])
    comp, _ = ik.collision_avoidance_component(model, cmodel, orc, q, dist_padding=0.05, gain=1.0)
    assert not np.all(comp == 0.0)


def test_point_jacobian_matches_finite_differences():
    from oracle.oracle import Oracle

    model, cmodel, _ = make_panda()
    rng = np.random.default_rng(3)
    for _ in range(5):
        q = rng.uniform(model.lowerPositionLimit, model.upperPositionLimit)
        T = ik.all_joint_poses(model, q)
        for j in (3, 7, 9):
            local = rng.normal(size=3) * 0.1
            p = T[j][:3, :3] @ local + T[j][:3, 3]
            J = ik.point_jacobian(model, T, j, p)
            for k in range(model.nq):
                dq = np.zeros(model.nq)
                dq[k] = 1e-6
                Tp = ik.all_joint_poses(model, q + dq)
                pp = Tp[j][:3, :3] @ local + Tp[j][:3, 3]
                np.testing.assert_allclose(J[:, k], (pp - p) / 1e-6, atol=2e-5)
