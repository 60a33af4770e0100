"""Cartesian planner mirror: option validation and time scaling on the CPU
(/root/reference/test/planning/test_cartesian_planner.py:15-31), planning on the GPU (:34-98)."""

import numpy as np
import pytest
from scipy.spatial.transform import Rotation

from conftest import make_panda
from pyroboplan_b200.model import SE3
from pyroboplan_b200.planning.cartesian_planner import CartesianPlanner, CartesianPlannerOptions, _UnitTrapezoid


@pytest.mark.parametrize("max_vals", [(0.0, 1.0, 1.0, 1.0), (1.0, 0.0, 1.0, 1.0), (1.0, 1.0, 0.0, 1.0), (1.0, 1.0, 1.0, 0.0)])
def test_invalid_cartesian_planner_options(max_vals):
    with pytest.raises(ValueError):
        CartesianPlannerOptions(max_linear_velocity=max_vals[0], max_linear_acceleration=max_vals[1],
                                max_angular_velocity=max_vals[2], max_angular_acceleration=max_vals[3])


def test_unit_trapezoid_profiles():
    # triangular: v^2 >= a -> peak sqrt(a), duration 2 / sqrt(a) (reference trapezoidal_velocity.py:88-92)
    tri = _UnitTrapezoid(5.0, 4.0)
    assert tri.duration == pytest.approx(1.0) and tri.position(0.5) == pytest.approx(0.5)
    # trapezoid: accelerate v/a, cruise (1 - v^2/a)/v (reference :113-119)
    tz = _UnitTrapezoid(1.0, 4.0)
    assert tz.duration == pytest.approx(0.25 + 0.75 + 0.25)
    assert tz.position(0.25) == pytest.approx(0.125) and tz.position(1.0) == pytest.approx(0.875)
    ts = np.linspace(0, tz.duration, 200)
    p = np.array([tz.position(t) for t in ts])
    assert p[0] == 0.0 and p[-1] == pytest.approx(1.0) and (np.diff(p) >= 0).all()
    assert np.max(np.diff(p) / np.diff(ts)) <= 1.0 + 1e-9


def test_segment_times_and_pose_interpolation():
    a = SE3(np.eye(3), np.zeros(3))
    b = SE3(Rotation.from_euler("z", 90, degrees=True).as_matrix(), np.array([0.5, 0.0, 0.0]))
    planner = CartesianPlanner(None, "f", [a, b, a], None, CartesianPlannerOptions(use_trapezoidal_scaling=False))
    # linear scaling: max(0.5 / 1, (pi / 2) / 1) per segment (reference :141-147)
    assert planner.segments[0][1] == pytest.approx(np.pi / 2) and planner.segments[1][1] == pytest.approx(np.pi)
    t_vec = planner._time_vector(0.1)
    assert t_vec[-1] == pytest.approx(np.pi) and np.all(np.diff(t_vec) > 0)
    poses = planner._poses(np.array([0.0, np.pi / 4, np.pi / 2, np.pi]))
    np.testing.assert_allclose(poses[1].translation, [0.25, 0, 0], atol=1e-12)
    np.testing.assert_allclose(poses[1].rotation, Rotation.from_euler("z", 45, degrees=True).as_matrix(), atol=1e-12)
    np.testing.assert_allclose(poses[2].translation, [0.5, 0, 0], atol=1e-12)
    np.testing.assert_allclose(poses[3].rotation, np.eye(3), atol=1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize("use_trapezoidal_scaling", [True, False])
def test_cartesian_planner(use_trapezoidal_scaling):
    from pyroboplan_b200.core.utils import check_collisions_at_state, extract_cartesian_pose
    from pyroboplan_b200.ik.differential_ik import DifferentialIk

    model, cmodel, _ = make_panda(objects=False)
    q_start = np.array([0.0, -0.785, 0.0, -2.356, 0.0, 1.571, 0.785, 0.0, 0.0])
    init = extract_cartesian_pose(model, "panda_hand", q_start)
    rot = Rotation.from_euler("z", 60, degrees=True).as_matrix()
    tforms = [init, init * SE3(rot, np.array([0.0, 0.0, 0.2])), init * SE3(np.eye(3), np.array([0.0, 0.2, 0.2]))]
    ik = DifferentialIk(model, collision_model=cmodel)
    planner = CartesianPlanner(model, "panda_hand", tforms, ik, options=CartesianPlannerOptions(use_trapezoidal_scaling=use_trapezoidal_scaling))
    success, t_vec, q_vec = planner.generate(q_start, 0.05)
    assert success
    assert q_vec.shape[0] == model.nq and q_vec.shape[1] == len(t_vec)
    # every sample reaches its pose and is collision free
    for k in range(0, len(t_vec), 5):
        got = extract_cartesian_pose(model, "panda_hand", q_vec[:, k])
        err = planner.generated_tforms[k].actInv(got)
        assert np.linalg.norm(err.translation) < 2e-3
        assert not check_collisions_at_state(model, cmodel, q_vec[:, k])
    # batched: the same plan raced from several seeds around q_start
    rng = np.random.default_rng(0)
    seeds = q_start + rng.normal(scale=0.02, size=(8, 9))
    seeds[:, 7:] = 0.0
    ok, t2, qb = planner.generate_batch(seeds, 0.05)
    assert qb.shape == (8, 9, len(t_vec)) and ok.sum() >= 6
    lane = int(np.nonzero(ok)[0][0])
    got = extract_cartesian_pose(model, "panda_hand", qb[lane, :, -1])
    assert np.linalg.norm(planner.generated_tforms[-1].actInv(got).translation) < 2e-3


@pytest.mark.gpu
def test_cartesian_planner_failure():
    from pyroboplan_b200.core.utils import extract_cartesian_pose
    from pyroboplan_b200.ik.differential_ik import DifferentialIk, DifferentialIkOptions

    model, cmodel, _ = make_panda(objects=False)
    q_start = np.array([0.0, -0.785, 0.0, -2.356, 0.0, 1.571, 0.785, 0.0, 0.0])
    init = extract_cartesian_pose(model, "panda_hand", q_start)
    tforms = [init, init * SE3(np.eye(3), np.array([10.0, 10.0, 10.0]))]
    ik = DifferentialIk(model, collision_model=cmodel, options=DifferentialIkOptions(max_retries=2))
    success, _, _ = CartesianPlanner(model, "panda_hand", tforms, ik).generate(q_start, 0.05)
    assert not success
