"""Batched differential IK (BASELINE.json config 5 / SURVEY.md 8a row a9) on the GPU vs the numpy oracle
and the reference's own IK tests (/root/reference/test/ik/test_differential_ik.py)."""

import numpy as np
import pytest

from conftest import make_panda, uniform_states
from oracle import ik_oracle as ik

pytestmark = pytest.mark.gpu


def se3_rows(T):
    return np.concatenate([T[:3, :3].reshape(9), T[:3, 3]])


@pytest.fixture(scope="module")
def panda_self():
    return make_panda(objects=False)


def test_ik_single_iteration_matches_oracle(panda_self):
    """One DLS iteration from random states: FK, log6, Jacobian, damped solve and step, to rounding."""
    import torch

    from pyroboplan_b200.engine import get_engine

    model, cmodel, _ = panda_self
    eng = get_engine(model, cmodel)
    fid = model.getFrameId("panda_hand")
    n = 300
    goal_q = uniform_states(model, n, 81).astype(np.float64)
    init_q = uniform_states(model, n, 82).astype(np.float64)
    targets = np.array([se3_rows(ik.frame_pose(model, fid, g)[0]) for g in goal_q])
    for kwargs in ({}, {"active_mask": 0x7F, "joint_weights": np.array([1.0, 2.0, 0.5, 1.0, 4.0, 1.0, 0.25, 1.0, 1.0])}):
        Q, status, iters, e0 = eng.ik_iterate(fid, torch.as_tensor(targets, device="cuda:0"), torch.as_tensor(init_q, device="cuda:0"),
                                              max_iters=1, **kwargs)
        Q, e0 = Q.cpu().numpy(), e0.cpu().numpy()
        assert (iters.cpu().numpy() == 1).all() and (status.cpu().numpy() == 0).all()
        for i in range(n):
            T = np.eye(4)
            T[:3, :3] = targets[i, :9].reshape(3, 3)
            T[:3, 3] = targets[i, 9:]
            okw = {}
            if kwargs:
                okw = {"active": list(range(7)), "joint_weights": 1.0 / kwargs["joint_weights"][:7]}  # W = inv(diag(weights))
            q_ref, conv, _, it_ref, e0_ref = ik.ik_try(model, fid, T, init_q[i], max_iters=1, **okw)
            assert not conv and it_ref == 1
            step = np.abs(q_ref - init_q[i]).max()
            np.testing.assert_allclose(Q[i], q_ref, atol=1e-9 * max(1.0, step / 1e-3))
            assert abs(e0[i] - e0_ref) < 1e-12


def test_ik_try_matches_oracle(panda_self):
    """One full try, target by target: same convergence verdict; where the two FP64 descents stay
    together (the rule), same iteration count and same final q to 1e-6."""
    import torch

    from pyroboplan_b200.engine import get_engine

    model, cmodel, _ = panda_self
    eng = get_engine(model, cmodel)
    fid = model.getFrameId("panda_hand")
    n = 160
    goal_q = uniform_states(model, n, 91).astype(np.float64)
    init_q = uniform_states(model, n, 92).astype(np.float64)
    targets = np.array([se3_rows(ik.frame_pose(model, fid, g)[0]) for g in goal_q])
    Q, status, iters, e0 = eng.ik_iterate(fid, torch.as_tensor(targets, device="cuda:0"), torch.as_tensor(init_q, device="cuda:0"))
    Q, status, iters = Q.cpu().numpy(), status.cpu().numpy(), iters.cpu().numpy()
    same_verdict = both = identical = 0
    for i in range(n):
        T = np.eye(4)
        T[:3, :3] = targets[i, :9].reshape(3, 3)
        T[:3, 3] = targets[i, 9:]
        q_ref, conv, within, it_ref, _ = ik.ik_try(model, fid, T, init_q[i])
        same_verdict += int(bool(status[i] & 1) == conv)
        if status[i] & 1:
            # every device solution must itself reach the target, whatever the oracle's descent did
            e = ik.log6(np.linalg.inv(T) @ ik.frame_pose(model, fid, Q[i])[0])
            assert np.linalg.norm(e[:3]) < 1e-3 and np.linalg.norm(e[3:]) < 1e-3
            inside = bool(np.all(Q[i] >= model.lowerPositionLimit) and np.all(Q[i] <= model.upperPositionLimit))
            assert bool(status[i] & 2) == inside
        if conv and (status[i] & 1):
            both += 1
            if int(iters[i]) == it_ref and np.abs(Q[i] - q_ref).max() < 1e-6:
                identical += 1
                assert bool(status[i] & 2) == within
    # With damping 1e-3 the iteration is chaotic on the few descents that wander through
    # singularities (a 1e-16 rounding difference -- Cholesky here, LU in numpy -- is amplified by up
    # to 1e6 per crossing), so bit-for-bit agreement of ALL trajectories is not a property even of
    # two FP64 implementations; the per-iteration test above pins the arithmetic instead.
    assert same_verdict >= n - 3
    assert identical >= 0.9 * both


def test_reference_ik_tests(panda_self):
    from pyroboplan_b200.core.utils import extract_cartesian_pose
    from pyroboplan_b200.ik.differential_ik import DifferentialIk, DifferentialIkOptions
    from pyroboplan_b200.model import SE3

    model, cmodel, _ = panda_self
    # trivial (reference :14-39)
    q = np.array([0.0, 1.57, 0.0, 0.0, 1.57, 1.57, 0.0, 0.0, 0.0])
    target = extract_cartesian_pose(model, "panda_hand", q)
    solver = DifferentialIk(model, options=DifferentialIkOptions())
    q_sol = solver.solve("panda_hand", target, init_state=q)
    np.testing.assert_almost_equal(q, q_sol, decimal=5)
    # 1 cm offset (reference :42-81)
    target2 = SE3(target.rotation.copy(), target.translation + np.array([0.0, 0.0, 0.01]))
    q_sol = solver.solve("panda_hand", target2, init_state=q)
    assert q_sol is not None
    got = extract_cartesian_pose(model, "panda_hand", q_sol)
    err = target2.inverse() * got
    np.testing.assert_almost_equal(np.identity(3), err.rotation, decimal=3)
    np.testing.assert_almost_equal(np.zeros(3), err.translation, decimal=3)
    # unreachable target (reference :84-103)
    np.random.seed(1234)
    far = SE3(np.identity(3), np.array([10.0, 10.0, 10.0]))
    assert DifferentialIk(model, options=DifferentialIkOptions(max_retries=2)).solve("panda_hand", far) is None
    # reachable only in self-collision (reference :106-133): the known-answer vector of the collision gate
    R = np.array([[0.206636, -0.430153, 0.878789], [-0.978337, -0.10235, 0.179946], [0.0125395, -0.896935, -0.441984]])
    T = np.array([0.155525, 0.0529695, 0.0259166])
    np.random.seed(1234)
    solver = DifferentialIk(model, collision_model=cmodel, options=DifferentialIkOptions())
    assert solver.solve("panda_hand", SE3(R, T)) is None


def _first_try_ok(model, cmodel, eng, fid, th, init, opts):
    """Which targets solve_batch settles in its first try (its pending set decides the layout of the restart draws)."""
    import torch

    from pyroboplan_b200.ik.differential_ik import DifferentialIk

    zero = type(opts)(max_retries=0, ignore_joint_indices=opts.ignore_joint_indices)
    _, ok, _ = DifferentialIk(model, collision_model=cmodel, options=zero).solve_batch("panda_hand", torch.as_tensor(th), init_states=init, seed=5)
    return ok.cpu().numpy()


def test_config5_batched_ik(panda_full, panda_oracle):
    """N targets = FK of free random states; every solved query must reach its target within the
    tolerances, inside the joint limits and collision free (checked with the float64 oracle)."""
    import torch

    from pyroboplan_b200.engine import get_engine
    from pyroboplan_b200.ik.differential_ik import DifferentialIk, DifferentialIkOptions

    model, cmodel, _ = panda_full
    eng = get_engine(model, cmodel)
    fid = model.getFrameId("panda_hand")
    n = 4096
    goals = eng.sample_free_states(n, seed=2)
    _, _, oMf = eng.fk(goals, joints=False, geoms=False, frames=True)
    targets = oMf[:, fid].contiguous()
    opts = DifferentialIkOptions(max_retries=3, ignore_joint_indices=[7, 8])  # fingers ignored (reference examples/differential_ik.py:35-38)
    Q, solved, tries = DifferentialIk(model, collision_model=cmodel, options=opts).solve_batch("panda_hand", targets, seed=5)
    assert Q.dtype == torch.float64
    Qh, sh = Q.cpu().numpy(), solved.cpu().numpy()
    th = targets.cpu().numpy().astype(np.float64)
    # success rate against the numpy oracle running the reference's solve loop (ik/differential_ik.py:134-320) on a
    # sample of the same targets from the SAME initial and restart states (the device sampler's draws, replayed)
    ns = 768
    R = opts.max_retries
    init = eng.sample_free_states(n, seed=5 * 7919 + 0).cpu().numpy().astype(np.float64)       # solve_batch's own draws
    pend = np.nonzero(~_first_try_ok(model, cmodel, eng, fid, th, init, opts))[0]
    restarts = eng.sample_free_states(len(pend) * R, seed=5 * 7919 + 1).cpu().numpy().astype(np.float64).reshape(len(pend), R, -1)
    where = {int(t): k for k, t in enumerate(pend)}
    solved_ref = np.zeros(ns, dtype=bool)
    for i in range(ns):
        T = np.eye(4)
        T[:3, :3], T[:3, 3] = th[i, :9].reshape(3, 3), th[i, 9:]
        starts = [init[i]] + ([restarts[where[i], r] for r in range(R)] if i in where else [])
        e0 = None
        for q0 in starts:
            q, conv, within, _, e0 = ik.ik_try(model, fid, T, q0, e0, active=list(range(7)))
            if conv and within and not panda_oracle.check_state(q.astype(np.float32).astype(np.float64)):
                solved_ref[i] = True
                break
    rate_gpu, rate_ref = sh[:ns].mean(), solved_ref.mean()
    agree = (sh[:ns] == solved_ref).mean()
    only_gpu, only_ref = int((sh[:ns] & ~solved_ref).sum()), int((~sh[:ns] & solved_ref).sum())
    print(f"[config 5] success rate on {ns} targets: engine {rate_gpu:.4f}, numpy oracle {rate_ref:.4f}; per-target agreement {agree:.4f} "
          f"(solved by the engine only {only_gpu}, by the oracle only {only_ref}); all {n} targets: engine {sh.mean():.4f}")
    # descents through singularities are chaotic (damping 1e-3 amplifies rounding by 1e6): whole tries of two FP64
    # implementations differ on a few per cent of the targets, in both directions
    assert abs(rate_gpu - rate_ref) < 0.04 and agree > 0.9 and min(only_gpu, only_ref) > 0
    idx = np.nonzero(sh)[0][:300]
    for i in idx:
        T = np.eye(4)
        T[:3, :3] = th[i, :9].reshape(3, 3)
        T[:3, 3] = th[i, 9:]
        e = ik.log6(np.linalg.inv(T) @ ik.frame_pose(model, fid, Qh[i])[0])
        assert np.linalg.norm(e[:3]) < 1e-3 + 1e-6 and np.linalg.norm(e[3:]) < 1e-3 + 1e-6   # targets are float32 FK poses
        assert np.all(Qh[i] >= model.lowerPositionLimit) and np.all(Qh[i] <= model.upperPositionLimit)
    hit, md, _, _ = panda_oracle.check_states(Qh[sh], want_all=True, use_cull=False)
    assert not (hit.astype(bool) & (md < -1e-6)).any() and hit.sum() <= 2
