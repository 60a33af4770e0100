"""GPU parity tests (run on the B200 box: `pytest -m gpu`).  Every call goes through the C ABI
(libpbp_engine.so via ctypes) and is compared with the float64 oracle on the same float32 inputs.

Bar (BASELINE.json north_star): identical collision booleans for every sample whose reference
minimum distance exceeds 1e-6 m; link poses and distances within 1e-5 m."""

import os

import numpy as np
import pytest

from conftest import COLLISION_STATE, FREE_STATE_1, FREE_STATE_2, ROOT, make_panda, make_two_dof, uniform_states

pytestmark = pytest.mark.gpu

BAND = 1e-6      # verdicts may differ only where |reference min distance - padding| <= BAND
DIST_TOL = 1e-5  # metres
GOLDEN = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def torch_cuda():
    import torch

    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch


@pytest.fixture(scope="module")
def panda_engine(torch_cuda, panda_full):
    from pyroboplan_b200.engine import CollisionEngine

    return CollisionEngine(panda_full[0], panda_full[1], device=0)


@pytest.fixture(scope="module")
def two_dof_engine(torch_cuda, two_dof_full):
    from pyroboplan_b200.engine import CollisionEngine

    return CollisionEngine(two_dof_full[0], two_dof_full[1], device=0)


def assert_verdicts(hit_gpu, hit_ref, min_dist_ref, padding=0.0):
    hit_gpu = np.asarray(hit_gpu).astype(bool)
    bad = np.nonzero(hit_gpu != hit_ref.astype(bool))[0]
    outside = [i for i in bad if abs(min_dist_ref[i] - padding) > BAND]
    assert not outside, f"{len(outside)} verdict mismatches outside the {BAND} m band, e.g. idx {outside[:5]} d={min_dist_ref[outside[:5]]}"
    return len(bad)


def assert_mesh_distances(dist, ref):
    """Min distances of a model with meshes: the oracle measures to the mesh TRIANGLES (the reference's BVHModel
    surfaces), and so does the engine wherever hull and inner core disagree -- equal within DIST_TOL everywhere."""
    err = np.abs(np.asarray(ref, dtype=np.float64) - np.asarray(dist, dtype=np.float64))
    assert err.max() < DIST_TOL, (err.max(), int(np.argmax(err)))


# ------------------------------------------------------------------------------------------ golden
def test_reference_known_answer_states_on_gpu(torch_cuda):
    """The reference's own known-answer test (test/core/test_utils.py:125-212) through the drop-in API."""
    from pyroboplan_b200.core.utils import check_collisions_along_path, check_collisions_at_state

    model, cmodel, _ = make_panda(objects=False)
    assert check_collisions_at_state(model, cmodel, COLLISION_STATE)
    assert not check_collisions_at_state(model, cmodel, FREE_STATE_1)
    assert isinstance(check_collisions_at_state(model, cmodel, FREE_STATE_1), np.bool_)
    assert check_collisions_along_path(model, cmodel, [FREE_STATE_1, COLLISION_STATE, FREE_STATE_2])
    assert not check_collisions_along_path(model, cmodel, [FREE_STATE_1, FREE_STATE_2, FREE_STATE_1])
    g = np.load(os.path.join(GOLDEN, "panda_self_reference_states.npz"))
    from pyroboplan_b200.core.utils import check_collisions_at_states

    hit, dist, pair = check_collisions_at_states(model, cmodel, g["q"], return_distance=True, return_pair=True)
    assert list(hit) == [True, False, False]
    np.testing.assert_allclose(dist, g["min_dist"], atol=DIST_TOL)
    assert list(pair) == list(g["first_pair"])


@pytest.mark.parametrize("name,which", [("panda_states", "panda"), ("panda_states_padding", "panda"), ("two_dof_states", "two_dof")])
def test_golden_states(name, which, panda_engine, two_dof_engine):
    g = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    eng = panda_engine if which == "panda" else two_dof_engine
    padding = float(g["padding"])
    hit = eng.check_states(g["q"], padding)
    md_v = g["min_dist"]
    assert_verdicts(hit, g["hit"], md_v, padding)
    hit2, dist, pair = eng.check_states(g["q"], padding, return_distance=True, return_pair=True)
    assert_verdicts(hit2, g["hit"], md_v, padding)
    free = g["min_dist"] > BAND
    if which == "panda":
        assert_mesh_distances(dist[free], g["min_dist"][free])
    else:
        assert np.abs(dist[free] - g["min_dist"][free]).max() < DIST_TOL
    assert np.all(dist[~free] <= DIST_TOL)
    clear = np.abs(g["min_dist"] - padding) > BAND
    assert (pair[clear] == g["first_pair"][clear]).all()


def test_golden_edges(panda_engine):
    g = np.load(os.path.join(GOLDEN, "panda_edges.npz"))
    hit = panda_engine.check_edges(g["q1"], g["q2"], float(g["step"]))
    assert (hit == g["hit"].astype(bool)).all()
    hit2, first_bad = panda_engine.check_edges(g["q1"], g["q2"], float(g["step"]), return_first_bad=True)
    assert (hit2 == g["hit"].astype(bool)).all()
    has_idx = g["first_bad"] >= 0  # -2 marks "q2 itself collides" (the reference tests q2 first)
    assert (first_bad[has_idx] == g["first_bad"][has_idx]).all()
    assert (first_bad[g["hit"] == 0] == -1).all()
    counts, offsets, states = panda_engine.discretize(g["q1"], g["q2"], float(g["step"]))
    assert (counts == g["counts"]).all()


# ------------------------------------------------------------------------------------------ live oracle
def test_fk_parity(panda_engine, panda_full, panda_oracle):
    q = uniform_states(panda_full[0], 500, 31)
    oMi, oMg, oMf = panda_engine.fk(q)
    worst = 0.0
    for i in range(len(q)):
        a, b = panda_oracle.fk(q[i].astype(np.float64))
        worst = max(worst, np.abs(oMi[i] - a).max(), np.abs(oMg[i] - b).max())
    assert worst < DIST_TOL
    # frames: the hand frame equals the hand geometry placement
    model = panda_full[0]
    fid = model.getFrameId("panda_hand")
    np.testing.assert_allclose(oMf[:, fid], oMg[:, 8], atol=1e-6)


def test_states_parity_200k(panda_engine, panda_full, panda_oracle, torch_cuda):
    q = uniform_states(panda_full[0], 200_000, 0)
    hit_ref, md_ref, fp_ref, _ = panda_oracle.check_states(q.astype(np.float64), want_all=True, use_cull=False)
    hit_host = panda_engine.check_states(q)
    assert_verdicts(hit_host, hit_ref, md_ref)
    qd = torch_cuda.as_tensor(q, device="cuda:0")
    hit_dev, dist_dev, pair_dev = panda_engine.check_states(qd, return_distance=True, return_pair=True)
    assert (hit_dev.cpu().numpy() == hit_host).all()  # device-pointer and host-pointer entry points agree
    free = md_ref > BAND
    assert_mesh_distances(dist_dev.cpu().numpy()[free], md_ref[free])
    clear = np.abs(md_ref) > BAND
    assert (pair_dev.cpu().numpy()[clear] == fp_ref[clear]).all()
    assert 0.38 < hit_ref.mean() < 0.42  # SURVEY.md 8d probe: P(collision) ~ 0.40


def test_config1_two_dof_10k(two_dof_engine, two_dof_full, two_dof_oracle):
    """BASELINE.json config 1: 2-DOF arm, 10 k random states, CPU reference vs GPU."""
    q = np.random.default_rng(0).uniform(-3.14, 3.14, size=(10_000, 2)).astype(np.float32)
    hit_ref, md_ref, _, _ = two_dof_oracle.check_states(q.astype(np.float64), want_all=True, use_cull=False)
    hit, dist = two_dof_engine.check_states(q, return_distance=True)
    assert_verdicts(hit, hit_ref, md_ref)
    assert_verdicts(two_dof_engine.check_states(q), hit_ref, md_ref)
    free = md_ref > BAND
    assert np.abs(dist[free] - md_ref[free]).max() < DIST_TOL
    assert 0.15 < hit_ref.mean() < 0.21


@pytest.mark.parametrize("padding", [0.0, 0.01, 0.05])
def test_padding_parity(panda_engine, panda_full, panda_oracle, padding):
    q = uniform_states(panda_full[0], 30_000, 41)
    hit_ref, _, _, _ = panda_oracle.check_states(q.astype(np.float64), padding=padding, use_cull=True)
    _, md_ref, _, _ = panda_oracle.check_states(q.astype(np.float64), want_all=True, use_cull=False)
    assert_verdicts(panda_engine.check_states(q, padding), hit_ref, md_ref, padding)


def test_edges_parity_and_properties(panda_engine, panda_full, panda_oracle, torch_cuda):
    model = panda_full[0]
    a = uniform_states(model, 6000, 51)
    b = uniform_states(model, 6000, 52)
    b[:4000] = a[:4000] + np.float32(0.1) * (b[:4000] - a[:4000])
    hit_ref, fb_ref, _, _ = panda_oracle.check_edges(a.astype(np.float64), b.astype(np.float64), 0.05)
    hit = panda_engine.check_edges(a, b, 0.05)
    assert (hit == hit_ref.astype(bool)).all()
    # property: an edge verdict equals the OR over its explicitly discretised states
    counts, offsets, states = panda_engine.discretize(a, b, 0.05)
    per_state = panda_engine.check_states(states)
    any_hit = np.logical_or.reduceat(per_state, offsets[:-1])
    assert (any_hit == hit).all()
    # property: check_paths on the explicit discretisation gives the same verdicts
    assert (panda_engine.check_paths(states, offsets) == hit).all()
    # property: reversing an edge does not change its verdict (same sample set up to rounding)
    rev = panda_engine.check_edges(b, a, 0.05)
    assert (rev != hit).sum() <= 2
    # zero-length edge -> one sample (reference test_planning_utils.py:32-34)
    c1, _, s1 = panda_engine.discretize(a[:3], a[:3], 0.05)
    assert list(c1) == [1, 1, 1] and np.array_equal(s1, a[:3])
    assert (panda_engine.check_edges(a[:50], a[:50], 0.05) == panda_engine.check_states(a[:50])).all()


def test_discretize_matches_reference_formula(panda_engine, panda_full):
    model = panda_full[0]
    a = uniform_states(model, 200, 61)
    b = uniform_states(model, 200, 62)
    for step in (0.05, 0.3):
        counts, offsets, states = panda_engine.discretize(a, b, step)
        for e in range(0, 200, 13):
            qa, qb = a[e].astype(np.float64), b[e].astype(np.float64)
            n = int(np.ceil(np.linalg.norm(qb - qa) / step)) + 1  # reference planning/utils.py:137
            assert counts[e] == n
            ref = np.array([qa + s * (qb - qa) for s in np.linspace(0.0, 1.0, n)])
            np.testing.assert_allclose(states[offsets[e] : offsets[e + 1]], ref, atol=2e-6)
    # golden vector of the reference's test: [0,0] -> [2,0] at 1.0 gives exactly 3 samples
    m2, c2, _ = make_two_dof()
    from pyroboplan_b200.engine import CollisionEngine

    e2 = CollisionEngine(m2, c2, device=0)
    c, o, s = e2.discretize(np.array([[0.0, 0.0]]), np.array([[2.0, 0.0]]), 1.0)
    assert list(c) == [3]
    np.testing.assert_array_equal(s, [[0.0, 0.0], [1.0, 0.0], [2.0, 0.0]])


def test_chunking_and_empty_inputs(panda_full, panda_engine, torch_cuda):
    from pyroboplan_b200.engine import CollisionEngine

    model, cmodel, _ = panda_full
    small = CollisionEngine(model, cmodel, device=0, max_chunk_states=1024)  # forces many rounds
    q = uniform_states(model, 10_000, 71)
    h_small, d_small = small.check_states(q, return_distance=True)
    h_big, d_big = panda_engine.check_states(q, return_distance=True)
    assert (h_small == h_big).all() and np.array_equal(d_small, d_big)
    a, b = q[:3000], q[3000:6000]
    assert (small.check_edges(a, b, 0.05) == panda_engine.check_edges(a, b, 0.05)).all()
    assert panda_engine.check_states(np.zeros((0, 9), dtype=np.float32)).shape == (0,)
    assert panda_engine.check_edges(np.zeros((0, 9)), np.zeros((0, 9)), 0.05).shape == (0,)
    ragged = np.array([0, 0, 5, 5, 12], dtype=np.int64)  # empty paths in the middle
    ph = panda_engine.check_paths(q[:12], ragged)
    per = panda_engine.check_states(q[:12])
    assert list(ph) == [False, bool(per[:5].any()), False, bool(per[5:12].any())]
    with pytest.raises(RuntimeError, match="padding"):
        panda_engine.check_states(q[:4], -0.1)
    small.close()


def test_dropin_single_calls_match_batched(panda_full, panda_engine, panda_oracle):
    from pyroboplan_b200.core.utils import (
        check_collisions_along_path,
        check_collisions_at_state,
        get_minimum_distance_at_state,
        get_random_collision_free_state,
    )
    from pyroboplan_b200.planning.utils import discretize_joint_space_path, has_collision_free_path, has_collision_free_paths

    model, cmodel, _ = panda_full
    q = uniform_states(model, 40, 81).astype(np.float64)
    ref, md, _, _ = panda_oracle.check_states(q, want_all=True, use_cull=False)
    for i in range(len(q)):
        assert bool(check_collisions_at_state(model, cmodel, q[i])) == bool(ref[i])
        if md[i] > BAND:
            assert abs(get_minimum_distance_at_state(model, cmodel, q[i]) - md[i]) <= DIST_TOL
    free_idx = np.nonzero(ref == 0)[0]
    for i, j in zip(free_idx[:-1:2], free_idx[1::2]):
        path = discretize_joint_space_path([q[i], q[j]], 0.05)
        expected = panda_oracle.check_path(path)
        assert check_collisions_along_path(model, cmodel, path) == expected
        assert has_collision_free_path(q[i], q[j], 0.05, model, cmodel) == (not expected)
    ok = has_collision_free_paths(q[:20], q[20:], 0.05, model, cmodel)
    assert ok.shape == (20,)
    np.random.seed(1234)
    s = get_random_collision_free_state(model, cmodel)
    assert s is not None and not panda_oracle.check_state(s)
    assert np.all(s >= model.lowerPositionLimit) and np.all(s <= model.upperPositionLimit)


def test_sampler(panda_full, panda_engine, panda_oracle):
    model = panda_full[0]
    s = panda_engine.sample_free_states(20_000, seed=5, as_numpy=True)
    assert s.shape == (20_000, 9)
    assert np.all(s >= model.lowerPositionLimit.astype(np.float32)) and np.all(s <= model.upperPositionLimit.astype(np.float32))
    hit, md, _, _ = panda_oracle.check_states(s.astype(np.float64), want_all=True, use_cull=False)
    assert not (hit.astype(bool) & (md < -BAND)).any() and hit.sum() <= 1  # at most a boundary sample
    s2 = panda_engine.sample_free_states(20_000, seed=5, as_numpy=True)
    assert np.array_equal(s, s2)  # counter-based stream: same seed, same states
    acceptance = 20_000 / panda_engine.last_sample_draws
    assert 0.3 < acceptance < 0.8  # ~60 % of uniform Panda states are free


def test_full_size_config2_properties(panda_engine, panda_full, panda_oracle, torch_cuda):
    """BASELINE.json config 2 at full size (1 M states): verdict parity with the oracle on the whole
    batch and size-independent properties (idempotence, permutation invariance, concatenation)."""
    n = 1_000_000
    q = uniform_states(panda_full[0], n, 0)
    qd = torch_cuda.as_tensor(q, device="cuda:0")
    hit = panda_engine.check_states(qd)
    assert (panda_engine.check_states(qd) == hit).all()
    perm = torch_cuda.randperm(n, device="cuda:0")
    assert (panda_engine.check_states(qd[perm]) == hit[perm]).all()
    halves = torch_cuda.cat([panda_engine.check_states(qd[: n // 3]), panda_engine.check_states(qd[n // 3 :])])
    assert (halves == hit).all()
    hit_ref, md_ref, _, _ = panda_oracle.check_states(q.astype(np.float64), want_all=True, use_cull=False)
    boundary = assert_verdicts(hit.cpu().numpy(), hit_ref, md_ref)
    assert boundary <= 5


def test_full_size_config3_edges(panda_engine, panda_full, panda_oracle, torch_cuda):
    """BASELINE.json config 3 at full size: 100 k random edges at 0.05 rad discretisation
    (~10.7 M samples); every edge verdict is compared with the oracle."""
    model = panda_full[0]
    rng = np.random.default_rng(1)
    q1 = rng.uniform(model.lowerPositionLimit, model.upperPositionLimit, size=(100_000, 9)).astype(np.float32)
    q2 = rng.uniform(model.lowerPositionLimit, model.upperPositionLimit, size=(100_000, 9)).astype(np.float32)
    hit = panda_engine.check_edges(torch_cuda.as_tensor(q1, device="cuda:0"), torch_cuda.as_tensor(q2, device="cuda:0"), 0.05)
    hit_ref, _, evals, _ = panda_oracle.check_edges(q1.astype(np.float64), q2.astype(np.float64), 0.05)
    bad = np.nonzero(hit.cpu().numpy() != hit_ref.astype(bool))[0]
    # an edge verdict can only differ through a sample inside the 1e-6 m band; none is expected here
    assert len(bad) <= 2, bad[:10]
    counts, offsets, _ = panda_engine.discretize(q1[:2000], q2[:2000], 0.05)
    assert 100 < counts.mean() < 115   # SURVEY.md 8d probe: ~107.5 samples per random edge


@pytest.mark.parametrize("n_edges", [1, 7, 300, 600, 5000, 20000])
def test_edge_batch_sizes_pick_different_level_schedules(panda_engine, panda_oracle, panda_full, n_edges):
    """One level (<= 512 edges; host path without any device -> host round trip), two levels
    (<= 16384) and the full 128..1 hierarchy give the same verdicts as the oracle's in-order walk."""
    import torch

    model = panda_full[0]
    a = uniform_states(model, n_edges, 300 + n_edges)
    b = uniform_states(model, n_edges, 900 + n_edges)
    b[: n_edges // 2] = a[: n_edges // 2] + np.float32(0.08) * (b[: n_edges // 2] - a[: n_edges // 2])
    ref, _, _, _ = panda_oracle.check_edges(a.astype(np.float64), b.astype(np.float64), 0.05)
    got_host = panda_engine.check_edges(a, b, 0.05)
    got_dev = panda_engine.check_edges(torch.as_tensor(a, device="cuda:0"), torch.as_tensor(b, device="cuda:0"), 0.05).cpu().numpy()
    assert (got_host == ref.astype(bool)).all()
    assert (got_dev == ref.astype(bool)).all()
    # zero-length edges (one sample each)
    z = panda_engine.check_edges(a[:3], a[:3], 0.05)
    assert (z == panda_oracle.check_states(a[:3].astype(np.float64))[0].astype(bool)).all()
