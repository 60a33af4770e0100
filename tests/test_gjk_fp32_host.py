"""The engine's FP32 GJK (pyroboplan_b200/csrc/pbp_gjk.cuh), compiled for the host with g++, vs the
float64 oracle on many random (state, pair) items.  No GPU needed: fmaf / IEEE sqrt / division give
the same float32 results on host and device (the engine is compiled with -fmad=false)."""

import ctypes

import numpy as np
import pytest

import emu
from conftest import make_panda, make_two_dof, uniform_states
from oracle.oracle import Oracle, _ptr
from pyroboplan_b200.engine import flatten_model

pytestmark = pytest.mark.skipif(emu.cuda_include() is None, reason="CUDA headers not available")


def _pair_sweep(model, cmodel, q, want_dist):
    arrays, _ = flatten_model(model, cmodel)
    orc = Oracle(model, cmodel)
    oMg = np.array([orc.fk(x)[1] for x in q])
    stats = dict(items=0, mismatch_far=0, mismatch_near=0, worst_dist_err=0.0, max_iters=0)
    for p in cmodel.collisionPairs:
        a, b = p.first, p.second
        T = emu.relative_placements(oMg, a, b)
        ca = np.asarray(cmodel.geometryObjects[a].geometry.bounding_sphere()[0])
        cb = np.asarray(cmodel.geometryObjects[b].geometry.bounding_sphere()[0])
        v0 = ca - (np.einsum("nij,j->ni", T[:, :9].reshape(-1, 3, 3), cb) + T[:, 9:])
        radii = sum(float(arrays["geom_params"][g][0]) for g in (a, b) if arrays["geom_shape"][g] in (0, 3))
        hit, dist, iters = emu.gjk_batch(arrays, a, b, T, v0, radii, want_dist)
        d_ref = np.array([orc.lib.oracle_pair_distance(ctypes.byref(orc._cm), _ptr(np.ascontiguousarray(oMg[i])), a, b, None, None)
                          for i in range(len(q))])
        bad = hit.astype(bool) != (d_ref <= 0)
        stats["items"] += len(q)
        stats["mismatch_far"] += int((bad & (np.abs(d_ref) > 1e-6)).sum())
        stats["mismatch_near"] += int((bad & (np.abs(d_ref) <= 1e-6)).sum())
        stats["max_iters"] = max(stats["max_iters"], int(iters.max()))
        if want_dist:
            free = d_ref > 1e-6
            if free.any():
                stats["worst_dist_err"] = max(stats["worst_dist_err"], float(np.abs(np.maximum(dist - radii, 0)[free] - d_ref[free]).max()))
    return stats


@pytest.mark.parametrize("want_dist", [False, True])
def test_fp32_gjk_panda(want_dist):
    model, cmodel, _ = make_panda()
    q = uniform_states(model, 400, 17).astype(np.float64)
    s = _pair_sweep(model, cmodel, q, want_dist)
    assert s["items"] == 400 * 74
    assert s["mismatch_far"] == 0, s          # identical verdicts away from the 1e-6 m boundary band
    assert s["max_iters"] < 40, s
    if want_dist:
        assert s["worst_dist_err"] < 1e-5, s  # distances within 1e-5 m (north_star tolerance)


@pytest.mark.parametrize("want_dist", [False, True])
def test_fp32_gjk_two_dof_planar(want_dist):
    # planar model: flat tetrahedra, sliver triangles and curved (cylinder) supports
    model, cmodel, _ = make_two_dof()
    q = uniform_states(model, 10000, 0).astype(np.float64)
    s = _pair_sweep(model, cmodel, q, want_dist)
    assert s["mismatch_far"] == 0, s
    if want_dist:
        assert s["worst_dist_err"] < 1e-5, s


def test_fp32_gjk_near_contact_band():
    """Box-box items squeezed to a few micrometres of clearance / penetration: the verdict may
    only differ from float64 inside the 1e-6 m band."""
    from pyroboplan_b200.model import Box, GeometryModel, GeometryObject, Model, SE3
    from pyroboplan_b200.model import CollisionPair, JOINT_PRISMATIC

    model = Model("slider")
    model.addJoint(0, JOINT_PRISMATIC, [1, 0, 0], SE3.Identity(), "slide", -1.0, 1.0)
    cm = GeometryModel()
    cm.addGeometryObject(GeometryObject("a", 0, SE3(np.eye(3), np.array([0.0, 0.0, 0.0])), Box(0.2, 0.3, 0.4)))
    rot = np.array([[0.8, -0.6, 0.0], [0.6, 0.8, 0.0], [0.0, 0.0, 1.0]])
    cm.addGeometryObject(GeometryObject("b", 1, SE3(rot, np.array([0.0, 0.02, 0.03])), Box(0.1, 0.1, 0.1)))
    cm.addCollisionPair(CollisionPair(0, 1))
    orc = Oracle(model, cm)
    arrays, _ = flatten_model(model, cm)
    # contact happens near x0: find it by bisection on the oracle, then sample +-20 um around it
    lo, hi = 0.0, 0.5
    for _ in range(60):
        mid = 0.5 * (lo + hi)
        if orc.check_state(np.array([mid])):
            lo = mid
        else:
            hi = mid
    qs = (lo + np.linspace(-2e-5, 2e-5, 401)).astype(np.float32).astype(np.float64).reshape(-1, 1)
    oMg = np.array([orc.fk(x)[1] for x in qs])
    T = emu.relative_placements(oMg, 0, 1)
    hit, dist, _ = emu.gjk_batch(arrays, 0, 1, T, -T[:, 9:], 0.0, False)
    d_ref = np.array([orc.pair_distance(x, 0, 1)[0] for x in qs])
    hit_ref = d_ref <= 0
    bad = hit.astype(bool) != hit_ref
    # penetrating side: clearance is reported as 0 by the oracle, so measure the band in q instead
    assert np.all(np.abs(qs[bad, 0] - lo) < 2e-6), (qs[bad, 0] - lo)


def test_support_maps_are_exact():
    """The per-hull support maps (pbp_supmap.h) return the same support VALUE as a scan of every
    vertex, for random directions and for directions on cell boundaries / coordinate axes."""
    model, cmodel, _ = make_panda()
    arrays, _ = flatten_model(model, cmodel)
    rng = np.random.default_rng(7)
    d = rng.normal(size=(60000, 3)).astype(np.float32)
    d[:2000] = np.round(d[:2000])                      # axis / face-diagonal directions
    d[2000:4000, 0] = np.abs(d[2000:4000, 1])          # |u| == 1 chart boundaries
    k = rng.integers(-4, 5, size=(2000, 2)).astype(np.float32) / 4.0
    d[4000:6000] = np.stack([np.ones(2000, dtype=np.float32), k[:, 0], k[:, 1]], axis=1)  # exact cell edges
    d = d[np.abs(d).sum(axis=1) > 0]
    for g in (0, 1, 5, 8, 9):
        entries, max_per_cell, mismatches = emu.supmap_check(arrays, g, d)
        assert mismatches == 0
        # 6 * 16 * 16 cells: ~2 candidates per cell instead of 100-150 vertices, never more than a handful
        assert entries / 1536 < 2.5 and max_per_cell <= 24, (entries, max_per_cell)


def test_support_maps_do_not_change_gjk_results():
    model, cmodel, _ = make_panda()
    arrays, _ = flatten_model(model, cmodel)
    orc = Oracle(model, cmodel)
    q = uniform_states(model, 150, 23).astype(np.float64)
    oMg = np.array([orc.fk(x)[1] for x in q])
    for a, b in ((1, 5), (0, 8), (14, 3), (12, 6), (5, 10)):
        T = emu.relative_placements(oMg, a, b)
        v0 = -T[:, 9:]
        radii = sum(float(arrays["geom_params"][g][0]) for g in (a, b) if arrays["geom_shape"][g] in (0, 3))
        h1, d1, _ = emu.gjk_batch(arrays, a, b, T, v0, radii, True, use_maps=True)
        h0, d0, _ = emu.gjk_batch(arrays, a, b, T, v0, radii, True, use_maps=False)
        assert (h1 == h0).all()
        np.testing.assert_allclose(d1, d0, atol=2e-6)
