"""The engine's FP32 GJK with witness tracking (pbp_gjk.cuh + pbp_witness.cuh, host build) on every
Panda / 2-DOF pair: distances equal the float64 oracle's within 1e-5 m, and each witness pair is
proved by GJK-independent certificates (points inside their shapes, |p2 - p1| = distance, the slab
between the supporting planes along the normal is exactly that wide)."""

import numpy as np
import pytest

import emu
from conftest import make_panda, make_two_dof, uniform_states
from oracle.oracle import Oracle
from pyroboplan_b200.engine import flatten_model
from test_oracle_golden import _contains, _support_extent

pytestmark = pytest.mark.skipif(emu.cuda_include() is None, reason="CUDA headers not available")


def world(T, p):
    return np.einsum("nij,nj->ni", T[:, :9].reshape(-1, 3, 3), p) + T[:, 9:]


@pytest.mark.parametrize("which", ["panda", "two_dof"])
def test_witness_points_certified(which):
    model, cmodel, _ = make_panda() if which == "panda" else make_two_dof()
    arrays, _ = flatten_model(model, cmodel)
    orc = Oracle(model, cmodel)
    geoms = cmodel.geometryObjects
    q = uniform_states(model, 60 if which == "panda" else 600, 31).astype(np.float64)
    oMg = np.array([orc.fk(x)[1] for x in q])
    checked = overlapping = deep = 0
    for p in cmodel.collisionPairs:
        a, b = p.first, p.second
        T = emu.relative_placements(oMg, a, b)
        ca = np.asarray(geoms[a].geometry.bounding_sphere()[0])
        cb = np.asarray(geoms[b].geometry.bounding_sphere()[0])
        v0 = ca - (np.einsum("nij,j->ni", T[:, :9].reshape(-1, 3, 3), cb) + T[:, 9:])
        w = emu.gjk_witness_batch(arrays, a, b, T, v0)
        assert not w["far"].any()   # no cutoff: every pair is resolved
        p1 = world(oMg[:, a], w["pa"].astype(np.float64))
        p2 = world(oMg[:, a], w["pb"].astype(np.float64))
        nw = np.einsum("nij,nj->ni", oMg[:, a, :9].reshape(-1, 3, 3), w["n"].astype(np.float64))
        for i in range(len(q)):
            d_ref, pa_ref, pb_ref, inter = orc.pair_distance(q[i], a, b)
            if inter or w["hit"][i]:
                # overlapping: coal's convention, negative distance = penetration depth (FP32 EPA vs the oracle's)
                overlapping += 1
                ds, pas, pbs, ns, exact = orc.pair_signed_distance(q[i], a, b)
                if not exact or abs(ds) < 1e-5:
                    assert abs(float(w["dist"][i])) < 2e-5 or not exact
                    continue
                tol = 2e-5 if which == "panda" else 2e-3     # curved supports (2-DOF cylinder): the polytope has 40 vertices at most
                assert abs(float(w["dist"][i]) - ds) < tol, (a, b, i, float(w["dist"][i]), ds)
                # depth along the reported normal (GJK/EPA-free certificate), witnesses in their shapes
                width = _support_extent(geoms[a], oMg[i, a], nw[i]) + _support_extent(geoms[b], oMg[i, b], -nw[i])
                assert -ds - tol <= width <= -ds + 30 * tol
                assert _contains(geoms[a], oMg[i, a], p1[i], 5e-6) and _contains(geoms[b], oMg[i, b], p2[i], 5e-6)
                np.testing.assert_allclose(p2[i] - p1[i], nw[i] * float(w["dist"][i]), atol=5e-6)
                deep += 1
                continue
            assert abs(float(w["dist"][i]) - d_ref) < 1e-5
            assert _contains(geoms[a], oMg[i, a], p1[i], 2e-6) and _contains(geoms[b], oMg[i, b], p2[i], 2e-6)
            assert np.linalg.norm(p2[i] - p1[i]) == pytest.approx(d_ref, abs=1e-5)
            assert np.linalg.norm(nw[i]) == pytest.approx(1.0, abs=1e-5)
            if d_ref > 1e-3:   # the normal points from object 1 to object 2, as the oracle's witnesses do
                np.testing.assert_allclose(nw[i], (pb_ref - pa_ref) / d_ref, atol=2e-3 + 2e-5 / d_ref)
            gap = -_support_extent(geoms[b], oMg[i, b], -nw[i]) - _support_extent(geoms[a], oMg[i, a], nw[i])
            # the slab along the reported normal: a direction error theta costs about lever * theta of
            # width, and an eps-optimal witness pair pins the direction only to sqrt(2 eps / d)
            assert d_ref - 3e-4 <= gap <= d_ref + 1e-5
            checked += 1
    assert checked > 2000 and overlapping > 0 and deep > 20


def test_witness_cutoff_only_skips_far_pairs():
    model, cmodel, _ = make_panda()
    arrays, _ = flatten_model(model, cmodel)
    orc = Oracle(model, cmodel)
    q = uniform_states(model, 80, 33).astype(np.float64)
    oMg = np.array([orc.fk(x)[1] for x in q])
    for a, b in ((1, 5), (11, 6), (12, 4), (14, 8)):
        T = emu.relative_placements(oMg, a, b)
        w_all = emu.gjk_witness_batch(arrays, a, b, T, -T[:, 9:])
        w_cut = emu.gjk_witness_batch(arrays, a, b, T, -T[:, 9:], cutoff=0.05)
        far = w_cut["far"]
        assert (w_all["dist"][far] > 0.05).all() and (w_cut["dist"][far] > 0.05).all()
        assert (w_cut["dist"][far] <= w_all["dist"][far] + 1e-6).all()     # a lower bound
        near = w_all["dist"] <= 0.05
        assert not far[near].any()
        np.testing.assert_array_equal(w_cut["dist"][~far], w_all["dist"][~far])
        np.testing.assert_array_equal(w_cut["pa"][~far], w_all["pa"][~far])
