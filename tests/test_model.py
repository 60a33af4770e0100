"""Host model containers vs the conventions the reference relies on."""

import os

import numpy as np
import pytest

from conftest import make_panda, make_two_dof
from pyroboplan_b200.core.utils import (
    get_collision_geometry_ids,
    get_collision_pair_indices_from_bodies,
    set_collisions,
)
from pyroboplan_b200.model import SE3, Convex, rpy_to_matrix

REF_MODELS = "/root/reference/src/pyroboplan/models"


def test_panda_dimensions():
    model, cmodel, _ = make_panda(False, False)
    assert model.nq == 9  # reference test/core/test_utils.py:65
    assert model.njoints == 10
    assert model.names[1:] == [f"panda_joint{i}" for i in range(1, 8)] + ["panda_finger_joint1", "panda_finger_joint2"]
    np.testing.assert_allclose(model.lowerPositionLimit, [-2.9671, -1.8326, -2.9671, -3.1416, -2.9671, -0.0873, -2.9671, 0.0, 0.0])
    np.testing.assert_allclose(model.upperPositionLimit, [2.9671, 1.8326, 2.9671, 0.0873, 2.9671, 3.8223, 2.9671, 0.04, 0.04])
    assert cmodel.ngeoms == 11
    assert len(cmodel.collisionPairs) == 0


def test_geometry_names_and_ids():
    # reference test/core/test_utils.py:215-232
    model, cmodel, _ = make_panda(False, False)
    assert get_collision_geometry_ids(model, cmodel, "panda_link1_0") == [cmodel.getGeometryId("panda_link1_0")]
    assert get_collision_geometry_ids(model, cmodel, "panda_link3") == [cmodel.getGeometryId("panda_link3_0")]
    assert get_collision_geometry_ids(model, cmodel, "panda_link1000") == []
    assert cmodel.getGeometryId("nope") == cmodel.ngeoms
    assert model.getFrameId("nope") == model.nframes


def test_set_collisions_pair_ids():
    # reference test/core/test_utils.py:234-246: geometry id == link index
    model, cmodel, _ = make_panda(False, False)
    set_collisions(model, cmodel, "panda_link4", "panda_link5", True)
    assert len(cmodel.collisionPairs) == 1
    assert cmodel.collisionPairs[0].first == 4
    assert cmodel.collisionPairs[0].second == 5
    set_collisions(model, cmodel, "panda_link4", "panda_link5", True)  # no duplicates
    assert len(cmodel.collisionPairs) == 1
    set_collisions(model, cmodel, "panda_link5", "panda_link4", False)  # unordered equality
    assert len(cmodel.collisionPairs) == 0


def test_panda_pair_list():
    # 20 self pairs after the SRDF, then 54 obstacle pairs = 74
    model, cmodel, vmodel = make_panda(True, False)
    self_pairs = [(p.first, p.second) for p in cmodel.collisionPairs]
    expected = [(a, b) for a in (0, 1, 2) for b in (5, 6, 7, 8, 9, 10)] + [(5, 9), (5, 10)]
    assert self_pairs == expected
    model, cmodel, vmodel = make_panda(True, True)
    assert cmodel.ngeoms == 16
    assert len(cmodel.collisionPairs) == 74
    obstacle_pairs = [(p.first, p.second) for p in cmodel.collisionPairs[20:]]
    names = [g.name for g in cmodel.geometryObjects]
    assert names[11:] == ["ground_plane", "obstacle_sphere_1", "obstacle_sphere_2", "obstacle_box_1", "obstacle_box_2"]
    # insertion order: ground, box_1, box_2, sphere_1, sphere_2, each against links 0..10; (ground, link0) removed
    exp = [(11, l) for l in range(1, 11)]
    for obs in (14, 15, 12, 13):
        exp += [(obs, l) for l in range(11)]
    assert obstacle_pairs == exp
    idx = get_collision_pair_indices_from_bodies(model, cmodel, ["ground_plane"])
    assert idx == list(range(20, 30))


def test_two_dof_model():
    model, cmodel, _ = make_two_dof(False)
    assert model.nq == 2 and cmodel.ngeoms == 2 and len(cmodel.collisionPairs) == 0
    model, cmodel, _ = make_two_dof(True)
    assert [g.name for g in cmodel.geometryObjects] == ["upperarm_0", "forearm_0", "obstacle_1", "obstacle_2"]
    assert [(p.first, p.second) for p in cmodel.collisionPairs] == [(2, 0), (2, 1), (3, 0), (3, 1)]
    np.testing.assert_allclose(cmodel.geometryObjects[0].placement.translation, [0.5, 0, 0])


def test_sphere_model():
    from pyroboplan_b200.models import panda

    model, cmodel, _ = panda.load_models(use_sphere_collisions=True)
    assert cmodel.ngeoms == 17 and model.nq == 9


def test_hand_placement_and_meshes():
    model, cmodel, _ = make_panda(False, False)
    hand = cmodel.geometryObjects[8]
    assert hand.name == "panda_hand_0" and hand.parentJoint == 7
    np.testing.assert_allclose(hand.placement.translation, [0, 0, 0.107])
    np.testing.assert_allclose(hand.placement.rotation, rpy_to_matrix(0, 0, -0.785398163397), atol=1e-12)
    rf = cmodel.geometryObjects[10]
    np.testing.assert_allclose(rf.placement.rotation, rpy_to_matrix(0, 0, 3.14159265359), atol=1e-12)
    counts = [len(g.geometry.points) for g in cmodel.geometryObjects]
    assert counts == [102, 152, 152, 152, 152, 152, 102, 102, 102, 18, 18]
    for g in cmodel.geometryObjects:
        assert isinstance(g.geometry, Convex)
        assert g.geometry.volume_ratio > 0.999  # the Panda collision meshes are convex to < 0.03 % volume
        c, r = g.geometry.bounding_sphere()
        assert np.all(np.linalg.norm(g.geometry.points - c, axis=1) <= r + 1e-12)
        ci, ri = g.geometry.inner_sphere()
        assert ri > 0 and np.all(g.geometry.planes[:, :3] @ ci + g.geometry.planes[:, 3] <= -ri + 1e-9)


def test_se3():
    rng = np.random.default_rng(0)
    a = SE3(rpy_to_matrix(*rng.normal(size=3)), rng.normal(size=3))
    b = SE3(rpy_to_matrix(*rng.normal(size=3)), rng.normal(size=3))
    p = rng.normal(size=3)
    np.testing.assert_allclose((a * b).act(p), a.act(b.act(p)), atol=1e-12)
    np.testing.assert_allclose((a * a.inverse()).homogeneous, np.eye(4), atol=1e-12)
    np.testing.assert_allclose(a.actInv(a.act(p)), p, atol=1e-12)


@pytest.mark.skipif(not os.path.isdir(REF_MODELS), reason="reference assets only exist in the authoring container")
def test_urdf_parser_matches_bundled_description():
    """The URDF/SRDF/STL path and the bundled JSON descriptions give the same models."""
    from pyroboplan_b200.model import buildModelsFromUrdf, removeCollisionPairs

    model, cmodel, _ = buildModelsFromUrdf(os.path.join(REF_MODELS, "panda_description/urdf/panda.urdf"), package_dirs=REF_MODELS)
    cmodel.addAllCollisionPairs()
    assert len(cmodel.collisionPairs) == 54
    removeCollisionPairs(model, cmodel, os.path.join(REF_MODELS, "panda_description/srdf/panda.srdf"))
    ref_model, ref_cmodel, _ = make_panda(True, False)
    assert [(p.first, p.second) for p in cmodel.collisionPairs] == [(p.first, p.second) for p in ref_cmodel.collisionPairs]
    for g, h in zip(cmodel.geometryObjects, ref_cmodel.geometryObjects):
        assert g.name == h.name and g.parentJoint == h.parentJoint and g.parentFrame == h.parentFrame
        np.testing.assert_allclose(g.geometry.points, h.geometry.points, rtol=0, atol=1e-9)
    for a, b in zip(model.jointPlacements, ref_model.jointPlacements):
        np.testing.assert_allclose(a.homogeneous, b.homogeneous, atol=1e-15)


def test_marble_and_sphere_panda_models():
    """Reference models/marble.py:28-78 (10 box-marble pairs, the ground plate is not paired) and
    models/panda.py:13-27 with use_sphere_collisions (panda_spheres.urdf: 17 link spheres)."""
    from pyroboplan_b200.models import marble, panda

    m, c, v = marble.load_models()
    assert m.nq == 2 and [g.name for g in c.geometryObjects] == ["body_0"]
    marble.add_object_collisions(m, c, v)
    assert len(c.geometryObjects) == 12 and len(c.collisionPairs) == 10
    body = c.getGeometryId("body_0")
    assert all(p.second == body and c.geometryObjects[p.first].name.startswith("box") for p in c.collisionPairs)
    np.testing.assert_array_equal(m.lowerPositionLimit, [0.0, 0.0])
    np.testing.assert_array_equal(m.upperPositionLimit, [2.0, 1.0])
    m, c, v = panda.load_models(use_sphere_collisions=True)
    assert m.nq == 9 and len(c.geometryObjects) == 17
    assert all(type(g.geometry).__name__ == "Sphere" for g in c.geometryObjects)


def test_enclosing_capsules_contain_their_hulls():
    """The broadphase's capsule proxies (engine.enclosing_capsule) must contain every hull vertex, in
    the float32 numbers the device sees; only clearly elongated links get one."""
    from pyroboplan_b200.engine import flatten_model

    model, cmodel, _ = make_panda()
    arrays, _ = flatten_model(model, cmodel)
    flagged = []
    for g, obj in enumerate(cmodel.geometryObjects):
        cap = arrays["geom_capsule"][g].astype(np.float64)
        if type(obj.geometry).__name__ != "Convex":
            assert not cap.any()
            continue
        pts = np.asarray(obj.geometry.points, dtype=np.float32).astype(np.float64)
        p0, p1, r = cap[:3], cap[3:6], cap[6]
        u = p1 - p0
        s = np.clip(((pts - p0) @ u) / max(u @ u, 1e-30), 0.0, 1.0)
        assert (np.linalg.norm(pts - (p0 + np.outer(s, u)), axis=1) <= r).all()
        if cap[7]:
            flagged.append(obj.name)
            assert r < 0.55 * arrays["geom_outer"][g][3]
    assert flagged == ["panda_link5_0", "panda_hand_0"]


def test_surface_tables_of_the_panda():
    """pbp_model_desc surface tables (engine.flatten_model): the hull planes contain every hull vertex and are
    sorted by face area, hull and inner core differ by less than 2 mm, and the enclosure flags name exactly the pairs
    whose second geometry is small enough to fit inside the first (never a primitive as the enclosing one)."""
    from pyroboplan_b200.engine import flatten_model
    from pyroboplan_b200.models import panda, two_dof

    model, cmodel, vmodel = panda.load_models()
    panda.add_self_collisions(model, cmodel)
    panda.add_object_collisions(model, cmodel, vmodel)
    a, sizes = flatten_model(model, cmodel)
    geoms = cmodel.geometryObjects
    assert sizes["nplanes"] == int(a["geom_plane_cnt"].sum()) > 0
    for g, obj in enumerate(geoms):
        cnt, off = int(a["geom_plane_cnt"][g]), int(a["geom_plane_off"][g])
        if type(obj.geometry).__name__ != "Convex":
            assert cnt == 0 and a["geom_surface_tol"][g] == 0
            continue
        pl = a["planes"][off:off + cnt].astype(np.float64)
        v = a["verts"][a["geom_vert_off"][g]: a["geom_vert_off"][g] + a["geom_vert_cnt"][g]].astype(np.float64)
        nh = int(a["geom_plane_hull_cnt"][g])
        assert nh == len(obj.geometry.planes) and nh < cnt          # hull faces first, then mesh triangles below the hull
        np.testing.assert_allclose(np.linalg.norm(pl[:, :3], axis=1), 1.0, atol=1e-6)
        slack = (v @ pl[:nh, :3].T - pl[:nh, 3]).max(axis=0)
        assert slack.max() < 2e-6 and slack.min() > -2e-6          # every hull plane supports the hull: it touches a vertex
        cut = (v @ pl[nh:, :3].T - pl[nh:, 3]).max(axis=0)
        assert cut.min() > -1e-7 and cut.max() < 2e-3              # the other planes cut a sliver off the hull (or touch it), at most ~1 mm deep
        assert 0.0 < a["geom_surface_tol"][g] < 1e-3                # hull boundary <-> mesh surface: sub-millimetre (link 3: 0.39 mm dents)
        assert 0.0 < a["geom_core_eps"][g] < 2e-3
    flagged = {(geoms[p.first].name, geoms[p.second].name): int(a["pair_enclose"][k]) for k, p in enumerate(cmodel.collisionPairs) if a["pair_enclose"][k]}
    assert len(flagged) == 17 and set(flagged.values()) == {1}
    for small in ("panda_leftfinger_0", "panda_rightfinger_0"):
        for big in ("panda_link0_0", "panda_link1_0", "panda_link2_0", "panda_link5_0"):
            assert (big, small) in flagged
    assert all("obstacle" not in x and "ground" not in x for pair in flagged for x in pair)
    # hull semantics on request; models without meshes never carry surface tables
    a_h, s_h = flatten_model(model, cmodel, mesh_semantics="hull")
    assert s_h["nplanes"] == 0 and not a_h["pair_enclose"].any()
    m2, c2, v2 = two_dof.load_models()
    two_dof.add_object_collisions(m2, c2, v2)
    assert flatten_model(m2, c2)[1]["nplanes"] == 0
    with pytest.raises(ValueError):
        flatten_model(model, cmodel, mesh_semantics="soup")


def test_ur_models_mirror_the_reference_modules():
    """reference models/ur5.py:9-47, ur10.py:9-47: 12 collision geometries, all pairs minus the SRDF's."""
    from pyroboplan_b200.engine import flatten_model
    from pyroboplan_b200.models import ur5, ur10

    for load, add in ((ur5.load_ur5_on_base_models, ur5.add_ur5_on_base_self_collisions), (ur10.load_ur10_on_base_models, ur10.add_ur10_on_base_self_collisions)):
        m, c, _ = load()
        assert m.nq == 9 and len(c.collisionPairs) == 0
        add(m, c)
        names = [g.name for g in c.geometryObjects]
        assert names[:8] == ["mobile_base_0", "base_link_0", "shoulder_link_0", "upper_arm_link_0", "forearm_link_0", "wrist_1_link_0", "wrist_2_link_0", "wrist_3_link_0"]
        assert len(c.collisionPairs) == 37
        pairs = {(names[p.first], names[p.second]) for p in c.collisionPairs}
        assert ("upper_arm_link_0", "forearm_link_0") not in pairs and ("base_link_0", "forearm_link_0") in pairs
        a, sizes = flatten_model(m, c)
        assert sizes["ntris"] > 5000 and (a["geom_tri_cnt"][1:8] > 300).all()
        # non-convex meshes: the inner cell is far inside the hull (or missing) -- the triangles decide
        assert (a["geom_core_eps"][1:8] > 0.05).sum() + (a["geom_core_cnt"][1:8] == 0).sum() >= 4
