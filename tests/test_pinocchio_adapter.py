"""from_pinocchio on stand-ins that carry Pinocchio's / coal's attribute names (no Pinocchio here).

The stand-ins are built FROM this package's Panda / 2-DOF containers, pushed through the adapter, and
the flattened device tables must come out identical -- so whatever the adapter reads (shortname(),
jointPlacements[i].rotation, geometry.halfSide, BVH vertices()/tri_indices(i), collisionPairs[k].first
...) is read with Pinocchio's spelling and semantics (half sizes, full cylinder length via halfLength).
Reference call sites: /root/reference/src/pyroboplan/core/utils.py:8-15, models/panda.py:27,78-79,98-154.
"""

import numpy as np
import pytest

from pyroboplan_b200.engine import flatten_model
from pyroboplan_b200.model import Box, Capsule, Convex, Cylinder, Sphere
from pyroboplan_b200.model import pinocchio_adapter as pa
from pyroboplan_b200.models import panda, two_dof


class FakeSE3:
    def __init__(self, T):
        self.rotation = np.array(T.rotation)
        self.translation = np.array(T.translation)


class FakeJointInner:
    def __init__(self, axis):
        self.axis = np.array(axis)


class FakeJoint:
    def __init__(self, short, idx_q, axis=None):
        self._short, self.idx_q, self.idx_v, self.nq, self.nv, self._axis = short, idx_q, idx_q, 1, 1, axis

    def shortname(self):
        return self._short

    def extract(self):
        return FakeJointInner(self._axis)


class FakeFrame:
    def __init__(self, f):
        self.name, self.parentJoint, self.parentFrame, self.placement, self.type = f.name, f.parentJoint, f.parentFrame, FakeSE3(f.placement), f.type


class FakeModel:
    pass


class FakePair:
    def __init__(self, a, b):
        self.first, self.second = a, b




def _coal_like(shape):
    """An object whose class name and attributes are coal's for this shape."""
    if isinstance(shape, Box):
        return type("Box", (), {"halfSide": np.array(shape.halfSide)})()
    if isinstance(shape, Sphere):
        return type("Sphere", (), {"radius": shape.radius})()
    if isinstance(shape, Cylinder):
        return type("Cylinder", (), {"radius": shape.radius, "halfLength": shape.halfLength})()
    if isinstance(shape, Capsule):
        return type("Capsule", (), {"radius": shape.radius, "halfLength": shape.halfLength})()
    if isinstance(shape, Convex) and shape.mesh_triangles is not None:
        verts, tris = np.array(shape.mesh_vertices), np.array(shape.mesh_triangles)
        return type("BVHModelOBBRSS", (), {"num_vertices": len(verts), "num_tris": len(tris), "vertices": lambda self: verts,
                                           "tri_indices": lambda self, i: tuple(int(x) for x in tris[i])})()
    pts = np.array(shape.points)
    return type("ConvexBase", (), {"num_points": len(pts), "points": lambda self: pts})()


def fake_pinocchio(model, cmodel, unaligned=False):
    fm = FakeModel()
    fm.name, fm.njoints, fm.nq, fm.nv = model.name, model.njoints, model.nq, model.nv
    fm.names, fm.parents = list(model.names), list(model.parents)
    fm.jointPlacements = [FakeSE3(T) for T in model.jointPlacements]
    fm.lowerPositionLimit, fm.upperPositionLimit = np.array(model.lowerPositionLimit), np.array(model.upperPositionLimit)
    fm.velocityLimit, fm.effortLimit = np.array(model.velocityLimit), np.array(model.effortLimit)
    fm.joints = [FakeJoint("JointModelRX", -1)]   # the universe slot is never read
    for j in model.joints[1:]:
        k = int(np.argmax(np.abs(j.axis)))
        aligned = np.allclose(np.abs(j.axis), np.eye(3)[k]) and j.axis[k] > 0 and not unaligned
        rev = j.type == 1
        short = ("JointModelR" if rev else "JointModelP") + "XYZ"[k] if aligned else ("JointModelRevoluteUnaligned" if rev else "JointModelPrismaticUnaligned")
        fm.joints.append(FakeJoint(short, j.idx_q, j.axis))
    fm.frames = [FakeFrame(f) for f in model.frames]
    fc = FakeModel()
    fc.geometryObjects = []
    for g in cmodel.geometryObjects:
        o = FakeModel()
        o.name, o.parentJoint, o.parentFrame, o.placement, o.geometry, o.disableCollision = g.name, g.parentJoint, g.parentFrame, FakeSE3(g.placement), _coal_like(g.geometry), False
        fc.geometryObjects.append(o)
    fc.collisionPairs = [FakePair(p.first, p.second) for p in cmodel.collisionPairs]
    return fm, fc


def _panda():
    model, cmodel, vmodel = panda.load_models()
    panda.add_self_collisions(model, cmodel)
    panda.add_object_collisions(model, cmodel, vmodel)
    return model, cmodel


@pytest.mark.parametrize("unaligned", [False, True])
def test_panda_tables_identical_through_the_adapter(unaligned):
    model, cmodel = _panda()
    fm, fc = fake_pinocchio(model, cmodel, unaligned)
    assert not pa.is_native_model(fm) and pa.is_native_model(model)
    m2, c2 = pa.from_pinocchio(fm, fc)
    assert [f.name for f in m2.frames] == [f.name for f in model.frames]
    assert m2.getFrameId("panda_hand") == model.getFrameId("panda_hand")
    assert [(p.first, p.second) for p in c2.collisionPairs] == [(p.first, p.second) for p in cmodel.collisionPairs]
    a, sa = flatten_model(model, cmodel)
    b, sb = flatten_model(m2, c2)
    assert sa == sb
    for k in a:
        np.testing.assert_array_equal(a[k], b[k], err_msg=k)


def test_two_dof_primitives_and_disabled_geometry():
    model, cmodel, vmodel = two_dof.load_models()
    two_dof.add_object_collisions(model, cmodel, vmodel)
    fm, fc = fake_pinocchio(model, cmodel)
    m2, c2 = pa.from_pinocchio(fm, fc)
    a, _ = flatten_model(model, cmodel)
    b, _ = flatten_model(m2, c2)
    for k in a:
        np.testing.assert_array_equal(a[k], b[k], err_msg=k)
    # Pinocchio skips the pairs of a geometry with disableCollision in computeCollisions
    fc.geometryObjects[fc.collisionPairs[0].first].disableCollision = True
    gone = fc.collisionPairs[0].first
    _, c3 = pa.from_pinocchio(fm, fc)
    assert all(gone not in (p.first, p.second) for p in c3.collisionPairs)
    assert len(c3.collisionPairs) < len(c2.collisionPairs)


def test_unsupported_joint_and_geometry_raise():
    model, cmodel = _panda()
    fm, fc = fake_pinocchio(model, cmodel)
    fm.joints[3] = FakeJoint("JointModelRUBZ", fm.joints[3].idx_q)     # continuous joint: (cos, sin) configuration
    with pytest.raises(NotImplementedError, match="JointModelRUBZ"):
        pa.from_pinocchio(fm, fc)
    fm, fc = fake_pinocchio(model, cmodel)
    fc.geometryObjects[0].geometry = type("Halfspace", (), {})()
    with pytest.raises(NotImplementedError, match="Halfspace"):
        pa.from_pinocchio(fm, fc)


def test_fingerprint_sees_in_place_edits():
    model, cmodel = _panda()
    fm, fc = fake_pinocchio(model, cmodel)
    f0 = pa.fingerprint(fm, fc)
    assert f0 == pa.fingerprint(fm, fc) == pa.fingerprint(*fake_pinocchio(model, cmodel))
    fc.geometryObjects[12].placement.translation[2] += 0.01         # an obstacle moved in place
    assert pa.fingerprint(fm, fc) != f0
    assert pa.fingerprint(model, cmodel) == pa.fingerprint(model, cmodel)      # native containers hash too
