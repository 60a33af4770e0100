"""``pinocchio.Model`` / ``pinocchio.GeometryModel`` -> this package's host containers.

The reference passes Pinocchio objects to every function of the hot path
(/root/reference/src/pyroboplan/core/utils.py:8-15, models/panda.py:27,78-79).  The engine compiles its
device tables from this package's own containers (``model/robot_model.py``); :func:`from_pinocchio` builds
those containers from a real Pinocchio model by reading nothing but public attributes, so the drop-in
functions accept either kind of model:

=========================================  =======================================================
Pinocchio / coal attribute                 becomes
=========================================  =======================================================
``model.joints[i].shortname()``            joint type + axis (``JointModelR{X,Y,Z}``, ``P{X,Y,Z}``,
                                           ``RevoluteUnaligned`` / ``PrismaticUnaligned`` via ``.extract().axis``)
``model.parents / jointPlacements``        tree + fixed placements (``.rotation``, ``.translation``)
``model.joints[i].idx_q / idx_v``          checked: one coordinate per joint, in joint order
``model.frames[i]``                        ``.name .parentJoint (.parent) .parentFrame .placement .type``
``model.lower/upperPositionLimit``         limits
``gm.geometryObjects[i].geometry``         by coal class name: ``Box .halfSide``, ``Sphere .radius``,
                                           ``Cylinder / Capsule .radius .halfLength``, ``Convex*`` points (a
                                           solid), ``BVHModel*`` ``vertices() / tri_indices(i)`` (a SURFACE:
                                           hull + inner core + triangles, model/surface.py)
``gm.geometryObjects[i].disableCollision`` pairs of a disabled geometry never collide: dropped
``gm.collisionPairs[k].first / .second``   active pairs, in order
=========================================  =======================================================

Pinocchio is not importable in the authoring container nor on the GPU box; the adapter is duck-typed and
unit-tested on stand-ins carrying Pinocchio's attribute names (tests/test_pinocchio_adapter.py), and
``tests/test_pinocchio_crosscheck.py`` runs it against the real thing wherever ``import pinocchio`` works.
"""

import hashlib

import numpy as np

from .robot_model import (
    JOINT_PRISMATIC,
    JOINT_REVOLUTE,
    CollisionPair,
    Frame,
    GeometryModel,
    GeometryObject,
    Model,
)
from .se3 import SE3
from .shapes import Box, Capsule, Convex, Cylinder, Sphere

_AXES = {"X": (1.0, 0.0, 0.0), "Y": (0.0, 1.0, 0.0), "Z": (0.0, 0.0, 1.0)}


def is_native_model(model):
    """True for this package's own containers (anything else is taken for a Pinocchio object)."""
    return isinstance(model, (Model, GeometryModel))


def _se3(M):
    return SE3(np.array(M.rotation, dtype=np.float64).reshape(3, 3), np.array(M.translation, dtype=np.float64).reshape(3))


def _joint_kind(jm, name):
    """(joint type code, unit axis) of a Pinocchio JointModel; raises for joints with nq != 1."""
    sn = jm.shortname()
    base = sn[len("JointModel"):] if sn.startswith("JointModel") else sn
    if base in ("RX", "RY", "RZ"):
        return JOINT_REVOLUTE, np.array(_AXES[base[1]])
    if base in ("PX", "PY", "PZ"):
        return JOINT_PRISMATIC, np.array(_AXES[base[1]])
    if base in ("RevoluteUnaligned", "PrismaticUnaligned"):
        inner = jm.extract() if hasattr(jm, "extract") else jm
        axis = np.array(inner.axis, dtype=np.float64).reshape(3)
        return (JOINT_REVOLUTE if base.startswith("Revolute") else JOINT_PRISMATIC), axis / np.linalg.norm(axis)
    raise NotImplementedError(
        f"joint {name!r}: {sn} is not supported by the collision engine (one-coordinate revolute / prismatic joints only; "
        "continuous joints with a (cos, sin) configuration, free-flyer, spherical, composite and mimic joints are not)"
    )


def _convex_points(g):
    pts = g.points
    pts = pts() if callable(pts) else pts
    if pts is None or np.ndim(pts) != 2:   # older bindings: points(i)
        pts = [np.array(g.points(i)) for i in range(int(g.num_points))]
    return np.array(pts, dtype=np.float64).reshape(-1, 3)


def _bvh_mesh(g):
    verts = g.vertices
    verts = verts() if callable(verts) else verts
    if verts is None or np.ndim(verts) != 2:
        verts = [np.array(g.vertices(i)) for i in range(int(g.num_vertices))]
    verts = np.array(verts, dtype=np.float64).reshape(-1, 3)
    tris = np.empty((int(g.num_tris), 3), dtype=np.int64)
    for i in range(len(tris)):
        t = g.tri_indices(i)
        tris[i] = (int(t[0]), int(t[1]), int(t[2]))
    return verts, tris


def convert_geometry(g, name=""):
    """coal / hpp-fcl ``CollisionGeometry`` -> this package's shape, dispatched on the class name."""
    cls = type(g).__name__
    if cls == "Box":
        h = np.array(g.halfSide, dtype=np.float64).reshape(3)
        return Box(2.0 * h[0], 2.0 * h[1], 2.0 * h[2])
    if cls == "Sphere":
        return Sphere(float(g.radius))
    if cls == "Cylinder":
        return Cylinder(float(g.radius), 2.0 * float(g.halfLength))
    if cls == "Capsule":
        return Capsule(float(g.radius), 2.0 * float(g.halfLength))
    if cls.startswith("BVHModel"):
        verts, tris = _bvh_mesh(g)
        return Convex(verts, source=name, triangles=tris)        # a surface: triangles kept
    if cls.startswith("Convex"):
        return Convex(_convex_points(g), source=name)            # a solid
    raise NotImplementedError(f"geometry {name!r}: coal class {cls} is not supported (Box, Sphere, Cylinder, Capsule, Convex, BVHModel*)")


def from_pinocchio(model, collision_model=None):
    """Build this package's ``(Model, GeometryModel)`` from Pinocchio objects (``collision_model`` optional).

    The result has the same joint, frame, geometry and pair indices as the input, so ids returned by either
    side mean the same thing.  Pairs touching a geometry with ``disableCollision`` are dropped (Pinocchio
    skips them in ``computeCollisions``)."""
    out = Model(getattr(model, "name", "model"))
    out.frames = []
    names = list(model.names)
    nq = 0
    for j in range(1, int(model.njoints)):
        jm = model.joints[j]
        kind, axis = _joint_kind(jm, names[j])
        if int(jm.idx_q) != nq or int(jm.idx_v) != nq:
            raise NotImplementedError(f"joint {names[j]!r}: configuration index {int(jm.idx_q)} is not its position in the joint order ({nq})")
        out.addJoint(int(model.parents[j]), kind, axis, _se3(model.jointPlacements[j]), names[j],
                     float(model.lowerPositionLimit[nq]), float(model.upperPositionLimit[nq]),
                     float(model.velocityLimit[nq]) if hasattr(model, "velocityLimit") else np.inf,
                     float(model.effortLimit[nq]) if hasattr(model, "effortLimit") else np.inf)
        nq += 1
    if nq != int(model.nq):
        raise NotImplementedError(f"model has nq = {int(model.nq)} but {nq} one-coordinate joints")
    for f in model.frames:
        parent = f.parentJoint if hasattr(f, "parentJoint") else f.parent
        pf = getattr(f, "parentFrame", getattr(f, "previousFrame", 0))
        out.frames.append(Frame(f.name, int(parent), _se3(f.placement), int(f.type), int(pf)))
    gm = None
    if collision_model is not None:
        gm = GeometryModel()
        disabled = set()
        for i, go in enumerate(collision_model.geometryObjects):
            pf = int(go.parentFrame) if hasattr(go, "parentFrame") else GeometryObject.UNSET_FRAME
            obj = GeometryObject(go.name, int(go.parentJoint), pf, _se3(go.placement), convert_geometry(go.geometry, go.name))
            gm.geometryObjects.append(obj)
            if bool(getattr(go, "disableCollision", False)):
                disabled.add(i)
        for p in collision_model.collisionPairs:
            a, b = int(p.first), int(p.second)
            if a in disabled or b in disabled:
                continue
            gm.collisionPairs.append(CollisionPair(a, b))
    return out, gm


def fingerprint(model, collision_model):
    """Content hash of what the engine compiles from (placements, limits, shape parameters, pairs): two
    calls agree iff the compiled device model would be the same.  Costs a walk over the model in Python
    (tens of microseconds to milliseconds): `engine.get_engine` uses it on request, not on every call."""
    h = hashlib.sha1()

    def put(x):
        h.update(np.ascontiguousarray(np.asarray(x, dtype=np.float64)).tobytes())

    put([int(model.njoints), int(model.nq)])
    put(list(model.parents))
    for M in model.jointPlacements:
        put(M.rotation), put(M.translation)
    put(model.lowerPositionLimit), put(model.upperPositionLimit)
    for j in range(1, int(model.njoints)):
        jm = model.joints[j]
        if hasattr(jm, "shortname"):
            h.update(jm.shortname().encode())
            if "Unaligned" in jm.shortname():
                put((jm.extract() if hasattr(jm, "extract") else jm).axis)
        else:
            put([jm.type]), put(jm.axis)
    if collision_model is not None:
        for go in collision_model.geometryObjects:
            g = go.geometry
            h.update(type(g).__name__.encode())
            put([int(go.parentJoint)]), put(go.placement.rotation), put(go.placement.translation)
            for attr in ("halfSide", "radius", "halfLength"):
                if hasattr(g, attr):
                    put(getattr(g, attr))
            for attr in ("num_vertices", "num_tris", "num_points"):
                if hasattr(g, attr) and not callable(getattr(g, attr)):
                    put([int(getattr(g, attr))])
            if hasattr(g, "volume") and not callable(g.volume):
                put([g.volume])
            put([bool(getattr(go, "disableCollision", False))])
        put([[int(p.first), int(p.second)] for p in collision_model.collisionPairs])
    return h.hexdigest()
