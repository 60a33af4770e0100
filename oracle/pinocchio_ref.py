"""Tier-A checker: the REAL reference arithmetic (Pinocchio + coal), used wherever ``import pinocchio`` works.

TEST INFRASTRUCTURE, like everything under oracle/: imported by tests/ and by bench.py's CPU legs only.
Pinocchio is importable neither in the authoring container nor on the GPU box of this project (no
wheel, no network), so nothing depends on this module; it exists so that the very first machine that
has Pinocchio upgrades the parity evidence from "restated" to "checked against the reference":

* :func:`available`        -- ``importlib.util.find_spec("pinocchio")`` probe (never raises)
* :func:`to_pinocchio`     -- this package's host containers -> real ``pinocchio.Model`` /
  ``GeometryModel`` (``JointModelR{X,Y,Z}`` / ``RevoluteUnaligned`` ..., ``coal.Box/Sphere/Cylinder/
  Capsule``, ``coal.BVHModelOBBRSS`` from the mesh triangles as ``buildModelsFromUrdf`` produces them)
* :func:`check_states`     -- the reference's ``check_collisions_at_state`` loop
  (/root/reference/src/pyroboplan/core/utils.py:8-51: security_margin on every request, then
  ``computeCollisions(..., stop_at_first_collision=True)``), the installed ``pyroboplan`` itself when it
  is importable
* :func:`min_distances`    -- ``get_minimum_distance_at_state`` (core/utils.py:54-87)
* :func:`timed_rate`       -- states/s of :func:`check_states` on one process or a pool (bench.py)
"""

import importlib.util
import os
import time

import numpy as np


def available():
    try:
        return importlib.util.find_spec("pinocchio") is not None
    except (ImportError, ValueError):
        return False


def _coal():
    try:
        import coal

        return coal
    except ImportError:
        import hppfcl

        return hppfcl


def to_pinocchio(model, cmodel):
    """(pinocchio.Model, pinocchio.GeometryModel) with the same joint / frame / geometry / pair indices."""
    import pinocchio as pin

    from pyroboplan_b200.model import Box, Capsule, Convex, Cylinder, Sphere

    coal = _coal()
    se3 = lambda T: pin.SE3(np.array(T.rotation, dtype=np.float64), np.array(T.translation, dtype=np.float64))
    pm = pin.Model()
    pm.name = model.name
    for j in range(1, model.njoints):
        jm_host = model.joints[j]
        ax = np.asarray(jm_host.axis, dtype=np.float64)
        k = int(np.argmax(np.abs(ax)))
        aligned = np.allclose(ax, np.eye(3)[k])
        if jm_host.type == 1:
            jm = (pin.JointModelRX, pin.JointModelRY, pin.JointModelRZ)[k]() if aligned else pin.JointModelRevoluteUnaligned(ax)
        else:
            jm = (pin.JointModelPX, pin.JointModelPY, pin.JointModelPZ)[k]() if aligned else pin.JointModelPrismaticUnaligned(ax)
        i = jm_host.idx_q
        lo, hi = np.array([model.lowerPositionLimit[i]]), np.array([model.upperPositionLimit[i]])
        jid = pm.addJoint(int(model.parents[j]), jm, se3(model.jointPlacements[j]), model.names[j],
                          np.array([model.effortLimit[i]]), np.array([model.velocityLimit[i]]), lo, hi)
        assert jid == j
    for f in model.frames[1:]:
        ftype = {1: pin.FrameType.OP_FRAME, 2: pin.FrameType.JOINT, 4: pin.FrameType.FIXED_JOINT, 8: pin.FrameType.BODY}[int(f.type)]
        pm.addFrame(pin.Frame(f.name, int(f.parentJoint), int(f.parentFrame), se3(f.placement), ftype))
    pcm = pin.GeometryModel()
    for g in cmodel.geometryObjects:
        s = g.geometry
        if isinstance(s, Box):
            geom = coal.Box(*(2.0 * np.asarray(s.halfSide)))
        elif isinstance(s, Sphere):
            geom = coal.Sphere(s.radius)
        elif isinstance(s, Cylinder):
            geom = coal.Cylinder(s.radius, 2.0 * s.halfLength)
        elif isinstance(s, Capsule):
            geom = coal.Capsule(s.radius, 2.0 * s.halfLength)
        elif isinstance(s, Convex) and s.mesh_triangles is not None:
            geom = coal.BVHModelOBBRSS()
            verts = np.asarray(s.mesh_vertices, dtype=np.float64)
            geom.beginModel(len(s.mesh_triangles), len(verts))
            for t in np.asarray(s.mesh_triangles):
                geom.addTriangle(verts[t[0]], verts[t[1]], verts[t[2]])
            geom.endModel()
        else:
            raise NotImplementedError(f"{g.name}: solid convex polytopes need coal.Convex construction from faces")
        pf = int(g.parentFrame) if g.parentFrame < model.nframes else 0
        try:
            obj = pin.GeometryObject(g.name, int(g.parentJoint), pf, se3(g.placement), geom)
        except Exception:   # Pinocchio 2 argument order
            obj = pin.GeometryObject(g.name, pf, int(g.parentJoint), geom, se3(g.placement))
        pcm.addGeometryObject(obj)
    for p in cmodel.collisionPairs:
        pcm.addCollisionPair(pin.CollisionPair(int(p.first), int(p.second)))
    return pm, pcm


def check_states(pm, pcm, Q, padding=0.0):
    """bool [N]: the reference's verdict per state."""
    import pinocchio as pin

    try:   # the reference itself, when it is installed next to Pinocchio
        from pyroboplan.core.utils import check_collisions_at_state as ref_check
    except ImportError:
        ref_check = None
    data, cdata = pm.createData(), pcm.createData()
    out = np.zeros(len(Q), dtype=bool)
    if ref_check is not None:
        for i, q in enumerate(Q):
            out[i] = bool(ref_check(pm, pcm, q, data, cdata, distance_padding=padding))
        return out
    for k in range(len(cdata.collisionRequests)):       # core/utils.py:45-46
        cdata.collisionRequests[k].security_margin = padding
    for i, q in enumerate(Q):                           # core/utils.py:48-51
        pin.computeCollisions(pm, data, pcm, cdata, np.asarray(q, dtype=np.float64), True)
        out[i] = bool(np.any([cr.isCollision() for cr in cdata.collisionResults]))
    return out


def min_distances(pm, pcm, Q):
    """float [N]: min over the active pairs of coal's min_distance (core/utils.py:54-87)."""
    import pinocchio as pin

    data, cdata = pm.createData(), pcm.createData()
    out = np.full(len(Q), np.inf)
    for i, q in enumerate(Q):
        pin.computeDistances(pm, data, pcm, cdata, np.asarray(q, dtype=np.float64))
        out[i] = min((dr.min_distance for dr in cdata.distanceResults), default=np.inf)
    return out


_POOL_MODELS = None


def _pool_init(desc):
    global _POOL_MODELS
    from pyroboplan_b200.model import build_models_from_description
    from pyroboplan_b200.models import panda

    model, cmodel, vmodel = build_models_from_description(desc)
    panda.add_self_collisions(model, cmodel)
    panda.add_object_collisions(model, cmodel, vmodel)
    _POOL_MODELS = to_pinocchio(model, cmodel)


def _pool_work(Q):
    return check_states(_POOL_MODELS[0], _POOL_MODELS[1], Q)


def timed_rate(Q, processes=None):
    """(states/s, verdicts) of the reference's own check on the config-2 Panda, one model copy per worker
    process (the reference is single-threaded; SURVEY.md 8d asks for the all-cores figure too)."""
    import multiprocessing as mp

    from pyroboplan_b200.models.utils import load_description

    processes = processes or len(os.sched_getaffinity(0))
    desc = load_description("panda")
    with mp.get_context("spawn").Pool(processes, initializer=_pool_init, initargs=(desc,)) as pool:
        parts = np.array_split(np.asarray(Q, dtype=np.float64), processes * 4)
        pool.map(_pool_work, [p[:8] for p in parts[:processes]])     # warm every worker
        t0 = time.perf_counter()
        res = pool.map(_pool_work, parts)
        dt = time.perf_counter() - t0
    return len(Q) / dt, np.concatenate(res)
