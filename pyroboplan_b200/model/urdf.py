"""URDF / SRDF / STL ingestion without Pinocchio.

Replaces, for the collision hot path, what the reference obtains from
``pinocchio.buildModelsFromUrdf(urdf, package_dirs=...)``
(/root/reference/src/pyroboplan/models/panda.py:27, models/two_dof.py:25) and
``pinocchio.removeCollisionPairs(model, collision_model, srdf)`` (models/panda.py:79).

Conventions restated: joints are numbered in depth-first order of the
link tree with the children of a link visited in joint-name order (urdfdom stores joints in a
name-sorted map); fixed joints are merged into the parent joint; collision geometries are
named ``"<link>_<i>"`` and numbered in depth-first link order; ``<mimic>`` is ignored.

A URDF is first lowered to a plain ``dict`` *description* (links, joints, inline mesh
vertices).  The bundled example models ship as such descriptions
(``pyroboplan_b200/models/assets/*.json``, produced by ``tools/make_model_assets.py``).
"""

import os
import struct
import xml.etree.ElementTree as ET

import numpy as np

from .robot_model import (
    FRAME_BODY,
    FRAME_FIXED_JOINT,
    FRAME_JOINT,
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


# --------------------------------------------------------------------------- STL
def read_stl(path):
    """Read a binary or ASCII STL. Returns (vertices [n,3] float64 unique, triangles [m,3] int)."""
    with open(path, "rb") as f:
        raw = f.read()
    tris = None
    if len(raw) >= 84:
        (count,) = struct.unpack_from("<I", raw, 80)
        if 84 + 50 * count == len(raw):
            rec = np.frombuffer(raw, dtype=np.uint8, count=50 * count, offset=84).reshape(count, 50)
            tris = rec[:, 12:48].copy().view("<f4").reshape(count, 3, 3).astype(np.float64)
    if tris is None:
        pts = []
        for line in raw.decode("ascii", errors="ignore").splitlines():
            tok = line.split()
            if len(tok) == 4 and tok[0] == "vertex":
                pts.append([float(tok[1]), float(tok[2]), float(tok[3])])
        if len(pts) == 0 or len(pts) % 3:
            raise ValueError(f"{path}: not a valid STL")
        tris = np.array(pts, dtype=np.float64).reshape(-1, 3, 3)
    verts, inverse = np.unique(tris.reshape(-1, 3), axis=0, return_inverse=True)
    return verts, inverse.reshape(-1, 3)


def mesh_signed_volume(verts, tris):
    a, b, c = verts[tris[:, 0]], verts[tris[:, 1]], verts[tris[:, 2]]
    return float(np.einsum("ij,ij->i", a, np.cross(b, c)).sum() / 6.0)


def _resolve_mesh(filename, urdf_dir, package_dirs):
    if filename.startswith("package://"):
        rel = filename[len("package://") :]
        for d in package_dirs:
            cand = os.path.join(d, rel)
            if os.path.exists(cand):
                return cand
        raise FileNotFoundError(f"cannot resolve {filename} in {package_dirs}")
    if filename.startswith("file://"):
        filename = filename[len("file://") :]
    if not os.path.isabs(filename):
        filename = os.path.join(urdf_dir, filename)
    return filename


# --------------------------------------------------------------------------- URDF -> description
def _floats(text, n, default):
    if text is None:
        return list(default)
    vals = [float(v) for v in text.split()]
    if len(vals) != n:
        raise ValueError(f"expected {n} numbers, got {text!r}")
    return vals


def _origin(elem):
    o = elem.find("origin") if elem is not None else None
    if o is None:
        return [0.0, 0.0, 0.0], [0.0, 0.0, 0.0]
    return _floats(o.get("xyz"), 3, [0, 0, 0]), _floats(o.get("rpy"), 3, [0, 0, 0])


def urdf_to_description(urdf_path, package_dirs=None, inline_meshes=True):
    """Parse a URDF file into the neutral description dict (see module docstring)."""
    if package_dirs is None:
        package_dirs = []
    if isinstance(package_dirs, (str, os.PathLike)):
        package_dirs = [package_dirs]
    urdf_dir = os.path.dirname(os.path.abspath(urdf_path))
    root = ET.parse(urdf_path).getroot()
    desc = {"name": root.get("name", "robot"), "links": [], "joints": [], "meshes": {}}
    for link in root.findall("link"):
        cols = []
        for col in link.findall("collision"):
            xyz, rpy = _origin(col)
            g = col.find("geometry")
            if g is None or len(g) == 0:
                continue
            shape = g[0]
            if shape.tag == "mesh":
                fn = shape.get("filename")
                geo = {"type": "mesh", "filename": fn, "scale": _floats(shape.get("scale"), 3, [1, 1, 1])}
                if inline_meshes and fn not in desc["meshes"]:
                    path = _resolve_mesh(fn, urdf_dir, package_dirs)
                    if not path.lower().endswith(".stl"):
                        raise ValueError(f"only STL collision meshes are supported ({path})")
                    verts, tris = read_stl(path)
                    desc["meshes"][fn] = _mesh_record(verts, tris)
            elif shape.tag == "box":
                geo = {"type": "box", "size": _floats(shape.get("size"), 3, None)}
            elif shape.tag == "sphere":
                geo = {"type": "sphere", "radius": float(shape.get("radius"))}
            elif shape.tag == "cylinder":
                geo = {"type": "cylinder", "radius": float(shape.get("radius")), "length": float(shape.get("length"))}
            elif shape.tag == "capsule":
                geo = {"type": "capsule", "radius": float(shape.get("radius")), "length": float(shape.get("length"))}
            else:
                raise ValueError(f"unsupported collision geometry <{shape.tag}>")
            cols.append({"xyz": xyz, "rpy": rpy, "geometry": geo})
        desc["links"].append({"name": link.get("name"), "collisions": cols})
    for joint in root.findall("joint"):
        xyz, rpy = _origin(joint)
        axis_el = joint.find("axis")
        lim = joint.find("limit")
        desc["joints"].append(
            {
                "name": joint.get("name"),
                "type": joint.get("type"),
                "parent": joint.find("parent").get("link"),
                "child": joint.find("child").get("link"),
                "xyz": xyz,
                "rpy": rpy,
                "axis": _floats(axis_el.get("xyz") if axis_el is not None else None, 3, [1, 0, 0]),
                "lower": float(lim.get("lower", 0.0)) if lim is not None else 0.0,
                "upper": float(lim.get("upper", 0.0)) if lim is not None else 0.0,
                "velocity": float(lim.get("velocity", "inf")) if lim is not None else float("inf"),
                "effort": float(lim.get("effort", "inf")) if lim is not None else float("inf"),
            }
        )
    return desc


def _mesh_record(verts, tris):
    from scipy.spatial import ConvexHull

    hull = ConvexHull(verts)
    vol = mesh_signed_volume(verts, tris)
    return {
        "vertices": np.asarray(verts, dtype=np.float64).tolist(),
        "triangles": np.asarray(tris, dtype=np.int64).tolist(),
        "n_triangles": int(len(tris)),
        "volume_ratio": float(abs(vol) / hull.volume) if hull.volume > 0 else None,
        "all_vertices_on_hull": bool(len(hull.vertices) == len(verts)),
    }


# --------------------------------------------------------------------------- description -> models
def _make_shape(geo, meshes):
    t = geo["type"]
    if t == "box":
        return Box(*geo["size"])
    if t == "sphere":
        return Sphere(geo["radius"])
    if t == "cylinder":
        return Cylinder(geo["radius"], geo["length"])
    if t == "capsule":
        return Capsule(geo["radius"], geo["length"])
    if t == "mesh":
        rec = meshes[geo["filename"]]
        pts = np.asarray(rec["vertices"], dtype=np.float64) * np.asarray(geo.get("scale", [1, 1, 1]))
        return Convex(
            pts, source=geo["filename"], volume_ratio=rec.get("volume_ratio"), triangles=rec.get("triangles")
        )
    raise ValueError(f"unknown geometry type {t!r}")


def build_models_from_description(desc):
    """Description dict -> (model, collision_model, visual_model).

    ``visual_model`` is an empty :class:`GeometryModel` (visual meshes are never needed on the
    collision path) kept so reference-style ``add_object_collisions(model, cm, vm)`` calls work.
    """
    links = {l["name"]: l for l in desc["links"]}
    children = {name: [] for name in links}
    child_names = set()
    for j in desc["joints"]:
        children[j["parent"]].append(j)
        child_names.add(j["child"])
    roots = [n for n in links if n not in child_names]
    if len(roots) != 1:
        raise ValueError(f"URDF must have exactly one root link, found {roots}")
    for lst in children.values():
        lst.sort(key=lambda j: j["name"])

    model = Model(desc.get("name", "robot"))
    cmodel = GeometryModel()
    vmodel = GeometryModel()
    model.addFrame(Frame("root_joint", 0, SE3.Identity(), FRAME_FIXED_JOINT, 0))

    # link name -> (joint id, placement of link frame in joint frame, body frame id)
    link_info = {}
    link_order = []

    def visit(link_name, joint_id, placement, parent_frame):
        fid = model.addFrame(Frame(link_name, joint_id, placement, FRAME_BODY, parent_frame))
        link_info[link_name] = (joint_id, placement, fid)
        link_order.append(link_name)
        for j in children[link_name]:
            origin = SE3.from_xyz_rpy(j["xyz"], j["rpy"])
            jt = j["type"]
            if jt == "fixed":
                pl = placement * origin
                jf = model.addFrame(Frame(j["name"], joint_id, pl, FRAME_FIXED_JOINT, fid))
                visit(j["child"], joint_id, pl, jf)
            elif jt in ("revolute", "prismatic"):
                code = JOINT_REVOLUTE if jt == "revolute" else JOINT_PRISMATIC
                jid = model.addJoint(
                    joint_id, code, j["axis"], placement * origin, j["name"],
                    j["lower"], j["upper"], j.get("velocity", np.inf), j.get("effort", np.inf),
                )
                jf = model.addFrame(Frame(j["name"], jid, SE3.Identity(), FRAME_JOINT, fid))
                visit(j["child"], jid, SE3.Identity(), jf)
            else:
                raise NotImplementedError(
                    f"joint {j['name']!r}: type {jt!r} is not supported (fixed/revolute/prismatic only)"
                )

    visit(roots[0], 0, SE3.Identity(), 1)

    meshes = desc.get("meshes", {})
    for link_name in link_order:
        joint_id, placement, fid = link_info[link_name]
        for i, col in enumerate(links[link_name]["collisions"]):
            shape = _make_shape(col["geometry"], meshes)
            pl = placement * SE3.from_xyz_rpy(col["xyz"], col["rpy"])
            cmodel.addGeometryObject(GeometryObject(f"{link_name}_{i}", joint_id, fid, pl, shape))
    return model, cmodel, vmodel


def buildModelsFromUrdf(urdf_path, package_dirs=None):
    """Drop-in for ``pinocchio.buildModelsFromUrdf`` on the collision path (fixed base)."""
    return build_models_from_description(urdf_to_description(urdf_path, package_dirs))


# --------------------------------------------------------------------------- SRDF
def srdf_disabled_pairs(srdf_path):
    """List of (link1, link2) from the ``<disable_collisions>`` entries of an SRDF file."""
    root = ET.parse(srdf_path).getroot()
    return [(e.get("link1"), e.get("link2")) for e in root.findall("disable_collisions")]


def removeCollisionPairs(model, collision_model, srdf, verbose=False):
    """Drop-in for ``pinocchio.removeCollisionPairs``.

    ``srdf`` is an SRDF file path or an iterable of (link1, link2) names.  For every entry, all
    pairs between geometries attached to the two links' BODY frames are removed.
    """
    pairs = srdf_disabled_pairs(srdf) if isinstance(srdf, (str, os.PathLike)) else list(srdf)
    for l1, l2 in pairs:
        f1 = model.getFrameId(l1, FRAME_BODY)
        f2 = model.getFrameId(l2, FRAME_BODY)
        if f1 >= model.nframes or f2 >= model.nframes:
            continue
        for i, gi in enumerate(collision_model.geometryObjects):
            if gi.parentFrame != f1:
                continue
            for k, gk in enumerate(collision_model.geometryObjects):
                if gk.parentFrame == f2 and i != k:
                    collision_model.removeCollisionPair(CollisionPair(i, k))
                    if verbose:
                        print(f"removed pair ({gi.name}, {gk.name})")
