"""ctypes binding of ``libpbp_engine.so`` (C ABI in ``include/pbp_engine.h``) + model flattening.

``CollisionEngine`` owns one device-resident compiled model and exposes the batched
operations of the hot path.  Inputs/outputs are either CUDA ``torch`` tensors (zero-copy, the
call is asynchronous on torch's current stream) or host ``numpy`` arrays (the ``*_host`` entry
points copy in, run and copy out).  PyTorch is used only for device memory and streams.

There is no CPU fallback: a missing library or a missing CUDA device raises ``RuntimeError``.
"""

import ctypes
import os

import numpy as np

from .model.shapes import Box, Capsule, Convex, Cylinder, Sphere
from .model.surface import N_INNER_BALLS, inner_balls, surface_tables

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpbp_engine.so")

PBP_ABI_VERSION = 6

# symbols declared in include/pbp_engine.h (tests check the library exports every one)
EXPORTED_SYMBOLS = [
    "pbp_last_error",
    "pbp_abi_version",
    "pbp_engine_create",
    "pbp_engine_destroy",
    "pbp_engine_device",
    "pbp_get_stats",
    "pbp_fk",
    "pbp_check_states",
    "pbp_check_states_host",
    "pbp_check_states_host_f64",
    "pbp_check_edges",
    "pbp_check_edges_host",
    "pbp_discretize_counts",
    "pbp_discretize_fill",
    "pbp_check_paths",
    "pbp_sample_free_states",
    "pbp_set_profiling",
    "pbp_get_profile",
    "pbp_measure_fp32_peak",
    "pbp_ik_iterate",
    "pbp_knn",
    "pbp_knn_blocks",
    "pbp_compute_distances",
    "pbp_collision_nullspace",
    "pbp_jit_broadphase_source",
    "pbp_jit_compile_check",
]


class _ModelDesc(ctypes.Structure):
    _fields_ = [
        ("nq", ctypes.c_int32),
        ("njoints", ctypes.c_int32),
        ("ngeoms", ctypes.c_int32),
        ("npairs", ctypes.c_int32),
        ("nverts", ctypes.c_int32),
        ("nframes", ctypes.c_int32),
        ("joint_parent", ctypes.c_void_p),
        ("joint_type", ctypes.c_void_p),
        ("joint_idx_q", ctypes.c_void_p),
        ("joint_placement", ctypes.c_void_p),
        ("joint_axis", ctypes.c_void_p),
        ("q_lower", ctypes.c_void_p),
        ("q_upper", ctypes.c_void_p),
        ("geom_parent", ctypes.c_void_p),
        ("geom_shape", ctypes.c_void_p),
        ("geom_placement", ctypes.c_void_p),
        ("geom_params", ctypes.c_void_p),
        ("geom_vert_off", ctypes.c_void_p),
        ("geom_vert_cnt", ctypes.c_void_p),
        ("geom_outer", ctypes.c_void_p),
        ("geom_inner", ctypes.c_void_p),
        ("verts", ctypes.c_void_p),
        ("pairs", ctypes.c_void_p),
        ("frame_parent", ctypes.c_void_p),
        ("frame_placement", ctypes.c_void_p),
        ("joint_placement_f64", ctypes.c_void_p),
        ("joint_axis_f64", ctypes.c_void_p),
        ("frame_placement_f64", ctypes.c_void_p),
        ("q_lower_f64", ctypes.c_void_p),
        ("q_upper_f64", ctypes.c_void_p),
        ("geom_capsule", ctypes.c_void_p),
        ("planes", ctypes.c_void_p),
        ("geom_plane_off", ctypes.c_void_p),
        ("geom_plane_cnt", ctypes.c_void_p),
        ("geom_surface_tol", ctypes.c_void_p),
        ("pair_enclose", ctypes.c_void_p),
        ("nplanes", ctypes.c_int32),
        ("ncore", ctypes.c_int32),
        ("nmverts", ctypes.c_int32),
        ("ntris", ctypes.c_int32),
        ("geom_plane_hull_cnt", ctypes.c_void_p),
        ("core_verts", ctypes.c_void_p),
        ("geom_core_off", ctypes.c_void_p),
        ("geom_core_cnt", ctypes.c_void_p),
        ("geom_core_eps", ctypes.c_void_p),
        ("mesh_verts", ctypes.c_void_p),
        ("geom_mvert_off", ctypes.c_void_p),
        ("tris", ctypes.c_void_p),
        ("geom_tri_off", ctypes.c_void_p),
        ("geom_tri_cnt", ctypes.c_void_p),
        ("geom_inner_balls", ctypes.c_void_p),
    ]


class _Stats(ctypes.Structure):
    _fields_ = [
        ("kernel_launches", ctypes.c_int64),
        ("states", ctypes.c_int64),
        ("narrowphase_items", ctypes.c_int64),
        ("chunks", ctypes.c_int64),
    ]


class _IkOptions(ctypes.Structure):
    _fields_ = [
        ("max_iters", ctypes.c_int32),
        ("max_translation_error", ctypes.c_double),
        ("max_rotation_error", ctypes.c_double),
        ("damping", ctypes.c_double),
        ("min_step_size", ctypes.c_double),
        ("max_step_size", ctypes.c_double),
    ]


class _Profile(ctypes.Structure):
    _fields_ = [
        ("broadphase_ms", ctypes.c_double),
        ("item_setup_ms", ctypes.c_double),
        ("narrowphase_ms", ctypes.c_double),
        ("rounds", ctypes.c_int64),
        ("narrowphase_items", ctypes.c_int64),
        ("surface_refine_ms", ctypes.c_double),
    ]


_lib = None


def load_library():
    """Load ``libpbp_engine.so``; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    lib_path = os.environ.get("PBP_ENGINE_LIB", LIB_PATH)   # developer override: A/B builds of the same ABI
    if not os.path.exists(lib_path):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). pyroboplan_b200 has no CPU fallback."
        )
    lib = ctypes.CDLL(lib_path)
    vp, i64, i32, f32, f64 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_float, ctypes.c_double
    lib.pbp_last_error.restype = ctypes.c_char_p
    lib.pbp_abi_version.restype = ctypes.c_int
    lib.pbp_engine_create.argtypes = [ctypes.POINTER(_ModelDesc), i32, i64, ctypes.POINTER(vp)]
    lib.pbp_engine_destroy.argtypes = [vp]
    lib.pbp_engine_destroy.restype = None
    lib.pbp_engine_device.argtypes = [vp]
    lib.pbp_get_stats.argtypes = [vp, ctypes.POINTER(_Stats), i32]
    lib.pbp_fk.argtypes = [vp, vp, i64, vp, vp, vp, vp]
    lib.pbp_check_states.argtypes = [vp, vp, i64, f32, vp, vp, vp, vp]
    lib.pbp_check_states_host.argtypes = [vp, vp, i64, f32, vp, vp, vp]
    lib.pbp_check_states_host_f64.argtypes = [vp, vp, i64, f32, vp, vp, vp]
    lib.pbp_check_edges.argtypes = [vp, vp, vp, i64, f64, f32, vp, vp, vp]
    lib.pbp_check_edges_host.argtypes = [vp, vp, vp, i64, f64, f32, vp, vp]
    lib.pbp_discretize_counts.argtypes = [vp, vp, vp, i64, f64, vp, vp]
    lib.pbp_discretize_fill.argtypes = [vp, vp, vp, i64, vp, vp, vp, vp]
    lib.pbp_check_paths.argtypes = [vp, vp, vp, i64, i64, f32, vp, vp]
    lib.pbp_sample_free_states.argtypes = [vp, i64, ctypes.c_uint64, f32, f32, i32, vp, ctypes.POINTER(i64), ctypes.POINTER(i64), vp]
    lib.pbp_set_profiling.argtypes = [vp, i32]
    lib.pbp_get_profile.argtypes = [vp, ctypes.POINTER(_Profile), i32]
    lib.pbp_measure_fp32_peak.argtypes = [i32, ctypes.POINTER(f64)]
    lib.pbp_ik_iterate.argtypes = [vp, i32, vp, vp, i64, ctypes.POINTER(_IkOptions), ctypes.c_uint64, vp, vp, vp, vp, vp, vp]
    lib.pbp_knn.argtypes = [vp, vp, i64, i32, f64, i32, vp, vp, vp]
    lib.pbp_knn_blocks.argtypes = [vp, vp, i64, i32, f64, i32, i64, i64, vp, vp, vp]
    lib.pbp_compute_distances.argtypes = [vp, vp, i64, f32, vp, vp, vp, vp, vp]
    lib.pbp_collision_nullspace.argtypes = [vp, vp, i64, f32, f32, vp, vp]
    if lib.pbp_abi_version() != PBP_ABI_VERSION:
        raise RuntimeError("libpbp_engine.so ABI version mismatch; rebuild it")
    _lib = lib
    return lib


def _check(rc):
    if rc != 0:
        msg = _lib.pbp_last_error().decode("utf-8", "replace") if _lib else ""
        raise RuntimeError(f"pbp_engine error {rc}: {msg}")


def _se3_row(T):
    return np.concatenate([np.asarray(T.rotation, dtype=np.float64).reshape(9), np.asarray(T.translation, dtype=np.float64)])


def _up32(x):
    """Smallest float32 >= x (x >= 0)."""
    y = np.float32(x)
    return y if float(y) >= x else np.nextafter(y, np.float32(np.inf))


def _shape_volume(s):
    if isinstance(s, Sphere):
        return 4.0 / 3.0 * np.pi * s.radius ** 3
    if isinstance(s, Box):
        return float(8.0 * np.prod(s.halfSide))
    if isinstance(s, Cylinder):
        return np.pi * s.radius ** 2 * 2.0 * s.halfLength
    if isinstance(s, Capsule):
        return np.pi * s.radius ** 2 * 2.0 * s.halfLength + 4.0 / 3.0 * np.pi * s.radius ** 3
    return float(s.volume)


def _could_enclose(outer, inner):
    """Necessary conditions for ``inner`` to fit inside ``outer`` in SOME relative placement: inclusion
    is monotone in the inscribed radius, the enclosing radius and the volume.  (Conservative: a
    pair flagged needlessly only costs its hull hits an enclosure test.)"""
    tol = 1.002   # the enclosing balls of hulls are computed iteratively
    return (inner.inner_sphere()[1] <= tol * outer.inner_sphere()[1] and inner.bounding_sphere()[1] <= tol * outer.bounding_sphere()[1]
            and _shape_volume(inner) <= tol * _shape_volume(outer))


def mesh_dent_depth(s):
    """How far the surface of a mesh lies below the surface of its convex hull (metres): the deepest
    sample of the mesh triangles' edges and centroids under the hull's face planes."""
    v = np.asarray(s.mesh_vertices, dtype=np.float64)[np.asarray(s.mesh_triangles)]          # [T, 3, 3]
    t = np.linspace(0.0, 1.0, 9)[1:-1]
    samples = [v.mean(axis=1)]
    for i, j in ((0, 1), (1, 2), (2, 0)):
        samples.append((v[:, i, None, :] * (1 - t)[None, :, None] + v[:, j, None, :] * t[None, :, None]).reshape(-1, 3))
    x = np.concatenate(samples)
    depth = (-(x @ s.planes[:, :3].T + s.planes[:, 3])).min(axis=1)                        # distance to the nearest face plane
    return float(max(depth.max(), 0.0))


def flatten_model(model, collision_model, mesh_semantics="surface"):
    """Host containers -> dict of contiguous float32 / int32 arrays matching ``pbp_model_desc``.

    ``geom_outer`` / ``geom_inner`` share one centre per geometry (the enclosing sphere's); radii
    are rounded outwards / inwards so the float32 broadphase stays conservative.

    ``mesh_semantics``: "surface" (default) treats geometries that came from a mesh file (they carry
    ``mesh_triangles``) as the reference does -- Pinocchio's BVHModel is the SURFACE, so a shape wholly
    inside the mesh does not collide with it --; "hull" treats them as convex solids.
    """
    geoms = collision_model.geometryObjects
    ng = len(geoms)
    a = {
        "joint_parent": np.array(model.parents, dtype=np.int32),
        "joint_type": np.array([j.type for j in model.joints], dtype=np.int32),
        "joint_idx_q": np.array([j.idx_q for j in model.joints], dtype=np.int32),
        "joint_placement": np.array([_se3_row(p) for p in model.jointPlacements], dtype=np.float32),
        "joint_axis": np.array([j.axis for j in model.joints], dtype=np.float32),
        "q_lower": np.asarray(model.lowerPositionLimit, dtype=np.float32),
        "q_upper": np.asarray(model.upperPositionLimit, dtype=np.float32),
        "geom_parent": np.array([g.parentJoint for g in geoms], dtype=np.int32).reshape(ng),
        "geom_shape": np.zeros(ng, dtype=np.int32),
        "geom_placement": np.zeros((ng, 12), dtype=np.float32),
        "geom_params": np.zeros((ng, 4), dtype=np.float32),
        "geom_vert_off": np.zeros(ng, dtype=np.int32),
        "geom_vert_cnt": np.zeros(ng, dtype=np.int32),
        "geom_outer": np.zeros((ng, 4), dtype=np.float32),
        "geom_inner": np.zeros((ng, 4), dtype=np.float32),
        "geom_capsule": np.zeros((ng, 8), dtype=np.float32),
        "geom_inner_balls": np.zeros((ng, N_INNER_BALLS, 4), dtype=np.float32),
    }
    verts = []
    nv = 0
    for i, g in enumerate(geoms):
        s = g.geometry
        a["geom_placement"][i] = _se3_row(g.placement)
        a["geom_shape"][i] = s.shape_type
        if isinstance(s, Sphere):
            a["geom_params"][i, 0] = s.radius
        elif isinstance(s, Box):
            a["geom_params"][i, :3] = s.halfSide
        elif isinstance(s, (Cylinder, Capsule)):
            a["geom_params"][i, :2] = [s.radius, s.halfLength]
        elif isinstance(s, Convex):
            pts = np.asarray(s.points, dtype=np.float32)
            a["geom_vert_off"][i] = nv
            a["geom_vert_cnt"][i] = len(pts)
            verts.append(pts)
            nv += len(pts)
        else:
            raise TypeError(f"unsupported collision geometry {type(s).__name__}")
        # One proxy centre per geometry, two radii: the enclosing radius (rounded outwards) and the
        # radius of the largest ball around the SAME centre that stays inside the shape (rounded
        # inwards), so the float32 broadphase stays conservative.
        oc, orad = s.bounding_sphere()
        oc32 = np.asarray(oc, dtype=np.float32)
        c64 = oc32.astype(np.float64)
        if isinstance(s, Convex):
            p64 = np.asarray(s.points, dtype=np.float32).astype(np.float64)
            orad = float(np.sqrt(((p64 - c64) ** 2).sum(axis=1).max()))
            irad = float((-(s.planes[:, :3] @ c64 + s.planes[:, 3])).min())
        elif isinstance(s, Box):
            irad = float(np.min(s.halfSide))
        elif isinstance(s, Cylinder):
            irad = min(s.radius, s.halfLength)
        else:  # Sphere, Capsule
            irad = s.radius
        if isinstance(s, Sphere):
            o32, i32r = np.float32(s.radius), np.float32(s.radius)  # exact pair tests use the true radius
        else:
            o32 = _up32(orad * (1.0 + 1e-6) + 1e-7)
            i32r = np.float32(max(irad - 2e-6, 0.0))
        a["geom_outer"][i] = [*oc32, o32]
        a["geom_inner"][i] = [*oc32, i32r]
        if isinstance(s, Convex):
            a["geom_capsule"][i] = enclosing_capsule(np.asarray(s.points, dtype=np.float32).astype(np.float64), float(o32))
            # balls inside the solid (a surface geometry gets balls inside its inner CORE below)
            a["geom_inner_balls"][i] = _cached_balls(s, "_pbp_inner_balls_hull", np.concatenate([s.planes[:, :3], -s.planes[:, 3:4]], axis=1),
                                                      np.asarray(s.points, dtype=np.float32).astype(np.float64))
        elif isinstance(s, Sphere):
            a["geom_inner_balls"][i, 0] = [0.0, 0.0, 0.0, float(np.float32(s.radius)) - 2e-6]
        elif isinstance(s, (Cylinder, Capsule)):
            r = float(np.float32(s.radius)) - 2e-6
            hl = s.halfLength if isinstance(s, Capsule) else max(s.halfLength - s.radius, 0.0)
            for k, z in enumerate(np.linspace(-hl, hl, N_INNER_BALLS) if hl > 0 else [0.0]):
                a["geom_inner_balls"][i, k] = [0.0, 0.0, z, min(r, s.halfLength) if isinstance(s, Cylinder) else r]
    a["verts"] = np.concatenate(verts).astype(np.float32) if verts else np.zeros((1, 3), dtype=np.float32)
    pairs = np.array([[p.first, p.second] for p in collision_model.collisionPairs], dtype=np.int32).reshape(-1, 2)
    a["pairs"] = pairs if len(pairs) else np.zeros((1, 2), dtype=np.int32)
    # surface semantics of mesh geometries (model/surface.py): planes of the exact inner core (hull faces
    # first, largest first), the simplified core's vertices, mesh vertices and triangles, pairs where one
    # shape could lie wholly inside the other's surface
    if mesh_semantics not in ("surface", "hull"):
        raise ValueError("mesh_semantics must be 'surface' or 'hull'")
    surface = [mesh_semantics == "surface" and isinstance(g.geometry, Convex) and g.geometry.mesh_triangles is not None for g in geoms]
    zi = lambda: np.zeros(ng, np.int32)
    planes, plane_off, plane_cnt, plane_hull, tol = [], zi(), zi(), zi(), np.zeros(ng, np.float32)
    cores, core_off, core_cnt, core_eps = [], zi(), zi(), np.zeros(ng, np.float32)
    mverts, mvert_off, tris, tri_off, tri_cnt = [], zi(), [], zi(), zi()
    for i, g in enumerate(geoms):
        plane_off[i] = sum(len(x) for x in planes)
        core_off[i] = sum(len(x) for x in cores)
        mvert_off[i] = sum(len(x) for x in mverts)
        tri_off[i] = sum(len(x) for x in tris)
        if not surface[i]:
            continue
        st = surface_tables(g.geometry)
        nh = st["n_hull_planes"]
        order = np.concatenate([np.argsort(-st["plane_areas"][:nh], kind="stable"), np.arange(nh, len(st["core_planes"]))])
        planes.append(st["core_planes"][order])
        plane_cnt[i], plane_hull[i] = len(order), nh
        tol[i] = _up32(st["surface_tol"])
        mverts.append(st["mesh_verts"])
        tris.append(st["tris"])
        tri_cnt[i] = len(st["tris"])
        if st["has_core"]:
            cores.append(st["core"])
            core_cnt[i] = len(st["core"])
            core_eps[i] = _up32(st["core_eps"])
            # the broadphase's "certainly colliding" ball must lie inside the CORE (and so inside the mesh)
            kp = st["core_hull_planes"]
            c64 = a["geom_inner"][i, :3].astype(np.float64)
            a["geom_inner"][i, 3] = np.float32(max(float((kp[:, 3] - kp[:, :3] @ c64).min()) - 2e-6, 0.0))
            a["geom_inner_balls"][i] = _cached_balls(g.geometry, "_pbp_inner_balls_core", kp, st["core"].astype(np.float64))
        else:
            a["geom_inner"][i, 3] = 0.0
            a["geom_inner_balls"][i] = 0.0    # no core: nothing certifies a hit
    enclose = np.zeros(max(len(pairs), 1), np.uint8)
    for k, (ga, gb) in enumerate(pairs):
        if surface[ga] and _could_enclose(geoms[ga].geometry, geoms[gb].geometry):
            enclose[k] |= 1
        if surface[gb] and _could_enclose(geoms[gb].geometry, geoms[ga].geometry):
            enclose[k] |= 2
    cat = lambda parts, shape, dt: np.concatenate(parts).astype(dt) if parts else np.zeros(shape, dtype=dt)
    a["planes"] = cat(planes, (1, 4), np.float32)
    a["geom_plane_off"], a["geom_plane_cnt"], a["geom_plane_hull_cnt"], a["geom_surface_tol"], a["pair_enclose"] = plane_off, plane_cnt, plane_hull, tol, enclose
    a["core_verts"], a["geom_core_off"], a["geom_core_cnt"], a["geom_core_eps"] = cat(cores, (1, 3), np.float32), core_off, core_cnt, core_eps
    a["mesh_verts"], a["geom_mvert_off"] = cat(mverts, (1, 3), np.float32), mvert_off
    a["tris"], a["geom_tri_off"], a["geom_tri_cnt"] = cat(tris, (1, 3), np.int32), tri_off, tri_cnt
    a["frame_parent"] = np.array([f.parentJoint for f in model.frames], dtype=np.int32)
    a["frame_placement"] = np.array([_se3_row(f.placement) for f in model.frames], dtype=np.float32)
    # float64 kinematic tables for the FP64 differential-IK kernel
    a["joint_placement_f64"] = np.array([_se3_row(p) for p in model.jointPlacements], dtype=np.float64)
    a["joint_axis_f64"] = np.array([j.axis for j in model.joints], dtype=np.float64)
    a["frame_placement_f64"] = np.array([_se3_row(f.placement) for f in model.frames], dtype=np.float64)
    a["q_lower_f64"] = np.asarray(model.lowerPositionLimit, dtype=np.float64)
    a["q_upper_f64"] = np.asarray(model.upperPositionLimit, dtype=np.float64)
    sizes = dict(nq=model.nq, njoints=model.njoints, ngeoms=ng, npairs=len(pairs), nverts=nv, nframes=len(model.frames),
                 nplanes=int(plane_cnt.sum()), ncore=int(core_cnt.sum()), nmverts=int(sum(len(x) for x in mverts)), ntris=int(tri_cnt.sum()))
    return {k: np.ascontiguousarray(v) for k, v in a.items()}, sizes


_BALL_CACHE = {}


def _cached_balls(shape, attr, planes, points):
    """inner_balls() per shape, cached on the shape and by content (models are loaded many times per process)."""
    import hashlib

    cached = getattr(shape, attr, None)
    if cached is None:
        key = hashlib.sha1(np.ascontiguousarray(planes).tobytes() + np.ascontiguousarray(points).tobytes()).hexdigest()
        cached = _BALL_CACHE.get(key)
        if cached is None:
            cached = _BALL_CACHE[key] = inner_balls(planes, points)
        setattr(shape, attr, cached)
    return cached


def enclosing_capsule(points, r_sphere, useful_ratio=None):
    """A capsule (segment + radius) that contains every hull vertex, hence the hull: a second, much
    tighter "certainly free" proxy for elongated links (Panda link5: radius 0.094 against a 0.173
    enclosing sphere).  Axis = first principal axis; the radius is the largest distance of a vertex
    from the axis, the segment is as short as that radius allows.  Returns
    [p0 (3), p1 (3), radius, useful flag] in float32, the radius re-measured against the
    float32-rounded segment and rounded up; the flag is 0 when the capsule is not clearly tighter
    than the sphere (radius >= useful_ratio * r_sphere): every capsule costs the specialised broadphase a
    point-segment (or segment-segment) test per pair and registers for its end points."""
    if useful_ratio is None:
        # measured on config 2 (1 Mi Panda states): 0 (no capsules) 348 us, 0.55 (link5, hand) 320 us,
        # 0.62 (+ fingers: two segment-segment tests, 92 registers) 337 us, 0.7 (+ links 1-4) 386 us
        useful_ratio = float(os.environ.get("PBP_CAPSULE_RATIO", 0.55))
    c = points.mean(axis=0)
    _, _, vt = np.linalg.svd(points - c)
    ax = vt[0]
    t = (points - c) @ ax
    rho = np.linalg.norm((points - c) - np.outer(t, ax), axis=1)
    r = float(rho.max())
    reach = np.sqrt(np.maximum(r * r - rho * rho, 0.0))
    lo, hi = float((t + reach).min()), float((t - reach).max())
    if lo > hi:
        lo = hi = 0.5 * (lo + hi)
    p0 = (c + lo * ax).astype(np.float32)
    p1 = (c + hi * ax).astype(np.float32)
    u = p1.astype(np.float64) - p0.astype(np.float64)
    uu = float(u @ u)
    s = np.clip(((points - p0.astype(np.float64)) @ u) / uu, 0.0, 1.0) if uu > 0 else np.zeros(len(points))
    dist = np.linalg.norm(points - (p0.astype(np.float64) + np.outer(s, u)), axis=1)
    r32 = _up32(float(dist.max()) * (1.0 + 1e-6) + 1e-7)
    return np.array([*p0, *p1, r32, 1.0 if float(r32) < useful_ratio * r_sphere else 0.0], dtype=np.float32)


def _torch():
    import torch

    return torch


class CollisionEngine:
    """Device-resident compiled (model, collision_model) + batched hot-path operations."""

    def __init__(self, model, collision_model, device=None, max_chunk_states=0, mesh_semantics="surface"):
        self.lib = load_library()
        torch = _torch()
        if not torch.cuda.is_available():
            raise RuntimeError("pyroboplan_b200 needs a CUDA device (B200); there is no CPU fallback")
        if device is None:
            device = torch.cuda.current_device()
        if not isinstance(device, int):     # torch.device, "cuda:1", "cuda"
            d = torch.device(device)
            if d.type != "cuda":
                raise ValueError(f"CollisionEngine needs a CUDA device, got {device!r}")
            device = d.index if d.index is not None else torch.cuda.current_device()
        self.device = int(device)
        self.torch_device = torch.device("cuda", self.device)
        from .model.pinocchio_adapter import from_pinocchio, is_native_model

        if not is_native_model(collision_model):      # Pinocchio objects, as the reference's callers hold them
            collision_model = from_pinocchio(model, collision_model)[1]
        if not is_native_model(model):
            model = from_pinocchio(model, None)[0]
        arrays, sizes = flatten_model(model, collision_model, mesh_semantics)
        self._arrays = arrays
        self.host_model, self.host_collision_model = model, collision_model   # the containers the tables were compiled from
        self.mesh_semantics = mesh_semantics
        self.nq, self.njoints, self.ngeoms = sizes["nq"], sizes["njoints"], sizes["ngeoms"]
        self.npairs, self.nframes = sizes["npairs"], sizes["nframes"]
        if sizes["ntris"] == 0:   # no surface geometry: the optional tables stay NULL
            arrays = {k: v for k, v in arrays.items() if k not in (
                "planes", "geom_plane_off", "geom_plane_cnt", "geom_plane_hull_cnt", "geom_surface_tol", "pair_enclose", "core_verts", "geom_core_off",
                "geom_core_cnt", "geom_core_eps", "mesh_verts", "geom_mvert_off", "tris", "geom_tri_off", "geom_tri_cnt")}
        desc = _ModelDesc()
        for k, v in sizes.items():
            setattr(desc, k, v)
        for k, v in arrays.items():
            setattr(desc, k, v.ctypes.data)
        handle = ctypes.c_void_p()
        _check(self.lib.pbp_engine_create(ctypes.byref(desc), self.device, int(max_chunk_states), ctypes.byref(handle)))
        self._h = handle

    def close(self):
        if getattr(self, "_h", None):
            self.lib.pbp_engine_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ helpers
    def _stream(self):
        return ctypes.c_void_p(_torch().cuda.current_stream(self.torch_device).cuda_stream)

    def _dev_q(self, Q):
        torch = _torch()
        if not (isinstance(Q, torch.Tensor) and Q.is_cuda):
            raise TypeError("expected a CUDA torch tensor")
        if Q.device != self.torch_device:
            raise ValueError(f"tensor on {Q.device}, engine on {self.torch_device}")
        Q = Q.to(torch.float32).contiguous().reshape(-1, self.nq)
        return Q

    @staticmethod
    def _host_q(Q, nq):
        return np.ascontiguousarray(np.asarray(Q, dtype=np.float32).reshape(-1, nq))

    def _is_dev(self, x):
        torch = _torch()
        return isinstance(x, torch.Tensor) and x.is_cuda

    def stats(self, reset=False):
        s = _Stats()
        _check(self.lib.pbp_get_stats(self._h, ctypes.byref(s), int(reset)))
        return {k: int(getattr(s, k)) for k, _ in s._fields_}

    def set_profiling(self, enabled=True):
        _check(self.lib.pbp_set_profiling(self._h, int(enabled)))

    def profile(self, reset=False):
        """Per-kernel CUDA-event timings (ms) and narrowphase item count since the last reset."""
        p = _Profile()
        _check(self.lib.pbp_get_profile(self._h, ctypes.byref(p), int(reset)))
        return {k: getattr(p, k) for k, _ in p._fields_}

    # ------------------------------------------------------------------ FK
    def fk(self, Q, joints=True, geoms=True, frames=True):
        """Placements [N, n*, 12] (R row-major, then t) for joints / geometries / frames."""
        torch = _torch()
        host = not self._is_dev(Q)
        Qd = torch.as_tensor(self._host_q(Q, self.nq), device=self.torch_device) if host else self._dev_q(Q)
        n = Qd.shape[0]
        mk = lambda cnt, want: torch.empty((n, cnt, 12), dtype=torch.float32, device=self.torch_device) if want else None
        oMi, oMg, oMf = mk(self.njoints, joints), mk(self.ngeoms, geoms), mk(self.nframes, frames)
        ptr = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
        _check(self.lib.pbp_fk(self._h, ptr(Qd), n, ptr(oMi), ptr(oMg), ptr(oMf), self._stream()))
        if host:
            conv = lambda t: t.cpu().numpy() if t is not None else None
            return conv(oMi), conv(oMg), conv(oMf)
        return oMi, oMg, oMf

    # ------------------------------------------------------------------ states
    def check_states(self, Q, padding=0.0, return_distance=False, return_pair=False, out=None):
        """Collision verdict per configuration.

        CUDA tensor in -> CUDA tensors out (bool [N], optionally float32 min distance [N] and
        int32 first colliding pair [N]); numpy / list in -> numpy out.
        ``out``: CUDA bool / uint8 tensor [N] that receives the verdicts (device path).  A caller that passes
        the same input and output buffers again and again gets the whole round replayed as one CUDA graph.
        """
        torch = _torch()
        if self._is_dev(Q):
            Qd = self._dev_q(Q)
            n = Qd.shape[0]
            if out is not None:
                if not (self._is_dev(out) and out.device == self.torch_device and out.numel() == n and out.is_contiguous()
                        and out.dtype in (torch.bool, torch.uint8)):
                    raise ValueError("out must be a contiguous CUDA bool / uint8 tensor with one element per state")
                hit = out
            else:
                hit = torch.empty(n, dtype=torch.bool, device=self.torch_device)  # kernels write 0/1 bytes
            dist = torch.empty(n, dtype=torch.float32, device=self.torch_device) if return_distance else None
            pair = torch.empty(n, dtype=torch.int32, device=self.torch_device) if return_pair else None
            ptr = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
            _check(self.lib.pbp_check_states(self._h, ptr(Qd), n, float(padding), ptr(hit), ptr(dist), ptr(pair), self._stream()))
            res = [hit]
        else:
            # float64 input (what the reference's callers hold) goes to the engine as it is: the library narrows
            # it slice by slice on a few host threads, overlapped with the copies and the kernels
            f64 = isinstance(Q, np.ndarray) and Q.dtype == np.float64 and Q.size >= 65536 * self.nq
            Qh = np.ascontiguousarray(Q).reshape(-1, self.nq) if f64 else self._host_q(Q, self.nq)
            n = Qh.shape[0]
            # large results land in page-locked memory (torch's caching pinned allocator): the
            # device -> host copies go straight there instead of through the engine's staging buffer
            big = n >= 65536
            hit = torch.empty(n, dtype=torch.uint8, pin_memory=True).numpy() if big else np.zeros(n, dtype=np.uint8)
            dist = (torch.empty(n, dtype=torch.float32, pin_memory=True).numpy() if big else np.zeros(n, dtype=np.float32)) if return_distance else None
            pair = (torch.empty(n, dtype=torch.int32, pin_memory=True).numpy() if big else np.zeros(n, dtype=np.int32)) if return_pair else None
            ptr = lambda t: ctypes.c_void_p(t.ctypes.data) if t is not None else None
            entry = self.lib.pbp_check_states_host_f64 if f64 else self.lib.pbp_check_states_host
            _check(entry(self._h, ptr(Qh), n, float(padding), ptr(hit), ptr(dist), ptr(pair)))
            res = [hit.view(np.bool_)]
        if return_distance:
            res.append(dist)
        if return_pair:
            res.append(pair)
        return res[0] if len(res) == 1 else tuple(res)

    # ------------------------------------------------------------------ edges
    def check_edges(self, Q1, Q2, max_step_size, padding=0.0, return_first_bad=False):
        """hit[E] = True iff the discretised segment Q1[e] -> Q2[e] has a colliding sample."""
        torch = _torch()
        if self._is_dev(Q1):
            A, B = self._dev_q(Q1), self._dev_q(Q2)
            n = A.shape[0]
            hit = torch.empty(n, dtype=torch.bool, device=self.torch_device)
            fb = torch.empty(n, dtype=torch.int32, device=self.torch_device) if return_first_bad else None
            ptr = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
            _check(self.lib.pbp_check_edges(self._h, ptr(A), ptr(B), n, float(max_step_size), float(padding), ptr(hit), ptr(fb), self._stream()))
            out = hit
        else:
            A, B = self._host_q(Q1, self.nq), self._host_q(Q2, self.nq)
            n = A.shape[0]
            hit = np.zeros(n, dtype=np.uint8)
            fb = np.zeros(n, dtype=np.int32) if return_first_bad else None
            ptr = lambda t: ctypes.c_void_p(t.ctypes.data) if t is not None else None
            _check(self.lib.pbp_check_edges_host(self._h, ptr(A), ptr(B), n, float(max_step_size), float(padding), ptr(hit), ptr(fb)))
            out = hit.view(np.bool_)
        return (out, fb) if return_first_bad else out

    def check_paths(self, Q, offsets, padding=0.0):
        """Ragged batch of pre-discretised paths: samples Q[offsets[p]:offsets[p+1]] form path p."""
        torch = _torch()
        host = not self._is_dev(Q)
        Qd = torch.as_tensor(self._host_q(Q, self.nq), device=self.torch_device) if host else self._dev_q(Q)
        off = torch.as_tensor(np.asarray(offsets, dtype=np.int64) if not isinstance(offsets, torch.Tensor) else offsets,
                              device=self.torch_device).to(torch.int64).contiguous()
        n_paths = off.numel() - 1
        hit = torch.empty(max(n_paths, 0), dtype=torch.bool, device=self.torch_device)
        if n_paths > 0:
            _check(self.lib.pbp_check_paths(self._h, ctypes.c_void_p(Qd.data_ptr()), ctypes.c_void_p(off.data_ptr()), n_paths,
                                            Qd.shape[0], float(padding), ctypes.c_void_p(hit.data_ptr()), self._stream()))
        return hit.cpu().numpy() if host else hit

    def discretize(self, Q1, Q2, max_step_size):
        """(counts [E] int32, offsets [E+1] int64, states [sum(counts), nq]) on the input's side."""
        torch = _torch()
        host = not self._is_dev(Q1)
        A = torch.as_tensor(self._host_q(Q1, self.nq), device=self.torch_device) if host else self._dev_q(Q1)
        B = torch.as_tensor(self._host_q(Q2, self.nq), device=self.torch_device) if host else self._dev_q(Q2)
        n = A.shape[0]
        counts = torch.empty(n, dtype=torch.int32, device=self.torch_device)
        ptr = lambda t: ctypes.c_void_p(t.data_ptr())
        _check(self.lib.pbp_discretize_counts(self._h, ptr(A), ptr(B), n, float(max_step_size), ptr(counts), self._stream()))
        offsets = torch.zeros(n + 1, dtype=torch.int64, device=self.torch_device)
        torch.cumsum(counts.to(torch.int64), dim=0, out=offsets[1:])
        total = int(offsets[-1].item()) if n else 0
        states = torch.empty((total, self.nq), dtype=torch.float32, device=self.torch_device)
        if total:
            _check(self.lib.pbp_discretize_fill(self._h, ptr(A), ptr(B), n, ptr(counts), ptr(offsets), ptr(states), self._stream()))
        if host:
            return counts.cpu().numpy(), offsets.cpu().numpy(), states.cpu().numpy()
        return counts, offsets, states

    # ------------------------------------------------------------------ differential IK
    def ik_iterate(self, frame_id, targets, Q, e0=None, max_iters=200, max_translation_error=1e-3, max_rotation_error=1e-3,
                   damping=1e-3, min_step_size=0.1, max_step_size=0.5, active_mask=None, joint_weights=None, nullspace=None):
        """One try of batched DLS differential IK (device tensors; see pbp_ik_iterate in the header).

        targets [N, 12] (R row-major, then t), Q [N, nq] initial states; both are used as float64.  Returns
        (Q_out [N, nq], status [N] int32 (bit 0 converged, bit 1 within limits), iters [N], e0 [N]).
        ``nullspace`` [N, nq]: null-space term held constant over the iterations of this call.
        """
        torch = _torch()
        f64 = torch.float64
        Qd = torch.as_tensor(Q, device=self.torch_device).to(f64).reshape(-1, self.nq).contiguous().clone()
        n = Qd.shape[0]
        Td = torch.as_tensor(targets, device=self.torch_device).to(f64).contiguous().reshape(n, 12)
        e0 = torch.zeros(n, dtype=f64, device=self.torch_device) if e0 is None else e0.to(f64).contiguous().clone()
        status = torch.zeros(n, dtype=torch.int32, device=self.torch_device)
        iters = torch.zeros(n, dtype=torch.int32, device=self.torch_device)
        opt = _IkOptions(int(max_iters), float(max_translation_error), float(max_rotation_error), float(damping),
                         float(min_step_size), float(max_step_size))
        mask = (1 << self.nq) - 1 if active_mask is None else int(active_mask)
        w = None
        if joint_weights is not None:
            w = np.ascontiguousarray(joint_weights, dtype=np.float64).reshape(self.nq)
        ns = None
        if nullspace is not None:
            ns = torch.as_tensor(nullspace, device=self.torch_device).to(f64).reshape(n, self.nq).contiguous()
        ptr = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
        _check(self.lib.pbp_ik_iterate(self._h, int(frame_id), ptr(Td), ptr(Qd), n, ctypes.byref(opt), ctypes.c_uint64(mask),
                                       ctypes.c_void_p(w.ctypes.data) if w is not None else None, ptr(ns), ptr(e0), ptr(status),
                                       ptr(iters), self._stream()))
        return Qd, status, iters, e0

    # ------------------------------------------------------------------ per-pair distances / avoidance
    def compute_distances(self, Q, cutoff=np.inf, witness=True):
        """Distance (+ witness points and normal) of every active pair for every state.

        CUDA tensor in -> CUDA tensors out, numpy in -> numpy out:
        signed dist [N, npairs] (negative = penetration depth); with ``witness`` also p1, p2, normal
        [N, npairs, 3] (world frame, normal from geometry 1 to 2, p2 - p1 = normal * dist; zero for pairs
        farther than ``cutoff``, whose dist is a lower bound).
        The device layout is pair-major; the returned arrays are transposed views."""
        torch = _torch()
        host = not self._is_dev(Q)
        Qd = torch.as_tensor(self._host_q(Q, self.nq), device=self.torch_device) if host else self._dev_q(Q)
        n, P = Qd.shape[0], self.npairs
        dist = torch.empty((P, n), dtype=torch.float32, device=self.torch_device)
        outs = [torch.empty((P, n, 3), dtype=torch.float32, device=self.torch_device) for _ in range(3)] if witness else [None] * 3
        ptr = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
        if P > 0:
            _check(self.lib.pbp_compute_distances(self._h, ptr(Qd), n, float(min(cutoff, 3.0e38)) if np.isfinite(cutoff) else float("inf"),
                                                  ptr(dist), ptr(outs[0]), ptr(outs[1]), ptr(outs[2]), self._stream()))
        res = [dist.transpose(0, 1)] + ([o.transpose(0, 1) for o in outs] if witness else [])
        if host:
            res = [r.cpu().numpy() for r in res]
        return res[0] if not witness else tuple(res)

    def collision_nullspace(self, Q, dist_padding=0.05, gain=1.0):
        """Batched collision-avoidance null-space term [N, nq] (float32); see pbp_collision_nullspace."""
        torch = _torch()
        host = not self._is_dev(Q)
        Qd = torch.as_tensor(self._host_q(Q, self.nq), device=self.torch_device) if host else self._dev_q(Q)
        out = torch.empty((Qd.shape[0], self.nq), dtype=torch.float32, device=self.torch_device)
        _check(self.lib.pbp_collision_nullspace(self._h, ctypes.c_void_p(Qd.data_ptr()), Qd.shape[0], float(dist_padding), float(gain),
                                                ctypes.c_void_p(out.data_ptr()), self._stream()))
        return out.cpu().numpy() if host else out

    # ------------------------------------------------------------------ k nearest neighbours
    KNN_QUERY_BLOCK = 16

    def knn(self, nodes, k, radius=np.inf, predecessors_only=False, block_first=0, block_stride=1):
        """k nearest nodes of every node (CUDA tensor [N, nq]) -> (idx [N, k] int32, -1 padded; dist [N, k] float32).

        ``block_first`` / ``block_stride``: only every ``block_stride``-th block of 16 consecutive query nodes,
        starting with block ``block_first`` (the interleaved share of one rank, see pbp_knn_blocks); the result
        then has one row per query of those blocks, in that order (rows past N hold -1 / inf)."""
        torch = _torch()
        Nd = self._dev_q(nodes)
        n = Nd.shape[0]
        nblocks = (n + self.KNN_QUERY_BLOCK - 1) // self.KNN_QUERY_BLOCK
        mine = max(0, (nblocks - int(block_first) + int(block_stride) - 1) // int(block_stride))
        rows = n if (block_first, block_stride) == (0, 1) else mine * self.KNN_QUERY_BLOCK
        idx = torch.full((rows, int(k)), -1, dtype=torch.int32, device=self.torch_device)
        dist = torch.full((rows, int(k)), float("inf"), dtype=torch.float32, device=self.torch_device)
        ptr = lambda t: ctypes.c_void_p(t.data_ptr())
        if rows:
            _check(self.lib.pbp_knn_blocks(self._h, ptr(Nd), n, int(k), float(radius), int(bool(predecessors_only)), int(block_first),
                                           int(block_stride), ptr(idx), ptr(dist), self._stream()))
        return idx, dist

    # ------------------------------------------------------------------ sampling
    def sample_free_states(self, n, seed=0, joint_padding=0.0, padding=0.0, max_rounds=64, as_numpy=False):
        """Up to n collision-free configurations (uniform in the padded joint limits), in draw order."""
        torch = _torch()
        out = torch.empty((n, self.nq), dtype=torch.float32, device=self.torch_device)
        found, drawn = ctypes.c_int64(0), ctypes.c_int64(0)
        _check(self.lib.pbp_sample_free_states(self._h, n, ctypes.c_uint64(seed), float(joint_padding), float(padding),
                                               int(max_rounds), ctypes.c_void_p(out.data_ptr()), ctypes.byref(found),
                                               ctypes.byref(drawn), self._stream()))
        out = out[: found.value]
        self.last_sample_draws = drawn.value
        return out.cpu().numpy() if as_numpy else out


def measure_fp32_peak(device=0):
    """Measured FP32 FMA throughput of `device` in TFLOP/s (microbenchmark inside the library)."""
    lib = load_library()
    out = ctypes.c_double(0.0)
    _check(lib.pbp_measure_fp32_peak(int(device), ctypes.byref(out)))
    return out.value


_FOREIGN_CACHE = {}    # (id(model), id(collision_model)) -> (key, engine): Pinocchio objects that refuse new attributes
_FOREIGN_CACHE_MAX = 8


def _cheap_key(model, collision_model, device):
    """What `get_engine` compares on every call: object identities, the containers' revision counters and
    the counts an add / remove changes.  In-place edits it cannot see (a placement assigned through
    ``geometryObjects[i].placement``, limits overwritten) need `invalidate_engine`, or PBP_STRICT_CACHE=1,
    which adds the content fingerprint of model/pinocchio_adapter.py to the key (a walk over the model
    per call)."""
    key = (id(model), id(collision_model), getattr(model, "_revision", 0), getattr(collision_model, "_revision", 0),
           int(model.njoints), len(collision_model.geometryObjects), len(collision_model.collisionPairs), device)
    if os.environ.get("PBP_STRICT_CACHE") == "1":
        from .model.pinocchio_adapter import fingerprint

        key += (fingerprint(model, collision_model),)
    return key


def _cache_get(collision_model, model):
    cached = getattr(collision_model, "_pbp_engine_cache", None)
    return cached if cached is not None else _FOREIGN_CACHE.get((id(model), id(collision_model)))


def _cache_put(collision_model, model, entry):
    try:
        collision_model._pbp_engine_cache = entry
    except (AttributeError, TypeError):     # an extension type without a __dict__
        if len(_FOREIGN_CACHE) >= _FOREIGN_CACHE_MAX:
            _, old = _FOREIGN_CACHE.pop(next(iter(_FOREIGN_CACHE)))
            old.close()
        _FOREIGN_CACHE[(id(model), id(collision_model))] = entry


def invalidate_engine(collision_model, model=None):
    """Forget the engine compiled for ``collision_model`` (call after editing a model in place)."""
    cached = _cache_get(collision_model, model)
    if cached is not None:
        cached[1].close()
    try:
        collision_model._pbp_engine_cache = None
    except (AttributeError, TypeError):
        pass
    for k in [k for k in _FOREIGN_CACHE if k[1] == id(collision_model)]:
        _FOREIGN_CACHE.pop(k)


def get_engine(model, collision_model, device=None):
    """Engine cached on the collision model; rebuilt when the model or its pair list changes.

    ``model`` / ``collision_model`` are this package's containers or Pinocchio objects
    (``pinocchio.Model`` / ``pinocchio.GeometryModel``, as the reference's callers hold them,
    /root/reference/src/pyroboplan/core/utils.py:8-15): those are converted once by
    ``model.pinocchio_adapter.from_pinocchio`` (same joint / frame / geometry / pair indices)."""
    key = _cheap_key(model, collision_model, device)
    cached = _cache_get(collision_model, model)
    if cached is not None and cached[0] == key:
        return cached[1]
    if cached is not None:
        cached[1].close()
    eng = CollisionEngine(model, collision_model, device=device)
    _cache_put(collision_model, model, (key, eng))
    return eng
