"""ctypes front end of the float64 CPU oracle (``pbp_oracle.c``).

TEST INFRASTRUCTURE ONLY -- see the header of ``pbp_oracle.c``.  Imported by ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs; never by ``pyroboplan_b200``.

The oracle flattens the host model containers (``pyroboplan_b200.model``: the neutral
description of joints, placements, shapes, pairs) into its own float64 tables; it shares no
code with the engine's float32 model compiler (``pyroboplan_b200/engine.py``).
"""

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libpbp_oracle.so")

SH_SPHERE, SH_BOX, SH_CYLINDER, SH_CAPSULE, SH_CONVEX = range(5)


def build(force=False):
    """Compile ``libpbp_oracle.so`` with gcc (no-op if it is newer than the source)."""
    src = os.path.join(_HERE, "pbp_oracle.c")
    if (not force) and os.path.exists(_LIB_PATH) and os.path.getmtime(_LIB_PATH) >= os.path.getmtime(src):
        return _LIB_PATH
    subprocess.check_call(
        ["gcc", "-O2", "-fPIC", "-pthread", "-std=gnu11", "-shared", "-o", _LIB_PATH, src, "-lm"]
    )
    return _LIB_PATH


class _CModel(ctypes.Structure):
    _fields_ = [
        ("nq", ctypes.c_int32),
        ("njoints", ctypes.c_int32),
        ("ngeoms", ctypes.c_int32),
        ("npairs", ctypes.c_int32),
        ("nverts", ctypes.c_int32),
        ("joint_parent", ctypes.c_void_p),
        ("joint_type", ctypes.c_void_p),
        ("joint_idx_q", ctypes.c_void_p),
        ("joint_placement", ctypes.c_void_p),
        ("joint_axis", ctypes.c_void_p),
        ("geom_parent", ctypes.c_void_p),
        ("geom_shape", ctypes.c_void_p),
        ("geom_placement", ctypes.c_void_p),
        ("geom_params", ctypes.c_void_p),
        ("geom_vert_off", ctypes.c_void_p),
        ("geom_vert_cnt", ctypes.c_void_p),
        ("geom_bsphere", ctypes.c_void_p),
        ("geom_isphere", ctypes.c_void_p),
        ("verts", ctypes.c_void_p),
        ("pairs", ctypes.c_void_p),
        ("tris", ctypes.c_void_p),
        ("tri_off", ctypes.c_void_p),
        ("tri_cnt", ctypes.c_void_p),
    ]


class Counters(ctypes.Structure):
    _fields_ = [
        ("pair_tests", ctypes.c_int64),
        ("gjk_iters", ctypes.c_int64),
        ("support_calls", ctypes.c_int64),
        ("support_verts", ctypes.c_int64),
        ("culled", ctypes.c_int64),
    ]

    def as_dict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_}


_lib = None


def _load():
    global _lib
    if _lib is None:
        build()
        lib = ctypes.CDLL(_LIB_PATH)
        P = ctypes.POINTER
        lib.oracle_fk.argtypes = [P(_CModel), ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        lib.oracle_fk.restype = None
        lib.oracle_pair_distance.argtypes = [P(_CModel), ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        lib.oracle_pair_distance.restype = ctypes.c_double
        lib.oracle_pair_signed_distance.argtypes = [P(_CModel), ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        lib.oracle_pair_signed_distance.restype = ctypes.c_double
        lib.oracle_check_states.argtypes = [
            P(_CModel), ctypes.c_void_p, ctypes.c_int64, ctypes.c_double, ctypes.c_int, ctypes.c_int, ctypes.c_int,
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, P(Counters),
        ]
        lib.oracle_check_states.restype = None
        lib.oracle_check_edges.argtypes = [
            P(_CModel), ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_double, ctypes.c_double,
            ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, P(ctypes.c_int64), P(Counters),
        ]
        lib.oracle_check_edges.restype = None
        lib.oracle_check_path.argtypes = [P(_CModel), ctypes.c_void_p, ctypes.c_int64, ctypes.c_double, ctypes.c_int]
        lib.oracle_check_path.restype = ctypes.c_int
        lib.oracle_segment_samples.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_double]
        lib.oracle_segment_samples.restype = ctypes.c_int64
        lib.oracle_discretize_segment.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_void_p]
        lib.oracle_discretize_segment.restype = None
        lib.oracle_num_threads.restype = ctypes.c_int
        lib.oracle_soup_pair_distance.argtypes = [P(_CModel), ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int,
                                                  ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_int]
        lib.oracle_soup_pair_distance.restype = ctypes.c_double
        lib.oracle_soup_check_states.argtypes = [P(_CModel), ctypes.c_void_p, ctypes.c_int64, ctypes.c_double, ctypes.c_void_p, ctypes.c_void_p,
                                                 ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                                 P(ctypes.c_int64)]
        lib.oracle_soup_check_states.restype = None
        lib.oracle_tri_tri_distance.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        lib.oracle_tri_tri_distance.restype = ctypes.c_double
        _lib = lib
    return _lib


def _se3_row(T):
    return np.concatenate([np.asarray(T.rotation, dtype=np.float64).reshape(9), np.asarray(T.translation, dtype=np.float64)])


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class Oracle:
    """Float64 CPU evaluation of the hot path for one (model, collision_model)."""

    def __init__(self, model, collision_model, mesh_semantics="surface"):
        """mesh_semantics: "surface" -- geometries that carry the triangles of a mesh file are surfaces,
        as Pinocchio's BVHModel meshes are in the reference (verdicts and min distances then come from
        exact triangle tests wherever the hulls are within the padding); "hull" -- convex solids."""
        self.lib = _load()
        self.nq = model.nq
        nj = model.njoints
        geoms = collision_model.geometryObjects
        ng = len(geoms)
        a = {}
        a["joint_parent"] = np.array(model.parents, dtype=np.int32)
        a["joint_type"] = np.array([j.type for j in model.joints], dtype=np.int32)
        a["joint_idx_q"] = np.array([j.idx_q for j in model.joints], dtype=np.int32)
        a["joint_placement"] = np.array([_se3_row(p) for p in model.jointPlacements], dtype=np.float64)
        a["joint_axis"] = np.array([j.axis for j in model.joints], dtype=np.float64)
        a["geom_parent"] = np.array([g.parentJoint for g in geoms], dtype=np.int32).reshape(ng)
        a["geom_shape"] = np.zeros(ng, dtype=np.int32)
        a["geom_placement"] = np.zeros((ng, 12), dtype=np.float64)
        a["geom_params"] = np.zeros((ng, 4), dtype=np.float64)
        a["geom_vert_off"] = np.zeros(ng, dtype=np.int32)
        a["geom_vert_cnt"] = np.zeros(ng, dtype=np.int32)
        a["geom_bsphere"] = np.zeros((ng, 4), dtype=np.float64)
        a["geom_isphere"] = np.zeros((ng, 4), dtype=np.float64)
        verts = []
        nv = 0
        for i, g in enumerate(geoms):
            s = g.geometry
            a["geom_placement"][i] = _se3_row(g.placement)
            ic, ir = s.inner_sphere()
            a["geom_isphere"][i] = [*np.asarray(ic, dtype=np.float64), max(ir - 1e-9, 0.0)]
            name = type(s).__name__
            if name == "Sphere":
                a["geom_shape"][i] = SH_SPHERE
                a["geom_params"][i, 0] = s.radius
                a["geom_bsphere"][i] = [0, 0, 0, s.radius]
            elif name == "Box":
                a["geom_shape"][i] = SH_BOX
                a["geom_params"][i, :3] = s.halfSide
                a["geom_bsphere"][i] = [0, 0, 0, np.linalg.norm(s.halfSide)]
            elif name == "Cylinder":
                a["geom_shape"][i] = SH_CYLINDER
                a["geom_params"][i, :2] = [s.radius, s.halfLength]
                a["geom_bsphere"][i] = [0, 0, 0, np.hypot(s.radius, s.halfLength)]
            elif name == "Capsule":
                a["geom_shape"][i] = SH_CAPSULE
                a["geom_params"][i, :2] = [s.radius, s.halfLength]
                a["geom_bsphere"][i] = [0, 0, 0, s.radius + s.halfLength]
            elif name == "Convex":
                a["geom_shape"][i] = SH_CONVEX
                pts = np.asarray(s.points, dtype=np.float64)
                a["geom_vert_off"][i] = nv
                a["geom_vert_cnt"][i] = len(pts)
                verts.append(pts)
                nv += len(pts)
                c = 0.5 * (pts.min(axis=0) + pts.max(axis=0))
                a["geom_bsphere"][i] = [*c, np.sqrt(((pts - c) ** 2).sum(axis=1).max())]
            else:
                raise TypeError(f"unsupported shape {name}")
        a["verts"] = np.concatenate(verts) if verts else np.zeros((1, 3))
        a["pairs"] = np.array([[p.first, p.second] for p in collision_model.collisionPairs], dtype=np.int32).reshape(-1, 2)
        self.npairs = len(a["pairs"])
        if self.npairs == 0:
            a["pairs"] = np.zeros((1, 2), dtype=np.int32)
        self._arrays = {k: np.ascontiguousarray(v) for k, v in a.items()}
        self.njoints, self.ngeoms = nj, ng
        cm = _CModel()
        cm.nq, cm.njoints, cm.ngeoms, cm.npairs, cm.nverts = self.nq, nj, ng, self.npairs, nv
        for k, v in self._arrays.items():
            setattr(cm, k, v.ctypes.data)
        self._cm = cm
        self.set_mesh_triangles(collision_model)
        self.mesh_semantics = mesh_semantics if self._tri_cnt.any() else "hull"
        if self.mesh_semantics == "surface":
            cm.tris, cm.tri_off, cm.tri_cnt = self._tris.ctypes.data, self._tri_off.ctypes.data, self._tri_cnt.ctypes.data

    # ------------------------------------------------------------------ FK
    def fk(self, q):
        """Returns (oMi [njoints,12], oMg [ngeoms,12]); rows = R row-major (9) then t (3)."""
        q = np.ascontiguousarray(q, dtype=np.float64).reshape(self.nq)
        oMi = np.zeros((self.njoints, 12))
        oMg = np.zeros((self.ngeoms, 12))
        self.lib.oracle_fk(ctypes.byref(self._cm), _ptr(q), _ptr(oMi), _ptr(oMg))
        return oMi, oMg

    def pair_distance(self, q, ga, gb):
        """(distance, witness_a[3], witness_b[3], intersecting) for geometries ga, gb at q."""
        _, oMg = self.fk(q)
        out = np.zeros(7)
        d = self.lib.oracle_pair_distance(ctypes.byref(self._cm), _ptr(oMg), int(ga), int(gb), _ptr(out), None)
        return d, out[:3].copy(), out[3:6].copy(), bool(out[6])

    def pair_distance_placed(self, oMg, ga, gb):
        out = np.zeros(7)
        d = self.lib.oracle_pair_distance(ctypes.byref(self._cm), _ptr(np.ascontiguousarray(oMg)), int(ga), int(gb), _ptr(out), None)
        return d, out[:3].copy(), out[3:6].copy(), bool(out[6])

    def pair_signed_distance(self, q, ga, gb):
        """coal's convention: (signed distance -- negative = penetration depth --, nearest/deepest point
        on a [3], on b [3], unit normal a -> b [3], exact flag) for geometries ga, gb at q."""
        _, oMg = self.fk(q)
        return self.pair_signed_distance_placed(oMg, ga, gb)

    def pair_signed_distance_placed(self, oMg, ga, gb):
        out = np.zeros(10)
        d = self.lib.oracle_pair_signed_distance(ctypes.byref(self._cm), _ptr(np.ascontiguousarray(oMg)), int(ga), int(gb), _ptr(out), None)
        return d, out[:3].copy(), out[3:6].copy(), out[6:9].copy(), bool(out[9])

    # ------------------------------------------------------------------ states
    def check_states(self, Q, padding=0.0, want_all=False, use_cull=True, nthreads=0):
        """hit[N] uint8, min_dist[N] (over evaluated pairs), first_pair[N], counters.

        use_cull: False/0 = every pair through GJK in list order (the reference's loop);
        True/1 = skip pairs whose bounding spheres are apart; 2 = work-counting mode with
        the cull-aware proxy broadphase (same verdicts; first_pair/min_dist not comparable)."""
        Q = np.ascontiguousarray(Q, dtype=np.float64).reshape(-1, self.nq)
        n = len(Q)
        hit = np.zeros(n, dtype=np.uint8)
        md = np.zeros(n, dtype=np.float64)
        fp = np.zeros(n, dtype=np.int32)
        cnt = Counters()
        self.lib.oracle_check_states(
            ctypes.byref(self._cm), _ptr(Q), n, float(padding), int(want_all), int(use_cull), int(nthreads),
            _ptr(hit), _ptr(md), _ptr(fp), ctypes.byref(cnt),
        )
        return hit, md, fp, cnt.as_dict()

    def check_state(self, q, padding=0.0):
        return bool(self.check_states(np.asarray(q).reshape(1, -1), padding, nthreads=1)[0][0])

    def min_distance(self, q):
        if self.npairs == 0:
            return np.inf
        return float(self.check_states(np.asarray(q).reshape(1, -1), 0.0, want_all=True, use_cull=False, nthreads=1)[1][0])

    # ------------------------------------------------------------------ paths / edges
    def check_path(self, q_path, padding=0.0, use_cull=True):
        Q = np.ascontiguousarray(q_path, dtype=np.float64).reshape(-1, self.nq)
        return bool(self.lib.oracle_check_path(ctypes.byref(self._cm), _ptr(Q), len(Q), float(padding), int(use_cull)))

    def check_edges(self, Q1, Q2, step, padding=0.0, use_cull=True, nthreads=0):
        """hit[E] (1 = edge invalid), first_bad[E], states_evaluated, counters."""
        Q1 = np.ascontiguousarray(Q1, dtype=np.float64).reshape(-1, self.nq)
        Q2 = np.ascontiguousarray(Q2, dtype=np.float64).reshape(-1, self.nq)
        n = len(Q1)
        hit = np.zeros(n, dtype=np.uint8)
        fb = np.zeros(n, dtype=np.int32)
        ev = ctypes.c_int64(0)
        cnt = Counters()
        self.lib.oracle_check_edges(
            ctypes.byref(self._cm), _ptr(Q1), _ptr(Q2), n, float(step), float(padding), int(use_cull), int(nthreads),
            _ptr(hit), _ptr(fb), ctypes.byref(ev), ctypes.byref(cnt),
        )
        return hit, fb, int(ev.value), cnt.as_dict()

    def segment_samples(self, q1, q2, step):
        q1 = np.ascontiguousarray(q1, dtype=np.float64)
        q2 = np.ascontiguousarray(q2, dtype=np.float64)
        return int(self.lib.oracle_segment_samples(_ptr(q1), _ptr(q2), len(q1), float(step)))

    def num_threads(self):
        return int(self.lib.oracle_num_threads())

    # ------------------------------------------------------------------ triangle-soup semantics (study)
    def set_mesh_triangles(self, collision_model):
        """Triangle tables for the surface semantics of Pinocchio's BVHModel meshes: geometries that carry
        ``mesh_vertices`` / ``mesh_triangles`` (URDF meshes) become soups, everything else stays a solid."""
        tris, off, cnt = [], [], []
        for g in collision_model.geometryObjects:
            t = getattr(g.geometry, "mesh_triangles", None)
            off.append(sum(len(x) for x in tris))
            if t is None:
                cnt.append(0)
                continue
            v = np.asarray(g.geometry.mesh_vertices, dtype=np.float64)
            tris.append(v[np.asarray(t)].reshape(-1, 9))
            cnt.append(len(t))
        self._tris = np.ascontiguousarray(np.concatenate(tris) if tris else np.zeros((0, 9)))
        self._tri_off = np.asarray(off, dtype=np.int32)
        self._tri_cnt = np.asarray(cnt, dtype=np.int32)

    def soup_pair_distance(self, oMg, ga, gb, stop_at=-1.0, brute=False):
        ta = self._tris[self._tri_off[ga]: self._tri_off[ga] + self._tri_cnt[ga]]
        tb = self._tris[self._tri_off[gb]: self._tri_off[gb] + self._tri_cnt[gb]]
        ta, tb = np.ascontiguousarray(ta), np.ascontiguousarray(tb)
        return float(self.lib.oracle_soup_pair_distance(ctypes.byref(self._cm), _ptr(np.ascontiguousarray(oMg)), int(ga), int(gb),
                                                        _ptr(ta), len(ta), _ptr(tb), len(tb), float(stop_at), int(brute)))

    def soup_check_states(self, Q, padding=0.0, nthreads=0):
        """(hull_hit[N], soup_hit[N], pairs per state that hit by hull but not by surface, #triangle tests run)."""
        Q = np.ascontiguousarray(Q, dtype=np.float64).reshape(-1, self.nq)
        n = len(Q)
        hh, sh, fl = np.zeros(n, np.uint8), np.zeros(n, np.uint8), np.zeros(n, np.int32)
        tests = ctypes.c_int64(0)
        self.lib.oracle_soup_check_states(ctypes.byref(self._cm), _ptr(Q), n, float(padding), _ptr(self._tris), _ptr(self._tri_off),
                                          _ptr(self._tri_cnt), int(nthreads), _ptr(hh), _ptr(sh), _ptr(fl), ctypes.byref(tests))
        return hh, sh, fl, int(tests.value)


def discretize_joint_space_path(q_path, max_angle_distance):
    """Oracle restatement of reference planning/utils.py:114-140 (list in, list out)."""
    lib = _load()
    out = []
    for idx in range(1, len(q_path)):
        a = np.ascontiguousarray(q_path[idx - 1], dtype=np.float64)
        b = np.ascontiguousarray(q_path[idx], dtype=np.float64)
        n = int(lib.oracle_segment_samples(_ptr(a), _ptr(b), len(a), float(max_angle_distance)))
        buf = np.zeros((n, len(a)))
        lib.oracle_discretize_segment(_ptr(a), _ptr(b), len(a), n, _ptr(buf))
        out.extend(list(buf))
    return out


def tri_tri_distance(a, b):
    """Exact float64 distance between two triangles ([3, 3] each; 0 when they intersect)."""
    a = np.ascontiguousarray(a, dtype=np.float64).reshape(9)
    b = np.ascontiguousarray(b, dtype=np.float64).reshape(9)
    return float(_load().oracle_tri_tri_distance(_ptr(a), _ptr(b)))
