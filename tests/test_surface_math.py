"""The arithmetic of the enclosure test (k_surface_refine, pbp_kernels.cuh) replayed in numpy on the
tables `flatten_model` hands to the engine, against the oracle's exact triangle tests: for a shape
inside a convex mesh surface, `min_f (w_f - h_B(n_f))` over the hull's face planes IS its distance to
the surface, and the static enclosure flags never miss a pair in which an enclosure occurs."""

import numpy as np

from conftest import make_panda, uniform_states
from oracle.oracle import Oracle
from pyroboplan_b200.engine import flatten_model
from test_gpu_surface import nested_shell_model


def _world(T, p):
    return p @ T[:9].reshape(3, 3).T + T[9:]


def _support_values(geom, T, normals):
    """h(n) = max over the shape of n.x for every row n of `normals` (world frame)."""
    s = geom.geometry
    name = type(s).__name__
    if name == "Sphere":
        return normals @ T[9:] + s.radius
    if name == "Box":
        h = s.halfSide
        pts = np.array([[sx * h[0], sy * h[1], sz * h[2]] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)])
    else:
        assert name == "Convex"
        pts = np.asarray(s.points)
    return (_world(T, pts) @ normals.T).max(axis=0)


def test_plane_clearance_equals_triangle_distance_for_enclosed_shapes():
    model, cm = nested_shell_model(3)
    arrays, _ = flatten_model(model, cm)
    orc, hull = Oracle(model, cm), Oracle(model, cm, mesh_semantics="hull")
    geoms = cm.geometryObjects
    q = uniform_states(model, 4000, 7).astype(np.float64)
    checked = {1: 0, 2: 0}
    for x in q:
        _, oMg = orc.fk(x)
        for k, p in enumerate(cm.collisionPairs):
            flag = int(arrays["pair_enclose"][k])
            outer, inner = (p.first, p.second) if flag == 1 else (p.second, p.first)
            if hull.pair_distance_placed(oMg, p.first, p.second)[0] > 0.0:
                continue
            exact = orc.soup_pair_distance(oMg, p.first, p.second)
            off, cnt = int(arrays["geom_plane_off"][outer]), int(arrays["geom_plane_cnt"][outer])
            pl = arrays["planes"][off:off + cnt].astype(np.float64)
            Ro, to = oMg[outer][:9].reshape(3, 3), oMg[outer][9:]
            n_world = pl[:, :3] @ Ro.T                                             # face normals in the world frame
            w_world = pl[:, 3] + n_world @ to                                      # n.(R x + t) <= w + n.t
            gap = (w_world - _support_values(geoms[inner], oMg[inner], n_world)).min()   # min_f (w_f - h_inner(n_f))
            if exact > 0.0:                                                        # enclosed: no triangle is touched
                assert abs(gap - exact) < 1e-6, (k, gap, exact)
                checked[flag] += 1
            else:
                assert gap <= 1e-6, (k, gap)
    assert checked[1] > 50 and checked[2] > 20


def test_enclosure_flags_cover_every_enclosure_on_the_panda():
    """Whenever the hulls of a Panda pair overlap while their surfaces do not, the pair carries a flag."""
    model, cmodel, _ = make_panda()
    arrays, _ = flatten_model(model, cmodel)
    orc, hull = Oracle(model, cmodel), Oracle(model, cmodel, mesh_semantics="hull")
    q = uniform_states(model, 40000, 3).astype(np.float64)
    hs = orc.check_states(q)[0].astype(bool)
    hh = hull.check_states(q)[0].astype(bool)
    seen = 0
    for i in np.nonzero(hh & ~hs)[0]:
        _, oMg = orc.fk(q[i])
        for k, p in enumerate(cmodel.collisionPairs):
            if hull.pair_distance_placed(oMg, p.first, p.second)[0] > 0.0:
                continue
            if hull.pair_signed_distance_placed(oMg, p.first, p.second)[0] < -1e-3:      # a real enclosure, not a dent
                assert arrays["pair_enclose"][k] & 1
                seen += 1
    assert seen > 20


def test_inner_balls_lie_inside_their_cores():
    """The hit certificate of the item setup (pbp_surface.cuh::inner_balls_hit) is only sound if every ball lies
    inside the core (inside the hull for a solid): checked against the core's face planes for every bundled mesh."""
    from pyroboplan_b200.engine import flatten_model
    from pyroboplan_b200.model.surface import surface_tables
    from pyroboplan_b200.models import panda, ur5

    for load in (panda.load_models, ur5.load_ur5_on_base_models):
        model, cmodel, _ = load()
        arrays, _ = flatten_model(model, cmodel)
        seen = spread = 0
        for g, obj in enumerate(cmodel.geometryObjects):
            shape = obj.geometry
            if getattr(shape, "mesh_triangles", None) is None:
                continue
            st = surface_tables(shape)
            balls = arrays["geom_inner_balls"][g].astype(np.float64)
            if not st["has_core"]:
                assert (balls[:, 3] == 0).all()          # no core: nothing may certify a hit
                continue
            kp = st["core_hull_planes"]
            for c in balls[balls[:, 3] > 0]:
                assert (kp[:, :3] @ c[:3] + c[3] <= kp[:, 3] + 1e-9).all(), (obj.name, c)
                seen += 1
            used = balls[balls[:, 3] > 0]
            spread = max(spread, len(used))
            if len(used) >= 2:   # the balls are not all the same one
                assert np.linalg.norm(used[0, :3] - used[1, :3]) > 1e-3
        assert seen >= 14 and spread == 4


def _l_shaped_prism():
    """A closed, consistently oriented NON-convex mesh (an L-shaped prism) whose hull's Chebyshev centre lies
    OUTSIDE the solid -- in the notch of the L."""
    poly = np.array([[0, 0], [3, 0], [3, 0.4], [0.4, 0.4], [0.4, 3], [0, 3]], dtype=np.float64)   # counter-clockwise
    n = len(poly)
    verts = np.array([[x, y, z] for z in (0.0, 2.0) for x, y in poly])
    tris = []
    for i in range(n):                                  # side walls, outward
        j = (i + 1) % n
        tris += [[i, j, n + j], [i, n + j, n + i]]
    fan = [[0, 1, 2], [0, 2, 3], [0, 3, 4], [0, 4, 5]]   # the L is star-shaped from its inner corner vertex 0
    tris += [[a, c, b] for a, b, c in fan]              # bottom, facing -z
    tris += [[n + a, n + b, n + c] for a, b, c in fan]  # top, facing +z
    return verts, np.array(tris)


def test_interior_point_of_a_non_convex_mesh():
    """model/surface.py::_interior_point: the inner cell must be built around a point INSIDE the solid; the hull's
    centre of an L-shaped prism is not one."""
    from pyroboplan_b200.model import Convex
    from pyroboplan_b200.model.surface import _interior_point, _is_closed, _winding_number, surface_tables

    verts, tris = _l_shaped_prism()
    assert _is_closed(tris)
    shape = Convex(verts, triangles=tris)
    centre, _ = shape.inner_sphere()
    assert abs(_winding_number(np.asarray(centre), verts, tris)) < 0.5          # the hull's centre is in the notch
    p = _interior_point(verts, tris, np.asarray(centre))
    assert p is not None and abs(_winding_number(p, verts, tris)) > 0.5
    in_l = (0 < p[2] < 2.0) and ((0 < p[0] < 3 and 0 < p[1] < 0.4) or (0 < p[0] < 0.4 and 0 < p[1] < 3))
    assert in_l, p
    st = surface_tables(shape)
    if st["has_core"]:      # whatever cell was built, it lies inside the L (every core vertex does)
        for v in st["core"].astype(np.float64):
            assert (-1e-6 <= v[2] <= 2.0 + 1e-6) and ((v[1] <= 0.4 + 1e-6) or (v[0] <= 0.4 + 1e-6)), v
