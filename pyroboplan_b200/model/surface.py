"""Surface tables of mesh geometries: what the engine needs to answer EXACTLY like a triangle mesh.

Pinocchio loads a URDF ``<mesh>`` as a coal ``BVHModel`` -- the surface triangles
(/root/reference/src/pyroboplan/models/panda.py:27 -> ``pinocchio.buildModelsFromUrdf``).  The engine's
hot kernels run GJK on convex polytopes, so a mesh M (enclosing the solid S) is bracketed by two of
them:

* the hull ``H`` of its vertices (``S`` lies inside ``H``): ``d(H_a, H_b)`` is a LOWER bound of the
  surface distance, a pair whose hulls are apart by more than the margin is certainly free;
* the inner core ``K`` = the intersection of the inner half-spaces of all mesh triangles (``K`` lies
  inside ``S`` when the mesh is closed and ``K`` is not empty): ``d(K_a, K_b)`` is an UPPER bound, a
  pair whose cores are within the margin certainly collides (or one shape encloses the other).

Only the items between the two bounds (the Panda's hulls and cores differ by 0.3 - 1 mm around the
near-planar quads its meshes triangulate along the valley) need the triangles themselves; the
kernels then test triangle pairs exactly (``pbp_surface.cuh``).  Everything here is host-side model
compilation (numpy / qhull), done once per geometry and cached on it.
"""

import numpy as np

_CACHE_ATTR = "_pbp_surface_tables"
_CACHE = {}            # content hash of (vertices, triangles) -> tables: models are loaded many times per process
_CORE_DELTA = 1e-4     # the simplified core may lie this far inside the exact one (metres)


def _simplify_core(hull_pts, kfull, kplanes, delta):
    """A subset of the exact core's vertices whose hull is within ``delta`` of it: the hull vertices the
    cuts left alone, then the farthest remaining core vertex until none is farther than delta.  (The
    exact core of a Panda link has ~1.7 x the hull's vertices, most of them on fans of nearly coincident
    cuts; the subset has about as many as the hull, so the GJK tables keep their size.  Any subset of
    core vertices spans a subset of the core: it stays inside the mesh.)  Returns float32 vertices."""
    from scipy.spatial import ConvexHull

    uncut = (hull_pts @ kplanes[:, :3].T - kplanes[:, 3]).max(axis=1) <= 1e-9
    pts = [p for p in hull_pts[uncut]]
    if len(pts) < 4:
        pts = [p for p in kfull]
    cand = kfull
    for _ in range(len(kfull)):
        eq = ConvexHull(np.array(pts)).equations
        ex = (cand @ eq[:, :3].T + eq[:, 3]).max(axis=1)
        j = int(np.argmax(ex))
        if ex[j] < delta:
            break
        pts.append(cand[j])
    p32 = np.unique(np.array(pts).astype(np.float32), axis=0)
    h = ConvexHull(p32.astype(np.float64))
    return np.ascontiguousarray(p32[np.sort(h.vertices)])


def _unit_planes_of_triangles(verts, tris, inside):
    """Outward unit plane (n, w), n.x <= w on the inner side, of every non-degenerate triangle;
    oriented by the interior point ``inside``.  Returns (planes [m, 4] = nx ny nz w, kept mask)."""
    t = verts[tris]
    n = np.cross(t[:, 1] - t[:, 0], t[:, 2] - t[:, 0])
    ln = np.linalg.norm(n, axis=1)
    scale = max(float(np.abs(verts).max()), 1e-30)
    ok = ln > 1e-12 * scale * scale
    n = n[ok] / ln[ok, None]
    w = np.einsum("ij,ij->i", n, t[ok, 0])
    flip = n @ inside - w > 0.0
    n[flip] *= -1.0
    w[flip] *= -1.0
    return np.concatenate([n, w[:, None]], axis=1), ok


def _is_closed(tris):
    """Every undirected edge is shared by exactly two triangles."""
    e = np.sort(np.concatenate([tris[:, [0, 1]], tris[:, [1, 2]], tris[:, [2, 0]]]), axis=1)
    _, cnt = np.unique(e, axis=0, return_counts=True)
    return bool(len(cnt)) and bool((cnt == 2).all())


def _winding_number(p, verts, tris):
    """Generalised winding number of the closed triangle mesh around p (Van Oosterom - Strackee solid angles):
    +-1 inside a consistently oriented closed mesh, 0 outside."""
    a, b, c = verts[tris[:, 0]] - p, verts[tris[:, 1]] - p, verts[tris[:, 2]] - p
    la, lb, lc = np.linalg.norm(a, axis=1), np.linalg.norm(b, axis=1), np.linalg.norm(c, axis=1)
    num = np.einsum("ij,ij->i", a, np.cross(b, c))
    den = la * lb * lc + np.einsum("ij,ij->i", a, b) * lc + np.einsum("ij,ij->i", b, c) * la + np.einsum("ij,ij->i", c, a) * lb
    return float(np.arctan2(num, den).sum() / (2.0 * np.pi))


def _inside_closed_mesh(p, verts, tris):
    """Is p inside the closed triangle mesh?  Crossing parity of three fixed rays, majority vote: independent of
    the triangles' orientation (qhull simplices and many exported meshes are not consistently wound, which rules
    out the winding number alone); the winding number settles a split vote."""
    t = verts[tris]
    a, e1, e2 = t[:, 0], t[:, 1] - t[:, 0], t[:, 2] - t[:, 0]
    votes = 0
    for d in (np.array([0.5377, 0.8339, -0.1253]), np.array([-0.3188, 0.2762, 0.9067]), np.array([0.8622, -0.4337, 0.2618])):
        d = d / np.linalg.norm(d)
        h = np.cross(d, e2)
        det = np.einsum("ij,ij->i", e1, h)
        ok = np.abs(det) > 1e-14
        inv = np.where(ok, 1.0 / np.where(ok, det, 1.0), 0.0)
        s = p - a
        u = np.einsum("ij,ij->i", s, h) * inv
        qv = np.cross(s, e1)
        v = (qv @ d) * inv
        tt = np.einsum("ij,ij->i", e2, qv) * inv
        crossings = int((ok & (u >= 0) & (v >= 0) & (u + v <= 1) & (tt > 1e-12)).sum())
        votes += crossings & 1
    if votes in (0, 3):
        return votes == 3
    return abs(_winding_number(p, verts, tris)) > 0.5 if votes == 2 else False


def _interior_point(verts, tris, first_guess):
    """A point strictly inside the closed mesh, or None: the guess (the hull's Chebyshev centre), else the
    candidate with the most clearance among points pushed inwards from the triangle centroids.  A convex cell
    of the triangle planes around ANY interior point lies inside the solid (no triangle crosses the open cell,
    and the cell is connected), which is all the inner core needs; for a non-convex mesh (UR upper arm, forearm)
    the hull's centre may lie outside the solid."""
    if _inside_closed_mesh(first_guess, verts, tris):
        return first_guess
    t = verts[tris]
    cen = t.mean(axis=1)
    n = np.cross(t[:, 1] - t[:, 0], t[:, 2] - t[:, 0])
    ln = np.linalg.norm(n, axis=1)
    ok = ln > 0
    n = n[ok] / ln[ok, None]
    cen = cen[ok]
    size = float(np.linalg.norm(verts.max(axis=0) - verts.min(axis=0)))
    best, best_clear = None, 0.0
    for depth in (0.02, 0.05, 0.1):
        for sign in (-1.0, 1.0):      # the orientation convention of the file is not assumed
            for p in (cen + sign * depth * size * n)[:: max(1, len(cen) // 64)]:
                if _inside_closed_mesh(p, verts, tris):
                    clear = float(_point_triangles_dist(p[None, :], t).min())
                    if clear > best_clear:
                        best, best_clear = p, clear
    return best


def _feasible_point_near(v, planes, centre, iters=60):
    """A point of {x : n.x <= w} close to v: cyclic projection on the most violated plane, then a pull
    towards the (strictly interior) centre.  |v - result| bounds the distance of v to the set."""
    x = v.copy()
    n, w = planes[:, :3], planes[:, 3]
    for _ in range(iters):
        e = n @ x - w
        k = int(np.argmax(e))
        if e[k] <= 0.0:
            return x
        x = x - e[k] * n[k]
    e = n @ x - w
    d = n @ (x - centre)
    t = np.where(e > 0.0, e / np.maximum(d, 1e-300), 0.0).max()
    return x + min(t, 1.0) * (centre - x)


def _point_triangles_dist(p, tri):
    """Distances from the points p [m, 3] to the closest of the triangles tri [k, 3, 3] (vectorised
    Ericson 5.1.5 on all pairs).  Used at model compilation only (a few thousand points)."""
    a, b, c = tri[None, :, 0], tri[None, :, 1], tri[None, :, 2]
    P = p[:, None, :]
    ab, ac, ap = b - a, c - a, P - a
    d1, d2 = (ab * ap).sum(-1), (ac * ap).sum(-1)
    bp = P - b
    d3, d4 = (ab * bp).sum(-1), (ac * bp).sum(-1)
    cp = P - c
    d5, d6 = (ab * cp).sum(-1), (ac * cp).sum(-1)
    va, vb, vc = d3 * d6 - d5 * d4, d5 * d2 - d1 * d6, d1 * d4 - d3 * d2
    best = np.full(d1.shape, np.inf)

    def upd(mask, q):
        d = np.linalg.norm(P - q, axis=-1)
        np.copyto(best, np.where(mask & (d < best), d, best))

    with np.errstate(divide="ignore", invalid="ignore"):
        upd((d1 <= 0) & (d2 <= 0), a + 0 * P)
        upd((d3 >= 0) & (d4 <= d3), b + 0 * P)
        upd((d6 >= 0) & (d5 <= d6), c + 0 * P)
        t = np.clip(np.nan_to_num(d1 / (d1 - d3)), 0, 1)
        upd((vc <= 0) & (d1 >= 0) & (d3 <= 0), a + t[..., None] * ab)
        t = np.clip(np.nan_to_num(d2 / (d2 - d6)), 0, 1)
        upd((vb <= 0) & (d2 >= 0) & (d6 <= 0), a + t[..., None] * ac)
        t = np.clip(np.nan_to_num((d4 - d3) / ((d4 - d3) + (d5 - d6))), 0, 1)
        upd((va <= 0) & ((d4 - d3) >= 0) & ((d5 - d6) >= 0), b + t[..., None] * (c - b))
        den = va + vb + vc
        v = np.nan_to_num(vb / den)
        w = np.nan_to_num(vc / den)
        inside = (va > 0) & (vb > 0) & (vc > 0)
        upd(inside, a + v[..., None] * ab + w[..., None] * ac)
    # whatever region test failed numerically: the vertices are always valid candidates
    for q in (a, b, c):
        upd(np.ones_like(inside), q + 0 * P)
    return best.min(axis=1)


def surface_tables(shape):
    """Tables of a mesh geometry (``Convex`` with ``mesh_triangles``), cached on the shape:

    ``mesh_verts``  [nv, 3] float32   unique mesh vertices (STL coordinates are float32 already)
    ``tris``        [nt, 3] int32     vertex indices of the non-degenerate triangles
    ``core``        [nk, 3] float32   vertices of the inner core K (== hull vertices when there is none)
    ``core_eps``    float             every point of the hull H is within this of K (metres); inf without a core
    ``core_planes`` [np, 4] float64   n.x <= w planes of K: the hull's face planes first ...
    ``n_hull_planes`` int             ... this many; the rest are planes of mesh triangles below the hull
    ``plane_areas`` [np]              area of the face behind each plane (hull planes), 0 for the rest
    ``surface_tol`` float             two-sided distance between the hull's boundary and the mesh (metres)
    ``core_hull_planes`` [m, 4]       face planes of the (simplified) core actually used: its inscribed radius
    ``has_core``    bool
    """
    cached = getattr(shape, _CACHE_ATTR, None)
    if cached is not None:
        return cached
    from scipy.spatial import ConvexHull, HalfspaceIntersection
    import hashlib

    v64 = np.asarray(shape.mesh_vertices, dtype=np.float64)
    key = hashlib.sha1(np.ascontiguousarray(v64).tobytes() + np.ascontiguousarray(np.asarray(shape.mesh_triangles, dtype=np.int64)).tobytes()).hexdigest()
    if key in _CACHE:
        setattr(shape, _CACHE_ATTR, _CACHE[key])
        return _CACHE[key]
    v32 = v64.astype(np.float32)
    uv, inv = np.unique(v32, axis=0, return_inverse=True)
    tris = inv.reshape(-1)[np.asarray(shape.mesh_triangles, dtype=np.int64)].astype(np.int32)
    nondeg = (tris[:, 0] != tris[:, 1]) & (tris[:, 1] != tris[:, 2]) & (tris[:, 0] != tris[:, 2])
    tris = np.ascontiguousarray(tris[nondeg])
    V = uv.astype(np.float64)

    hull_planes = np.concatenate([shape.planes[:, :3], -shape.planes[:, 3:4]], axis=1)    # n.x <= w
    hull_areas = np.asarray(shape.plane_areas, dtype=np.float64)
    centre, r_in = shape.inner_sphere()
    centre = np.asarray(centre, dtype=np.float64)

    out = {
        "mesh_verts": np.ascontiguousarray(uv), "tris": tris,
        "core": np.ascontiguousarray(np.asarray(shape.points, dtype=np.float32)), "core_eps": np.inf,
        "core_planes": hull_planes, "n_hull_planes": len(hull_planes), "plane_areas": hull_areas,
        "surface_tol": 0.0, "has_core": False, "core_hull_planes": hull_planes,
    }

    # ---- two-sided distance between the hull's boundary and the mesh surface
    T = V[tris]
    ts = np.linspace(0.0, 1.0, 7)[1:-1]
    samples = [T.mean(axis=1)]
    for i, j in ((0, 1), (1, 2), (2, 0)):
        samples.append((T[:, i, None, :] * (1 - ts)[None, :, None] + T[:, j, None, :] * ts[None, :, None]).reshape(-1, 3))
    x = np.concatenate(samples)
    below = (-(x @ hull_planes[:, :3].T - hull_planes[:, 3])).min(axis=1)     # mesh point -> nearest hull plane
    tol_in = float(max(below.max(), 0.0))
    hull = ConvexHull(np.asarray(shape.points, dtype=np.float64))
    HT = np.asarray(shape.points, dtype=np.float64)[hull.simplices]
    bary = np.array([[1 / 3, 1 / 3, 1 / 3], [0.5, 0.5, 0], [0, 0.5, 0.5], [0.5, 0, 0.5], [0.6, 0.2, 0.2], [0.2, 0.6, 0.2], [0.2, 0.2, 0.6]])
    hp = np.einsum("sb,fbk->fsk", bary, HT).reshape(-1, 3)
    # only hull faces that are not mesh triangles themselves can be away from the mesh
    tol_out = 0.0
    for s0 in range(0, len(hp), 256):
        tol_out = max(tol_out, float(_point_triangles_dist(hp[s0:s0 + 256], T).max()))
    out["surface_tol"] = 1.5 * max(tol_in, tol_out) + 1e-7     # samples, not maxima: keep a margin

    # ---- inner core: K = hull /\ the half-spaces of every mesh triangle that hold an interior point of the solid
    closed = _is_closed(tris)
    inside = _interior_point(V, tris, centre) if closed else None
    if inside is not None:
        centre = inside
    tri_planes, _ = _unit_planes_of_triangles(V, tris, centre)
    slack_c = (tri_planes[:, :3] @ centre - tri_planes[:, 3]).max() if len(tri_planes) else 0.0
    if closed and inside is not None and slack_c < -1e-6 and r_in > 1e-6:
        # the planes of mesh triangles that lie ON a hull face add nothing
        onhull = np.zeros(len(tri_planes), dtype=bool)
        for k, pl in enumerate(tri_planes):
            onhull[k] = bool((np.abs(hull_planes[:, :3] @ pl[:3] - 1.0) < 1e-9).any() and
                             (np.abs(hull_planes[:, 3] - pl[3])[np.abs(hull_planes[:, :3] @ pl[:3] - 1.0) < 1e-9] < 1e-9).any())
        extra = tri_planes[~onhull]
        allp = np.concatenate([hull_planes, extra])
        try:
            hs = HalfspaceIntersection(np.concatenate([allp[:, :3], -allp[:, 3:4]], axis=1), centre)
            kfull = np.unique(np.round(hs.intersections, 12), axis=0)
            core = _simplify_core(np.asarray(shape.points, dtype=np.float64), kfull, allp, _CORE_DELTA)
            kh = ConvexHull(core.astype(np.float64))
            # every hull vertex is within eps of the float32 core actually used on the device
            kp = np.unique(np.round(kh.equations, 10), axis=0)
            kpl = np.concatenate([kp[:, :3], -kp[:, 3:4]], axis=1)
            kc = core.astype(np.float64).mean(axis=0)
            eps = 0.0
            for p in np.asarray(shape.points, dtype=np.float64):
                if (kpl[:, :3] @ p - kpl[:, 3]).max() <= 0.0:
                    continue
                eps = max(eps, float(np.linalg.norm(p - _feasible_point_near(p, kpl, kc))))
            out.update(core=core, core_eps=eps * 1.02 + 2e-7, core_planes=allp, plane_areas=np.concatenate([hull_areas, np.zeros(len(extra))]),
                       core_hull_planes=kpl, has_core=True)
        except Exception:   # qhull could not build the intersection: the geometry stays without a core
            pass
    setattr(shape, _CACHE_ATTR, out)
    _CACHE[key] = out
    return out


# --------------------------------------------------------------------------------------------------
# Inner balls: cheap "certainly colliding" certificates
# --------------------------------------------------------------------------------------------------
N_INNER_BALLS = 4


def inner_balls(planes, points, k=N_INNER_BALLS, rmin_ratio=0.4):
    """Up to ``k`` balls inside the convex set {x : n.x <= w} (``planes`` [m, 4] = nx ny nz w, unit normals;
    ``points``: its vertices), spread along its principal axis: for 31 stations along the axis the largest
    ball centred in the station's cross-section (one small LP each), then greedily the ball that reaches
    farthest beyond those already chosen.  Two shapes whose balls overlap certainly collide -- the item setup
    kernel tests that before it spends a GJK run on the pair (it settles ~60 % of the colliding Panda items).
    Returns float32 [k, 4] (cx cy cz r; r = 0: unused), radii re-measured at the float32 centres and rounded
    inwards."""
    from scipy.optimize import linprog

    out = np.zeros((k, 4), dtype=np.float32)
    planes = np.asarray(planes, dtype=np.float64)
    pts = np.asarray(points, dtype=np.float64)
    if len(planes) < 4 or len(pts) < 4:
        return out
    c0 = pts.mean(axis=0)
    _, _, vt = np.linalg.svd(pts - c0)
    ax = vt[0]
    ts = (pts - c0) @ ax
    n, w = planes[:, :3], planes[:, 3]
    A = np.hstack([n, np.ones((len(n), 1))])
    cand = []
    for t in np.linspace(ts.min(), ts.max(), 33)[1:-1]:
        res = linprog(c=[0, 0, 0, -1.0], A_ub=A, b_ub=w, A_eq=[[*ax, 0.0]], b_eq=[t + c0 @ ax], bounds=[(None, None)] * 3 + [(0, None)], method="highs")
        if res.success and res.x[3] > 1e-4:
            cand.append((res.x[:3], float(res.x[3])))
    if not cand:
        return out
    rmax = max(r for _, r in cand)
    cand = [x for x in cand if x[1] >= rmin_ratio * rmax]
    picked = []
    while cand and len(picked) < k:
        def reach(cr):
            c, r = cr
            return r if not picked else min(r - (ri - np.linalg.norm(ci - c)) for ci, ri in picked)

        j = int(np.argmax([reach(x) for x in cand]))
        if picked and reach(cand[j]) < 2e-3:
            break
        picked.append(cand.pop(j))
    for i, (c, _) in enumerate(picked):
        c32 = c.astype(np.float32)
        r = float((w - n @ c32.astype(np.float64)).min()) - 2e-6
        if r > 0.0:
            r32 = np.float32(r)
            out[i] = [*c32, r32 if float(r32) <= r else np.nextafter(r32, np.float32(0.0))]
    return out
