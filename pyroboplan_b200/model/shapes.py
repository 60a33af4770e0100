"""Collision shapes (host side), mirroring the ``coal`` constructors the reference calls.

Reference call sites: ``coal.Box(x, y, z)`` (full side lengths), ``coal.Sphere(r)``,
``coal.Cylinder(radius, length)`` (centred, axis = local z) at
/root/reference/src/pyroboplan/models/panda.py:102,112,122,132-136,146-150 and
/root/reference/src/pyroboplan/models/two_dof.py:45,55.  URDF ``<mesh>`` collision geometry is
represented by :class:`Convex` (the convex hull of the mesh vertices); see DESIGN.md for why
this matches the reference's triangle-soup semantics on the bundled Panda meshes.
"""

import numpy as np

# Shape type codes shared with the CUDA engine (include/pbp_engine.h: pbp_shape_type).
SHAPE_SPHERE = 0
SHAPE_BOX = 1
SHAPE_CYLINDER = 2
SHAPE_CAPSULE = 3
SHAPE_CONVEX = 4


class CollisionGeometry:
    shape_type = -1

    def bounding_sphere(self):
        """(center[3], radius) of a sphere enclosing the shape, in the shape frame."""
        raise NotImplementedError

    def inner_sphere(self):
        """(center[3], radius) of a sphere contained in the shape, in the shape frame."""
        raise NotImplementedError


class Sphere(CollisionGeometry):
    shape_type = SHAPE_SPHERE

    def __init__(self, radius):
        self.radius = float(radius)

    def bounding_sphere(self):
        return np.zeros(3), self.radius

    def inner_sphere(self):
        return np.zeros(3), self.radius


class Box(CollisionGeometry):
    shape_type = SHAPE_BOX

    def __init__(self, x, y=None, z=None):
        if y is None:
            x, y, z = np.asarray(x, dtype=np.float64).reshape(3)
        self.halfSide = 0.5 * np.array([x, y, z], dtype=np.float64)

    def bounding_sphere(self):
        return np.zeros(3), float(np.linalg.norm(self.halfSide))

    def inner_sphere(self):
        return np.zeros(3), float(np.min(self.halfSide))


class Cylinder(CollisionGeometry):
    shape_type = SHAPE_CYLINDER

    def __init__(self, radius, lz):
        self.radius = float(radius)
        self.halfLength = 0.5 * float(lz)

    def bounding_sphere(self):
        return np.zeros(3), float(np.hypot(self.radius, self.halfLength))

    def inner_sphere(self):
        return np.zeros(3), float(min(self.radius, self.halfLength))


class Capsule(CollisionGeometry):
    shape_type = SHAPE_CAPSULE

    def __init__(self, radius, lz):
        self.radius = float(radius)
        self.halfLength = 0.5 * float(lz)

    def bounding_sphere(self):
        return np.zeros(3), self.radius + self.halfLength

    def inner_sphere(self):
        return np.zeros(3), self.radius


class Convex(CollisionGeometry):
    """Convex polytope given by a point cloud; only the hull vertices are kept.

    ``points`` : array [n, 3].  ``source`` records where the points came from (mesh path).
    ``volume_ratio`` = (signed volume of the source triangle mesh) / (hull volume) when the
    shape was built from a closed mesh, else None; 1.0 means the mesh itself is convex.
    """

    shape_type = SHAPE_CONVEX

    def __init__(self, points, source=None, volume_ratio=None, triangles=None):
        from scipy.spatial import ConvexHull

        pts = np.asarray(points, dtype=np.float64).reshape(-1, 3)
        # The source triangles (if any) are kept verbatim: the oracle evaluates the reference's surface
        # semantics on them; the engine consumes the hull (``self.points``, ``self.planes``) and only
        # needs to know THAT the geometry is a surface.
        self.mesh_vertices = pts.copy()
        self.mesh_triangles = None if triangles is None else np.asarray(triangles, dtype=np.int64).reshape(-1, 3)
        pts = np.unique(pts, axis=0)
        hull = ConvexHull(pts)
        keep = np.sort(hull.vertices)
        self.points = np.ascontiguousarray(pts[keep])
        # Facet planes n.x + d <= 0 inside (unit normals), de-duplicated (qhull triangulates).
        eq = hull.equations
        eq = eq[np.lexsort(np.round(eq, 9).T[::-1])]
        uniq = [eq[0]]
        for row in eq[1:]:
            if np.max(np.abs(row - uniq[-1])) > 1e-9:
                uniq.append(row)
        self.planes = np.array(uniq)
        # area of the hull face behind every plane (the engine sorts the planes by it: the enclosure
        # test of the surface semantics tries the largest faces first)
        tri = pts[hull.simplices]
        tri_area = 0.5 * np.linalg.norm(np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0]), axis=1)
        owner = np.abs(hull.equations[:, None, :] - self.planes[None, :, :]).sum(axis=2).argmin(axis=1)
        self.plane_areas = np.bincount(owner, weights=tri_area, minlength=len(self.planes))
        self.volume = float(hull.volume)
        self.source = source
        self.volume_ratio = volume_ratio
        self._bs = None
        self._is = None

    def bounding_sphere(self):
        if self._bs is None:
            self._bs = _min_enclosing_sphere(self.points)
        return self._bs

    def inner_sphere(self):
        if self._is is None:
            self._is = _chebyshev_ball(self.planes, self.points.mean(axis=0))
        return self._is


def _min_enclosing_sphere(pts, iters=2000):
    """Near-minimal enclosing sphere (Badoiu-Clarkson), radius padded to cover every point."""
    c = pts.mean(axis=0)
    for k in range(1, iters + 1):
        far = pts[np.argmax(np.einsum("ij,ij->i", pts - c, pts - c))]
        c = c + (far - c) / (k + 1.0)
    r = float(np.sqrt(np.max(np.einsum("ij,ij->i", pts - c, pts - c))))
    return c, r * (1.0 + 1e-9)


def _chebyshev_ball(planes, fallback_center):
    """Largest ball inside {x : n.x + d <= 0}: maximise r s.t. n.c + r <= -d."""
    from scipy.optimize import linprog

    n = planes[:, :3]
    d = planes[:, 3]
    A = np.hstack([n, np.ones((len(n), 1))])
    res = linprog(
        c=[0, 0, 0, -1.0],
        A_ub=A,
        b_ub=-d,
        bounds=[(None, None)] * 3 + [(0, None)],
        method="highs",
    )
    if not res.success:
        return fallback_center, 0.0
    return res.x[:3], float(res.x[3]) * (1.0 - 1e-9)
