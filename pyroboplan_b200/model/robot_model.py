"""Host-side kinematic / geometry model containers.

These mirror the slice of ``pinocchio.Model`` / ``pinocchio.GeometryModel`` that the
reference's hot path reads, so the drop-in wrappers in ``pyroboplan_b200.core.utils`` can keep
the reference's signatures (``model, collision_model, q, data, collision_data``):

* ``model.lowerPositionLimit / upperPositionLimit / nq / nv / nframes / getFrameId / createData``
  -- /root/reference/src/pyroboplan/core/utils.py:190-192,251-255,318-319,402-408
* ``collision_model.collisionPairs / geometryObjects / ngeoms / getGeometryId /
  addCollisionPair / removeCollisionPair / addAllCollisionPairs / addGeometryObject /
  createData`` -- /root/reference/src/pyroboplan/core/utils.py:379-475,
  /root/reference/src/pyroboplan/models/panda.py:78-79,105-172

Conventions restated from Pinocchio 3.7: joint 0 is ``universe``;
fixed URDF joints are merged into the parent joint's body and survive only as frames;
``getFrameId`` / ``getGeometryId`` return ``nframes`` / ``ngeoms`` when the name is unknown;
``addAllCollisionPairs`` adds every (i < j) whose parent joints differ.
"""

import numpy as np

from .se3 import SE3

# Joint type codes shared with the CUDA engine (include/pbp_engine.h: pbp_joint_type).
JOINT_FIXED = 0  # only joint 0 (universe)
JOINT_REVOLUTE = 1  # rotation about `axis` by q
JOINT_PRISMATIC = 2  # translation along `axis` by q

# Frame types (subset of pinocchio.FrameType)
FRAME_OP = 1
FRAME_JOINT = 2
FRAME_FIXED_JOINT = 4
FRAME_BODY = 8


class Frame:
    """Named frame rigidly attached to a joint (``pinocchio.Frame`` subset)."""

    def __init__(self, name, parent_joint, placement, frame_type, parent_frame=0):
        self.name = name
        self.parentJoint = int(parent_joint)
        self.parent = self.parentJoint  # pinocchio 2 alias
        self.parentFrame = int(parent_frame)
        self.placement = placement
        self.type = frame_type


class JointModel:
    def __init__(self, joint_type, axis, idx_q):
        self.type = joint_type
        self.axis = np.asarray(axis, dtype=np.float64)
        self.idx_q = idx_q
        self.idx_v = idx_q
        self.nq = 0 if joint_type == JOINT_FIXED else 1
        self.nv = self.nq


class Data:
    """Scratch mirroring ``pinocchio.Data``: joint and frame placements of the last FK."""

    def __init__(self, model):
        self.oMi = [SE3.Identity() for _ in range(model.njoints)]
        self.oMf = [SE3.Identity() for _ in range(model.nframes)]

    def __bool__(self):
        return True


class Model:
    """Kinematic tree (``pinocchio.Model`` subset)."""

    def __init__(self, name="model"):
        self.name = name
        self.names = ["universe"]
        self.parents = [0]
        self.jointPlacements = [SE3.Identity()]
        self.joints = [JointModel(JOINT_FIXED, [0.0, 0.0, 0.0], -1)]
        self.frames = [Frame("universe", 0, SE3.Identity(), FRAME_FIXED_JOINT)]
        self.nq = 0
        self.nv = 0
        self.lowerPositionLimit = np.zeros(0)
        self.upperPositionLimit = np.zeros(0)
        self.velocityLimit = np.zeros(0)
        self.effortLimit = np.zeros(0)
        self._revision = 0

    # --- construction -------------------------------------------------------------
    def addJoint(
        self, parent, joint_type, axis, placement, name, lower, upper, velocity=np.inf, effort=np.inf
    ):
        if joint_type not in (JOINT_REVOLUTE, JOINT_PRISMATIC):
            raise ValueError("only revolute and prismatic joints are supported")
        axis = np.asarray(axis, dtype=np.float64)
        norm = np.linalg.norm(axis)
        if norm == 0.0:
            raise ValueError(f"joint {name}: zero axis")
        self.names.append(name)
        self.parents.append(int(parent))
        self.jointPlacements.append(placement)
        self.joints.append(JointModel(joint_type, axis / norm, self.nq))
        self.nq += 1
        self.nv += 1
        self.lowerPositionLimit = np.append(self.lowerPositionLimit, float(lower))
        self.upperPositionLimit = np.append(self.upperPositionLimit, float(upper))
        self.velocityLimit = np.append(self.velocityLimit, float(velocity))
        self.effortLimit = np.append(self.effortLimit, float(effort))
        self._revision += 1
        return len(self.names) - 1

    def addFrame(self, frame):
        self.frames.append(frame)
        self._revision += 1
        return len(self.frames) - 1

    # --- queries ------------------------------------------------------------------
    @property
    def njoints(self):
        return len(self.names)

    @property
    def nframes(self):
        return len(self.frames)

    def getFrameId(self, name, frame_type=None):
        for i, f in enumerate(self.frames):
            if f.name == name and (frame_type is None or f.type & frame_type):
                return i
        return self.nframes

    def existFrame(self, name):
        return self.getFrameId(name) < self.nframes

    def getJointId(self, name):
        for i, n in enumerate(self.names):
            if n == name:
                return i
        return self.njoints

    def createData(self):
        return Data(self)


class CollisionPair:
    """Unordered pair of geometry indices; keeps (first, second) as given."""

    __slots__ = ("first", "second")

    def __init__(self, first, second):
        if first == second:
            raise ValueError("CollisionPair: first == second")
        self.first = int(first)
        self.second = int(second)

    def __eq__(self, other):
        return (self.first == other.first and self.second == other.second) or (
            self.first == other.second and self.second == other.first
        )

    def __hash__(self):
        return hash(frozenset((self.first, self.second)))

    def __repr__(self):
        return f"CollisionPair({self.first}, {self.second})"


class GeometryObject:
    """``pinocchio.GeometryObject`` subset.

    Accepts the two call forms the reference ecosystem uses:
    ``GeometryObject(name, parent_joint, placement, geometry)`` (as in
    /root/reference/src/pyroboplan/models/panda.py:98-103; ``parentFrame`` is then "unset")
    and ``GeometryObject(name, parent_joint, parent_frame, placement, geometry)``.
    """

    UNSET_FRAME = np.iinfo(np.int64).max

    def __init__(self, name, parent_joint, *args):
        if len(args) == 2:
            placement, geometry = args
            parent_frame = GeometryObject.UNSET_FRAME
        elif len(args) == 3:
            parent_frame, placement, geometry = args
        else:
            raise TypeError("GeometryObject(name, parent_joint, [parent_frame,] placement, geometry)")
        self.name = name
        self.parentJoint = int(parent_joint)
        self.parentFrame = int(parent_frame)
        self.placement = placement
        self.geometry = geometry
        self.meshColor = np.array([0.5, 0.5, 0.5, 1.0])
        self.meshPath = getattr(geometry, "source", None) or ""
        self.meshScale = np.ones(3)
        self.disableCollision = False


class _Request:
    def __init__(self):
        self.security_margin = 0.0


class _CollisionResult:
    def __init__(self):
        self._hit = False

    def isCollision(self):
        return self._hit


class _DistanceResult:
    def __init__(self):
        self.min_distance = np.inf
        self.normal = np.zeros(3)
        self._p1 = np.zeros(3)
        self._p2 = np.zeros(3)

    def getNearestPoint1(self):
        return self._p1

    def getNearestPoint2(self):
        return self._p2


class GeometryData:
    """Scratch mirroring ``pinocchio.GeometryData`` (placements, per-pair requests/results)."""

    def __init__(self, geometry_model):
        n = len(geometry_model.collisionPairs)
        self.oMg = [SE3.Identity() for _ in range(geometry_model.ngeoms)]
        self.collisionRequests = [_Request() for _ in range(n)]
        self.collisionResults = [_CollisionResult() for _ in range(n)]
        self.distanceResults = [_DistanceResult() for _ in range(n)]

    def __bool__(self):
        return True


class GeometryModel:
    """Collision geometry set + active pair list (``pinocchio.GeometryModel`` subset)."""

    def __init__(self):
        self.geometryObjects = []
        self.collisionPairs = []
        self._revision = 0

    @property
    def ngeoms(self):
        return len(self.geometryObjects)

    def addGeometryObject(self, obj):
        for existing in self.geometryObjects:
            if existing.name == obj.name:
                raise ValueError(f"geometry {obj.name!r} already exists")
        self.geometryObjects.append(obj)
        self._revision += 1
        return len(self.geometryObjects) - 1

    def getGeometryId(self, name):
        for i, g in enumerate(self.geometryObjects):
            if g.name == name:
                return i
        return self.ngeoms

    def existGeometryName(self, name):
        return self.getGeometryId(name) < self.ngeoms

    def existCollisionPair(self, pair):
        return pair in self.collisionPairs

    def addCollisionPair(self, pair):
        if pair.first >= self.ngeoms or pair.second >= self.ngeoms:
            raise ValueError("CollisionPair index out of range")
        if not self.existCollisionPair(pair):
            self.collisionPairs.append(pair)
            self._revision += 1

    def removeCollisionPair(self, pair):
        if pair in self.collisionPairs:
            self.collisionPairs.remove(pair)
            self._revision += 1

    def removeAllCollisionPairs(self):
        self.collisionPairs = []
        self._revision += 1

    def addAllCollisionPairs(self):
        self.removeAllCollisionPairs()
        for i in range(self.ngeoms):
            for j in range(i + 1, self.ngeoms):
                if self.geometryObjects[i].parentJoint != self.geometryObjects[j].parentJoint:
                    self.collisionPairs.append(CollisionPair(i, j))
        self._revision += 1

    def createData(self):
        return GeometryData(self)
