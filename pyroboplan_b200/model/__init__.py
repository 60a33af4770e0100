"""Host-side model containers and URDF/SRDF/STL ingestion (no Pinocchio needed)."""

from .se3 import SE3, rpy_to_matrix, axis_angle_to_matrix
from .shapes import (
    Box,
    Capsule,
    CollisionGeometry,
    Convex,
    Cylinder,
    Sphere,
    SHAPE_BOX,
    SHAPE_CAPSULE,
    SHAPE_CONVEX,
    SHAPE_CYLINDER,
    SHAPE_SPHERE,
)
from .robot_model import (
    CollisionPair,
    Data,
    Frame,
    GeometryData,
    GeometryModel,
    GeometryObject,
    Model,
    JOINT_FIXED,
    JOINT_PRISMATIC,
    JOINT_REVOLUTE,
    FRAME_BODY,
    FRAME_FIXED_JOINT,
    FRAME_JOINT,
    FRAME_OP,
)
from .urdf import (
    buildModelsFromUrdf,
    build_models_from_description,
    read_stl,
    removeCollisionPairs,
    srdf_disabled_pairs,
    urdf_to_description,
)
