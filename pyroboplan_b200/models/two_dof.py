"""2-DOF planar arm example -- mirror of /root/reference/src/pyroboplan/models/two_dof.py."""

import numpy as np

from ..core.utils import set_collisions
from ..model import Box, Cylinder, GeometryObject, SE3, build_models_from_description
from .utils import load_description


def load_models():
    """(model, collision_model, visual_model); cf. reference models/two_dof.py:12-25."""
    return build_models_from_description(load_description("two_dof"))


def add_object_collisions(model, collision_model, visual_model):
    """Cylinder + box obstacles and their pairs; cf. reference models/two_dof.py:28-68."""
    objects = [
        ("obstacle_1", [1.0, 1.0, 0.0], Cylinder(0.3, 0.1), [0.0, 1.0, 0.0, 0.5]),
        ("obstacle_2", [-1.0, -0.75, 0.0], Box(0.5, 0.5, 0.1), [1.0, 0.0, 0.0, 0.5]),
    ]
    for name, xyz, shape, color in objects:
        obj = GeometryObject(name, 0, SE3(np.eye(3), np.array(xyz)), shape)
        obj.meshColor = np.array(color)
        visual_model.addGeometryObject(obj)
        collision_model.addGeometryObject(obj)

    collision_names = [c.name for c in collision_model.geometryObjects if "arm" in c.name]
    for obstacle_name in ["obstacle_1", "obstacle_2"]:
        for collision_name in collision_names:
            set_collisions(model, collision_model, obstacle_name, collision_name, True)
