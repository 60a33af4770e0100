"""Franka Emika Panda example model -- mirror of /root/reference/src/pyroboplan/models/panda.py.

Same functions, same argument meaning; models are this package's containers instead of
Pinocchio objects, obstacles are this package's ``Box`` / ``Sphere`` instead of ``coal`` ones.
"""

import numpy as np

from ..core.utils import set_collisions
from ..model import Box, GeometryObject, SE3, Sphere, build_models_from_description, removeCollisionPairs
from .utils import load_description


def load_models(use_sphere_collisions=False):
    """(model, collision_model, visual_model) of the Panda; cf. reference models/panda.py:13-27."""
    return build_models_from_description(
        load_description("panda_spheres" if use_sphere_collisions else "panda")
    )


def add_self_collisions(model, collision_model, srdf_filename=None):
    """All pairs minus the SRDF-disabled ones; cf. reference models/panda.py:57-79."""
    collision_model.addAllCollisionPairs()
    if srdf_filename is None:
        disabled = [tuple(p) for p in load_description("panda")["disabled_collisions"]]
        removeCollisionPairs(model, collision_model, disabled)
    else:
        removeCollisionPairs(model, collision_model, srdf_filename)


def add_object_collisions(model, collision_model, visual_model, inflation_radius=0.0):
    """Ground plane, two spheres, two boxes and their pairs; cf. reference models/panda.py:82-172."""
    r = inflation_radius
    objects = [
        ("ground_plane", [0.0, 0.0, -0.151], Box(2.0, 2.0, 0.3), [0.5, 0.5, 0.5, 0.5]),
        ("obstacle_sphere_1", [0.0, 0.1, 1.1], Sphere(0.2 + r), [0.0, 1.0, 0.0, 0.5]),
        ("obstacle_sphere_2", [0.5, 0.5, 0.5], Sphere(0.25 + r), [1.0, 1.0, 0.0, 0.5]),
        ("obstacle_box_1", [-0.5, 0.5, 0.7], Box(0.25 + 2.0 * r, 0.55 + 2.0 * r, 0.55 + 2.0 * r), [1.0, 0.0, 0.0, 0.5]),
        ("obstacle_box_2", [-0.5, -0.5, 0.75], Box(0.33 + 2.0 * r, 0.33 + 2.0 * r, 0.33 + 2.0 * r), [0.0, 0.0, 1.0, 0.5]),
    ]
    for name, xyz, shape, color in objects:
        obj = GeometryObject(name, 0, SE3(np.eye(3), np.array(xyz)), shape)
        obj.meshColor = np.array(color)
        visual_model.addGeometryObject(obj)
        collision_model.addGeometryObject(obj)

    collision_names = [c.name for c in collision_model.geometryObjects if "panda" in c.name]
    obstacle_names = [
        "ground_plane",
        "obstacle_box_1",
        "obstacle_box_2",
        "obstacle_sphere_1",
        "obstacle_sphere_2",
    ]
    for obstacle_name in obstacle_names:
        for collision_name in collision_names:
            set_collisions(model, collision_model, obstacle_name, collision_name, True)

    # The base link rests on the ground.
    set_collisions(model, collision_model, "panda_link0", "ground_plane", False)
