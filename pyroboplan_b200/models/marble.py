"""Marble-labyrinth example (a sphere sliding in the plane between box walls) -- mirror of
/root/reference/src/pyroboplan/models/marble.py."""

import numpy as np

from ..core.utils import set_collisions
from ..model import Box, GeometryObject, SE3, build_models_from_description
from .utils import load_description


def load_models():
    """(model, collision_model, visual_model); cf. reference models/marble.py:12-25."""
    return build_models_from_description(load_description("marble"))


# (x, y, width, height) of the boxes standing on the plane, reference models/marble.py:55-71
_WALLS = [(-0.05, -0.05, 0.05, 1.1), (2.0, -0.05, 0.05, 1.1), (-0.05, -0.05, 2.1, 0.05), (-0.05, 1.0, 2.1, 0.05)]
_OBSTACLES = [(0.3, 0.0, 0.05, 0.7), (0.75, 0.2, 0.05, 0.8), (1.3, 0.0, 0.05, 0.3), (0.8, 0.5, 0.5, 0.05),
              (1.6, 0.45, 0.05, 0.4), (1.65, 0.45, 0.35, 0.05)]


def add_object_collisions(model, collision_model, visual_model):
    """Ground plate + ten boxes; every box is paired with the marble ("body"); cf. reference
    models/marble.py:28-78 (the ground plate gets no pair there either)."""
    def add(name, xyz, shape, color):
        obj = GeometryObject(name, 0, SE3(np.eye(3), np.array(xyz)), shape)
        obj.meshColor = np.array(color)
        visual_model.addGeometryObject(obj)
        collision_model.addGeometryObject(obj)

    add("ground", [1.0, 0.5, -0.05 - 0.02], Box(2.1, 1.1, 0.1), [0.8, 0.8, 0.8, 1.0])
    for k, (x, y, w, h) in enumerate(_WALLS + _OBSTACLES, start=1):
        add(f"box{k}", [x + w / 2, y + h / 2, 0.0], Box(w, h, 0.04), [0.3, 0.3, 0.3, 1.0])
    for obj in list(collision_model.geometryObjects):
        if obj.name.startswith("box"):
            set_collisions(model, collision_model, obj.name, "body", True)
