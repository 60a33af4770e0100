"""Universal Robots UR5 on a base -- mirror of /root/reference/src/pyroboplan/models/ur5.py:9-47.

The UR collision meshes are NOT convex (upper arm and forearm fill 64-68 % of their hulls): they go through
the engine as surfaces without an inner core -- a hull hit proves nothing, the mesh triangles decide every
one of them (pbp_surface.cuh, PAIR_NOCORE).
"""

from ..model import build_models_from_description, removeCollisionPairs
from .utils import load_description


def load_ur5_on_base_models():
    """(model, collision_model, visual_model) of the UR5 on its base; cf. reference models/ur5.py:9-22."""
    return build_models_from_description(load_description("ur5_on_base"))


def add_ur5_on_base_self_collisions(model, collision_model, srdf_filename=None):
    """All pairs minus the SRDF-disabled ones; cf. reference models/ur5.py:25-47."""
    collision_model.addAllCollisionPairs()
    if srdf_filename is None:
        removeCollisionPairs(model, collision_model, [tuple(p) for p in load_description("ur5_on_base")["disabled_collisions"]])
    else:
        removeCollisionPairs(model, collision_model, srdf_filename)
