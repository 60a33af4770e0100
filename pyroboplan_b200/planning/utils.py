"""Drop-in for the edge-validation helpers of ``pyroboplan.planning.utils``.

Reference: /root/reference/src/pyroboplan/planning/utils.py:15-140.
"""

import numpy as np

from ..core.utils import check_collisions_at_state, configuration_distance, _engine


def extend_robot_state(q_parent, q_sample, max_connection_distance):
    """Step from ``q_parent`` toward ``q_sample`` by at most ``max_connection_distance``.

    Reference: planning/utils.py:15-48.  Returns ``q_sample`` itself (same object) when it is
    within reach, because planners compare nodes with ``np.array_equal``.
    """
    q_diff = q_sample - q_parent
    distance = np.linalg.norm(q_diff)
    if distance == 0.0:
        return q_sample
    q_increment = max_connection_distance * q_diff / distance
    if configuration_distance(q_parent, q_sample) > max_connection_distance:
        return q_parent + q_increment
    return q_sample


def discretize_joint_space_path(q_path, max_angle_distance):
    """Waypoints -> samples, ``ceil(||dq|| / step) + 1`` per segment, both ends included.

    Reference: planning/utils.py:114-140.  Host float64 (list in, list of float64 arrays out,
    bit-identical to the reference, so planners can keep storing the result); the device
    discretisation used for collision checking is ``has_collision_free_path(s)`` /
    ``CollisionEngine.discretize``.
    """
    q_discretized = []
    for idx in range(1, len(q_path)):
        q_start = q_path[idx - 1]
        q_end = q_path[idx]
        q_diff = q_end - q_start
        num_steps = int(np.ceil(np.linalg.norm(q_diff) / max_angle_distance)) + 1
        step_vec = np.linspace(0.0, 1.0, num_steps)
        q_discretized.extend([q_start + step * q_diff for step in step_vec])
    return q_discretized


def has_collision_free_path(
    q1, q2, max_step_size, model, collision_model, data=None, collision_data=None, distance_padding=0.0
):
    """True if the straight joint-space segment q1 -> q2 is collision free.

    Reference: planning/utils.py:51-111 (checks q2, then every sample of the discretised
    segment).  Here one device call discretises and tests the segment.
    """
    if len(collision_model.collisionPairs) == 0:
        return True
    hit = _engine(model, collision_model).check_edges(
        np.asarray(q1, dtype=np.float64).reshape(1, -1),
        np.asarray(q2, dtype=np.float64).reshape(1, -1),
        max_step_size,
        distance_padding,
    )
    return not bool(hit[0])


def has_collision_free_paths(Q1, Q2, max_step_size, model, collision_model, distance_padding=0.0):
    """Batched ``has_collision_free_path``: edges Q1[e] -> Q2[e] (numpy or CUDA tensors) -> bool [E]."""
    if len(collision_model.collisionPairs) == 0:
        try:
            import torch

            if isinstance(Q1, torch.Tensor):
                return torch.ones(Q1.reshape(-1, model.nq).shape[0], dtype=torch.bool, device=Q1.device)
        except ImportError:
            pass
        return np.ones(np.asarray(Q1).reshape(-1, model.nq).shape[0], dtype=bool)
    hit = _engine(model, collision_model).check_edges(Q1, Q2, max_step_size, distance_padding)
    return ~hit
