"""Null-space components for differential IK -- mirror of
/root/reference/src/pyroboplan/ik/nullspace_components.py, each with a batched sibling.

The single-state functions keep the reference's signatures (a numpy vector in, a numpy vector out).
The ``*_batch`` functions take a CUDA tensor Q [N, nq] and return a tensor [N, nq]; they are what
``DifferentialIk.solve_batch`` calls once per iteration.
"""

import numpy as np


def zero_nullspace_component(model, q):
    """Reference: nullspace_components.py:9-24."""
    return np.zeros_like(model.lowerPositionLimit)


def joint_limit_nullspace_component(model, q, gain=1.0, padding=0.0):
    """Pushes coordinates beyond the (padded) limits back. Reference: nullspace_components.py:27-57."""
    upper = model.upperPositionLimit - padding
    lower = model.lowerPositionLimit + padding
    q = np.asarray(q, dtype=np.float64)
    grad = zero_nullspace_component(model, q)
    over, under = q > upper, q < lower
    grad[over] = -gain * (q[over] - upper[over])
    grad[under & ~over] = -gain * (q[under & ~over] - lower[under & ~over])
    return grad


def joint_center_nullspace_component(model, q, gain=1.0):
    """Pulls towards the middle of the joint ranges. Reference: nullspace_components.py:60-79."""
    return gain * (0.5 * (model.lowerPositionLimit + model.upperPositionLimit) - q)


def collision_avoidance_nullspace_component(model, data, collision_model, collision_data, q, dist_padding=0.05, gain=1.0):
    """Pushes the joints away from every pair closer than ``dist_padding``.

    Reference: nullspace_components.py:82-142.  ``data`` / ``collision_data`` are ignored; distances,
    witness points and the gradient come from the device (``pbp_collision_nullspace``, FP32).
    """
    from ..engine import get_engine

    if len(collision_model.collisionPairs) == 0:
        return np.zeros_like(model.lowerPositionLimit)
    out = get_engine(model, collision_model).collision_nullspace(np.asarray(q, dtype=np.float64).reshape(1, -1), dist_padding, gain)
    return out[0].astype(np.float64)


# ----------------------------------------------------------------------------------------------
# batched siblings (CUDA tensors)
# ----------------------------------------------------------------------------------------------
def zero_nullspace_component_batch(model, Q):
    import torch

    return torch.zeros_like(Q)


def joint_limit_nullspace_component_batch(model, Q, gain=1.0, padding=0.0):
    import torch

    upper = torch.as_tensor(model.upperPositionLimit - padding, dtype=Q.dtype, device=Q.device)
    lower = torch.as_tensor(model.lowerPositionLimit + padding, dtype=Q.dtype, device=Q.device)
    return -gain * (torch.clamp(Q - upper, min=0.0) + torch.clamp(Q - lower, max=0.0))


def joint_center_nullspace_component_batch(model, Q, gain=1.0):
    import torch

    centre = torch.as_tensor(0.5 * (model.lowerPositionLimit + model.upperPositionLimit), dtype=Q.dtype, device=Q.device)
    return gain * (centre - Q)


def collision_avoidance_nullspace_component_batch(model, collision_model, Q, dist_padding=0.05, gain=1.0):
    import torch

    from ..engine import get_engine

    if len(collision_model.collisionPairs) == 0:
        return torch.zeros_like(Q)
    return get_engine(model, collision_model).collision_nullspace(Q.to(torch.float32), dist_padding, gain).to(Q.dtype)
