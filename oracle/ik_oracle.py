"""Float64 numpy restatement of one try of the reference's differential IK.

TEST INFRASTRUCTURE ONLY (see the header of ``pbp_oracle.c``).

Follows /root/reference/src/pyroboplan/ik/differential_ik.py:208-298 line by line; the Pinocchio
calls are restated from their published definitions:
  * ``framesForwardKinematics`` -> ``frame_pose`` (forward product of joint transforms),
  * ``computeFrameJacobian(..., LOCAL)`` -> ``frame_jacobian_local``: world-frame geometric Jacobian
    columns [z x (p_f - p_j); z] (revolute) / [z; 0] (prismatic) rotated into the frame,
  * ``pinocchio.log`` (SE(3) logarithm, Motion ordered [linear; angular]) -> ``log6``.
The GPU kernel derives FK and Jacobian differently (one backward sweep), so agreement is a check.
"""

import numpy as np


def _joint_motion(joint, q):
    T = np.eye(4)
    a = joint.axis
    if joint.type == 1:  # revolute
        c, s = np.cos(q), np.sin(q)
        K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
        T[:3, :3] = np.eye(3) + s * K + (1 - c) * (K @ K)
    else:
        T[:3, 3] = a * q
    return T


def chain(model, frame_id):
    j = model.frames[frame_id].parentJoint
    out = []
    while j > 0:
        out.append(j)
        j = model.parents[j]
    return out[::-1]


def frame_pose(model, frame_id, q):
    """World pose of the frame and of every joint on its chain (4x4 matrices)."""
    T = np.eye(4)
    joint_T = {}
    for j in chain(model, frame_id):
        T = T @ model.jointPlacements[j].homogeneous @ _joint_motion(model.joints[j], q[model.joints[j].idx_q])
        joint_T[j] = T.copy()
    return T @ model.frames[frame_id].placement.homogeneous, joint_T


def frame_jacobian_local(model, frame_id, q):
    """6 x nv Jacobian of the frame twist expressed in the frame (rows: linear, angular)."""
    oMf, joint_T = frame_pose(model, frame_id, q)
    J = np.zeros((6, model.nv))
    Rf, pf = oMf[:3, :3], oMf[:3, 3]
    for j, T in joint_T.items():
        z = T[:3, :3] @ model.joints[j].axis
        col = model.joints[j].idx_v
        if model.joints[j].type == 1:
            J[:3, col] = Rf.T @ np.cross(z, pf - T[:3, 3])
            J[3:, col] = Rf.T @ z
        else:
            J[:3, col] = Rf.T @ z
    return J


def log3(R):
    v = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
    s = 0.5 * np.linalg.norm(v)
    c = 0.5 * (np.trace(R) - 1.0)
    theta = np.arctan2(s, c)
    if c > -0.99:
        k = theta / (2.0 * s) if theta > 1e-8 else 0.5 * (1.0 + theta * theta / 6.0)
        return k * v, theta
    d = np.diag(R)
    kx = int(np.argmax(d))
    ax = np.zeros(3)
    ax[kx] = np.sqrt(max((d[kx] - c) / (1.0 - c), 0.0))
    for o in (1, 2):
        m = (kx + o) % 3
        ax[m] = (R[m, kx] + R[kx, m]) / (2.0 * ax[kx] * (1.0 - c))
    if ax @ v < 0:
        ax = -ax
    return theta * ax / np.linalg.norm(ax), theta


def log6(M):
    """se(3) logarithm of a 4x4 transform as [linear; angular]."""
    w, theta = log3(M[:3, :3])
    p = M[:3, 3]
    t2 = theta * theta
    if theta < 1e-4:
        alpha = 1.0 - t2 / 12.0 - t2 * t2 / 720.0
        beta = 1.0 / 12.0 + t2 / 720.0
    else:
        st, ct = np.sin(theta), np.cos(theta)
        alpha = theta * st / (2.0 * (1.0 - ct))
        beta = 1.0 / t2 - st / (2.0 * theta * (1.0 - ct))
    v = alpha * p - 0.5 * np.cross(w, p) + beta * (w @ p) * w
    return np.concatenate([v, w])


def ik_try(model, frame_id, target, q0, e0=None, max_iters=200, max_translation_error=1e-3, max_rotation_error=1e-3,
           damping=1e-3, min_step_size=0.1, max_step_size=0.5, active=None, joint_weights=None, nullspace_components=()):
    """One try (inner while loop of the reference).  target: 4x4.  Returns (q, converged, within_limits, iters, e0).
    nullspace_components: callables ``comp(model, q)`` as in the reference (:274-283)."""
    q = np.array(q0, dtype=np.float64)
    active = list(range(model.nq)) if active is None else list(active)
    W = np.eye(len(active)) if joint_weights is None else np.linalg.inv(np.diag(np.asarray(joint_weights, dtype=np.float64)))
    Tinv = np.linalg.inv(target)
    converged = False
    n_iters = 0
    while n_iters < max_iters:
        cur, _ = frame_pose(model, frame_id, q)
        error = -log6(Tinv @ cur)
        if np.linalg.norm(error[:3]) < max_translation_error and np.linalg.norm(error[3:]) < max_rotation_error:
            q = (q + np.pi) % (2 * np.pi) - np.pi
            converged = True
            break
        J = frame_jacobian_local(model, frame_id, q)[:, active]
        jjt = J @ W @ J.T + damping**2 * np.eye(6)
        error_norm = np.linalg.norm(error)
        if not e0:
            e0 = error_norm
        alpha = min_step_size + (1.0 - error_norm / e0) * (max_step_size - min_step_size)
        if not nullspace_components:
            q_step = alpha * W @ J.T @ np.linalg.solve(jjt, error)
        else:
            ns = sum(comp(model, q)[active] for comp in nullspace_components)
            q_step = alpha * (W @ J.T @ np.linalg.solve(jjt, error - J @ ns) + ns)
        for dq, idx in zip(q_step, active):
            q[idx] += dq
        n_iters += 1
        if np.any(np.isinf(q)):
            break
    within = bool(np.all(q >= model.lowerPositionLimit) and np.all(q <= model.upperPositionLimit)) if converged else False
    return q, converged, within, n_iters, e0


# ------------------------------------------------------------------------------------------------
# collision-avoidance null-space term (reference ik/nullspace_components.py:82-142 with
# core/utils.py:478-538), float64, from the C oracle's per-pair distances and witness points
# ------------------------------------------------------------------------------------------------
def point_jacobian(model, joint_T, parent_joint, point):
    """3 x nv Jacobian of the linear velocity of a world point rigidly attached to `parent_joint`
    (= [I, -skew(r)] @ LOCAL_WORLD_ALIGNED frame Jacobian of the reference)."""
    J = np.zeros((3, model.nv))
    j = parent_joint
    while j > 0:
        T = joint_T[j]
        z = T[:3, :3] @ model.joints[j].axis
        J[:, model.joints[j].idx_v] = np.cross(z, point - T[:3, 3]) if model.joints[j].type == 1 else z
        j = model.parents[j]
    return J


def all_joint_poses(model, q):
    T = {0: np.eye(4)}
    for j in range(1, model.njoints):
        T[j] = T[model.parents[j]] @ model.jointPlacements[j].homogeneous @ _joint_motion(model.joints[j], q[model.joints[j].idx_q])
    return T


def collision_vector_and_jacobians(model, collision_model, orc, pair_idx, q):
    """(normal * signed distance, Jcoll1, Jcoll2, signed distance, exact) for one pair -- coal's
    convention: negative distance = penetration depth, normal from geometry 1 to geometry 2."""
    cp = collision_model.collisionPairs[pair_idx]
    d, pa, pb, normal, exact = orc.pair_signed_distance(q, cp.first, cp.second)
    joint_T = all_joint_poses(model, q)
    J1 = point_jacobian(model, joint_T, collision_model.geometryObjects[cp.first].parentJoint, pa)
    J2 = point_jacobian(model, joint_T, collision_model.geometryObjects[cp.second].parentJoint, pb)
    return normal * d, J1, J2, d, exact


def collision_avoidance_component(model, collision_model, orc, q, dist_padding=0.05, gain=1.0):
    """Returns (component [nv], every contributing pair exact?)."""
    comp = np.zeros(model.nv)
    all_exact = True
    for k in range(len(collision_model.collisionPairs)):
        vec, J1, J2, d, exact = collision_vector_and_jacobians(model, collision_model, orc, k, q)
        if d > dist_padding:
            continue
        all_exact &= exact
        comp -= (vec - dist_padding) @ (J2 - J1)
    return gain * comp, all_exact
