"""Drop-in for the collision hot path of ``pyroboplan.core.utils``.

Same names, arity and return conventions as /root/reference/src/pyroboplan/core/utils.py; the
arithmetic runs on the B200 through ``libpbp_engine.so`` (see ``pyroboplan_b200.engine``).
There is no CPU path: without the CUDA library or a CUDA device these functions raise.
"""

import numpy as np

from ..model import CollisionPair


# ----------------------------------------------------------------------------------------------
# Pair / geometry bookkeeping (integer work, host side) -- reference core/utils.py:379-475
# ----------------------------------------------------------------------------------------------
def get_collision_geometry_ids(model, collision_model, body):
    """Collision geometry ids for a geometry name or a frame (link) name.

    Reference: core/utils.py:379-413 (geometry name first, then frame -> ``parentFrame`` match).
    """
    ids = []
    gid = collision_model.getGeometryId(body)
    if gid < collision_model.ngeoms:
        ids.append(gid)
    else:
        fid = model.getFrameId(body)
        if fid < model.nframes:
            for i, obj in enumerate(collision_model.geometryObjects):
                if obj.parentFrame == fid:
                    ids.append(i)
    return ids


def get_collision_pair_indices_from_bodies(model, collision_model, body_list):
    """Indices of the active pairs touching any of the bodies. Reference: core/utils.py:416-446."""
    ids = set()
    for body in body_list:
        ids.update(get_collision_geometry_ids(model, collision_model, body))
    return [
        k
        for k, p in enumerate(collision_model.collisionPairs)
        if p.first in ids or p.second in ids
    ]


def _pair_type(collision_model):
    """``CollisionPair`` of the model's own kind: this package's, or ``pinocchio.CollisionPair`` when the
    caller holds a real ``pinocchio.GeometryModel`` (reference core/utils.py:471)."""
    from ..model import GeometryModel

    if isinstance(collision_model, GeometryModel):
        return CollisionPair
    import pinocchio

    return pinocchio.CollisionPair


def set_collisions(model, collision_model, body1, body2, enable):
    """Enable/disable all pairs between two bodies. Reference: core/utils.py:449-475."""
    for id1 in get_collision_geometry_ids(model, collision_model, body1):
        for id2 in get_collision_geometry_ids(model, collision_model, body2):
            pair = _pair_type(collision_model)(id1, id2)
            if enable:
                collision_model.addCollisionPair(pair)
            else:
                collision_model.removeCollisionPair(pair)


# ----------------------------------------------------------------------------------------------
# Collision queries -- single-state signatures of the reference, executed on the GPU engine
# ----------------------------------------------------------------------------------------------
def _engine(model, collision_model):
    from ..engine import get_engine

    return get_engine(model, collision_model)


def check_collisions_at_state(
    model, collision_model, q, data=None, collision_data=None, distance_padding=0.0
):
    """True if ``q`` is in collision (some active pair closer than ``distance_padding``).

    Reference: core/utils.py:8-51.  ``data`` / ``collision_data`` are accepted for signature
    compatibility and ignored (the engine owns its device workspaces); the padding is a
    per-call argument, not sticky state.  Zero active pairs -> False (``np.any([])``).
    """
    if len(collision_model.collisionPairs) == 0:
        return np.bool_(False)
    hit = _engine(model, collision_model).check_states(np.asarray(q, dtype=np.float64).reshape(1, -1), distance_padding)
    return np.bool_(hit[0])


def get_minimum_distance_at_state(model, collision_model, q, data=None, collision_data=None):
    """Minimum distance over the active pairs (``np.inf`` without pairs); negative when penetrating
    (the penetration depth, coal's convention).

    Reference: core/utils.py:54-87.  The batched hot path clamps distances at 0; a state it reports
    as touching / penetrating is re-evaluated pair by pair with the witness kernel, whose EPA gives
    the signed value.
    """
    if len(collision_model.collisionPairs) == 0:
        return np.inf
    eng = _engine(model, collision_model)
    q = np.asarray(q, dtype=np.float64).reshape(1, -1)
    _, dist = eng.check_states(q, 0.0, return_distance=True)
    if float(dist[0]) > 0.0:
        return float(dist[0])
    return float(eng.compute_distances(q, witness=False)[0].min())


def check_collisions_along_path(
    model, collision_model, q_path, data=None, collision_data=None, distance_padding=0.0
):
    """True if any configuration of the (already discretised) path collides.

    Reference: core/utils.py:90-132 (sequential with early return; same verdict).
    """
    if len(q_path) == 0 or len(collision_model.collisionPairs) == 0:
        return False
    Q = np.asarray(q_path, dtype=np.float64).reshape(len(q_path), -1)
    hit = _engine(model, collision_model).check_paths(Q, [0, len(Q)], distance_padding)
    return bool(hit[0])


def configuration_distance(q_start, q_end):
    """Euclidean joint-space distance. Reference: core/utils.py:135-151."""
    return np.linalg.norm(q_end - q_start)


def get_path_length(q_path):
    """Sum of segment distances. Reference: core/utils.py:154-171."""
    total_distance = 0.0
    for idx in range(1, len(q_path)):
        total_distance += configuration_distance(q_path[idx - 1], q_path[idx])
    return total_distance


def get_random_state(model, padding=0.0):
    """Uniform sample in the padded joint limits, drawn from the legacy global ``np.random``
    stream exactly as the reference does (core/utils.py:174-192), so seeded callers agree."""
    return np.random.uniform(model.lowerPositionLimit + padding, model.upperPositionLimit - padding)


def get_random_collision_free_state(
    model, collision_model, joint_padding=0.0, distance_padding=0.0, max_tries=100
):
    """Rejection sampling, one host draw + one device check per try.

    Reference: core/utils.py:195-229 (prints and returns None after ``max_tries`` failures).
    """
    num_tries = 0
    while num_tries < max_tries:
        state = get_random_state(model, padding=joint_padding)
        if not check_collisions_at_state(model, collision_model, state, distance_padding=distance_padding):
            return state
        num_tries += 1
    print(f"Could not generate collision-free state after {max_tries} tries.")
    return None


def check_within_limits(model, q):
    """Reference: core/utils.py:303-319."""
    return np.all(q >= model.lowerPositionLimit) and np.all(q <= model.upperPositionLimit)


_FK_HOLDERS = {}


def fk_holder(model):
    """An empty geometry model per kinematic model: the key the FK-only engine of ``model`` is cached on."""
    from ..model import GeometryModel

    holder = getattr(model, "_pbp_fk_holder", None)
    if holder is None:
        holder = _FK_HOLDERS.get(id(model))
    if holder is None:
        holder = GeometryModel()
        try:
            model._pbp_fk_holder = holder
        except (AttributeError, TypeError):
            _FK_HOLDERS[id(model)] = holder
    return holder


def extract_cartesian_pose(model, target_frame, q, data=None):
    """Pose of a model frame at ``q`` (device FK). Reference: core/utils.py:322-346."""
    return extract_cartesian_poses(model, target_frame, [q])[0]


def extract_cartesian_poses(model, target_frame, q_vec, data=None):
    """Poses of a model frame for a list of configurations. Reference: core/utils.py:349-376."""
    from ..engine import get_engine
    from ..model import GeometryModel, SE3

    fid = model.getFrameId(target_frame)
    if fid >= model.nframes:
        raise ValueError(f"unknown frame {target_frame!r}")
    eng = get_engine(model, fk_holder(model))
    Q = np.asarray(q_vec, dtype=np.float64).reshape(len(q_vec), -1)
    _, _, oMf = eng.fk(Q, joints=False, geoms=False, frames=True)
    return [SE3(oMf[i, fid, :9].reshape(3, 3).astype(np.float64), oMf[i, fid, 9:].astype(np.float64)) for i in range(len(Q))]


def calculate_collision_vector_and_jacobians(model, collision_model, data, collision_data, pair_idx, q):
    """Collision vector (normal * distance, object 1 -> object 2) and the 3 x nv Jacobians of the two
    witness points of pair ``pair_idx`` at ``q``.

    Reference: core/utils.py:478-538.  The reference reads ``collision_data.distanceResults`` that a
    previous ``computeDistances`` left behind; here ``data`` / ``collision_data`` are ignored and the
    pair's distance and witness points are computed on the device for ``q``.  The Jacobians --
    ``[I, -skew(r)] @ J_frame(LOCAL_WORLD_ALIGNED)``, i.e. column k = z_k x (p - o_k) for a revolute
    joint on the chain of the geometry's parent joint and z_k for a prismatic one -- are small
    host arithmetic on the device FK.  Overlapping pairs have a negative distance (penetration depth).
    """
    eng = _engine(model, collision_model)
    q = np.asarray(q, dtype=np.float64).reshape(1, -1)
    dist, p1, p2, normal = eng.compute_distances(q)
    oMi, _, _ = eng.fk(q, joints=True, geoms=False, frames=False)
    hm, hcm = eng.host_model, eng.host_collision_model     # this package's containers (converted from Pinocchio's if need be)
    cp = hcm.collisionPairs[pair_idx]
    jacs = []
    for gid, point in ((cp.first, p1[0, pair_idx]), (cp.second, p2[0, pair_idx])):
        J = np.zeros((3, hm.nv))
        j = hcm.geometryObjects[gid].parentJoint
        while j > 0:
            R, o = oMi[0, j, :9].reshape(3, 3).astype(np.float64), oMi[0, j, 9:].astype(np.float64)
            z = R @ np.asarray(hm.joints[j].axis, dtype=np.float64)
            J[:, hm.joints[j].idx_v] = np.cross(z, point.astype(np.float64) - o) if hm.joints[j].type == 1 else z
            j = hm.parents[j]
        jacs.append(J)
    return normal[0, pair_idx].astype(np.float64) * float(dist[0, pair_idx]), jacs[0], jacs[1]


# ----------------------------------------------------------------------------------------------
# Batched siblings (the point of the engine)
# ----------------------------------------------------------------------------------------------
def check_collisions_at_states(model, collision_model, Q, distance_padding=0.0, return_distance=False, return_pair=False):
    """Batched ``check_collisions_at_state``: ``Q`` [N, nq] (numpy or CUDA tensor) -> bool [N]
    (+ float32 min distance [N], + int32 first colliding pair index [N])."""
    return _engine(model, collision_model).check_states(Q, distance_padding, return_distance, return_pair)


def check_collisions_along_paths(model, collision_model, Q, offsets, distance_padding=0.0):
    """Batched ``check_collisions_along_path`` over a ragged array of pre-discretised paths."""
    return _engine(model, collision_model).check_paths(Q, offsets, distance_padding)


def sample_collision_free_states(model, collision_model, n, seed=0, joint_padding=0.0, distance_padding=0.0, as_numpy=True):
    """Batched ``get_random_collision_free_state``: ``n`` free states from a Philox stream keyed
    by ``seed`` (not the global ``np.random`` stream)."""
    return _engine(model, collision_model).sample_free_states(n, seed, joint_padding, distance_padding, as_numpy=as_numpy)


def compute_distances_at_states(model, collision_model, Q, cutoff=np.inf):
    """Batched ``pinocchio.computeDistances``: (dist [N, npairs], p1, p2, normal [N, npairs, 3])."""
    return _engine(model, collision_model).compute_distances(Q, cutoff)
