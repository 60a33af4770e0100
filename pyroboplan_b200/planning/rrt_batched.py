"""Batched edge validation for the spots of RRT / RRT* / shortcutting where many independent
edges meet.  The planners' control flow stays with the caller (the
reference's own ``RRTPlanner`` runs unchanged on this package's ``has_collision_free_path``);
these helpers replace its inner per-edge loops by one device call and then apply the
reference's SEQUENTIAL selection rule on the host, so the outcome is identical.

Reference: /root/reference/src/pyroboplan/planning/rrt.py:376-402 (RRT* rewire),
/root/reference/src/pyroboplan/planning/rrt.py:150-260 (extend / connect steps),
/root/reference/src/pyroboplan/planning/path_shortcutting.py:110-145.
"""

import numpy as np

from ..core.utils import _engine


def rewire_choice(model, collision_model, node_qs, node_costs, q_new, parent_index, max_rewire_dist, max_step_size,
                  distance_padding=0.0):
    """RRT* parent selection for a node just attached to ``parent_index``.

    node_qs [M, nq] / node_costs [M]: the tree's nodes in iteration order (``tree.nodes``), NOT
    including the new node.  Returns ``(best_parent_index, new_node_cost)``.

    The reference walks the nodes in order and re-parents whenever ``other.cost + dist`` beats the
    running minimum and the segment ``q_new -> other.q`` is collision free (rrt.py:376-402).
    Only nodes that beat the INITIAL cost can ever be chosen, so exactly those segments are
    validated, all at once; the running-minimum rule is then replayed on the host.
    """
    node_qs = np.asarray(node_qs, dtype=np.float64).reshape(len(node_costs), -1)
    node_costs = np.asarray(node_costs, dtype=np.float64)
    q_new = np.asarray(q_new, dtype=np.float64)
    # Distances exactly as the reference forms them -- ``np.linalg.norm`` of ONE difference vector
    # (core/utils.py:135-151) -- for every node a vectorised pre-filter (with slack) lets through: on an
    # RRT-Connect branch the nodes are collinear, ``other.cost + dist`` ties with the running minimum up
    # to rounding, and a row-wise norm (different summation order) would break such ties differently.
    rough = np.linalg.norm(node_qs - q_new, axis=1)
    cost0 = node_costs[parent_index] + np.linalg.norm(node_qs[parent_index] - q_new)
    slack = 1e-9 * (1.0 + abs(cost0))
    pre = np.nonzero((rough <= max_rewire_dist + slack) & (node_costs + rough < cost0 + slack))[0]
    dist = np.full(len(node_costs), np.inf)
    for k in pre:
        dist[k] = np.linalg.norm(node_qs[k] - q_new)
    new_cost = node_costs + dist
    cand = np.nonzero((dist <= max_rewire_dist) & (new_cost < cost0) & (np.arange(len(dist)) != parent_index)
                      & ~np.all(node_qs == q_new, axis=1))[0]
    best, min_cost = parent_index, cost0
    if len(cand) == 0:
        return best, min_cost
    if len(collision_model.collisionPairs) == 0:
        free = np.ones(len(cand), dtype=bool)
    else:
        hit = _engine(model, collision_model).check_edges(np.repeat(q_new[None, :], len(cand), axis=0), node_qs[cand],
                                                          max_step_size, distance_padding)
        free = ~np.asarray(hit)
    for k, ok in zip(cand, free):       # tree order, running minimum: the reference's rule
        if new_cost[k] < min_cost and ok:
            best, min_cost = int(k), float(new_cost[k])
    return best, min_cost


def validate_extensions(model, collision_model, q_from, q_to, max_step_size, distance_padding=0.0):
    """``has_collision_free_path`` for many (q_from[e] -> q_to[e]) extension steps at once -> bool [E].
    Useful for batched / multi-query RRT variants that extend several trees or samples per step."""
    q_from = np.asarray(q_from, dtype=np.float64)
    q_to = np.asarray(q_to, dtype=np.float64)
    if len(collision_model.collisionPairs) == 0:
        return np.ones(len(q_from), dtype=bool)
    return ~np.asarray(_engine(model, collision_model).check_edges(q_from, q_to, max_step_size, distance_padding))


def shortcut_candidates_free(model, collision_model, q_lows, q_highs, max_step_size=0.05):
    """Collision verdicts of many candidate shortcuts (q_lows[e] -> q_highs[e]) of ONE fixed path in
    one call -> bool [E] (True = the shortcut is collision free).  The reference samples and tests
    one candidate per iteration (path_shortcutting.py:118-143); a caller may draw the candidates
    of several iterations up front and consume them until the first success changes the path."""
    return validate_extensions(model, collision_model, q_lows, q_highs, max_step_size)
