"""Planner-loop harness: the reference planners' CALL ORDER on a pluggable collision backend.

Not a re-implementation of the reference's planners (no Graph class, no A*, no visualisation): just the
sequence of hot-path calls their loops make, with the bookkeeping those calls depend on, so that the same
seeded loop can run on the drop-in functions (GPU engine) and on the float64 oracle and must grow the SAME
tree / roadmap.  Followed call sites:
  RRTPlanner.plan                /root/reference/src/pyroboplan/planning/rrt.py:125-261
  RRTPlanner.extend_or_connect   rrt.py:263-318   (has_collision_free_path per extension step)
  RRTPlanner.add_node_to_tree    rrt.py:350-402   (RRT* rewire: discretize + check_collisions_along_path)
  PRMPlanner.construct_roadmap   /root/reference/src/pyroboplan/planning/prm.py:115-168
  PRMPlanner.connect_node        prm.py:170-205   (k nearest within radius, has_collision_free_path each)
"""

import numpy as np

f32 = lambda q: np.asarray(q, dtype=np.float32).astype(np.float64)   # the rounding the engine applies to its inputs


class DropInBackend:
    """The package's drop-in functions (reference names and signatures): every call goes to the GPU engine."""

    name = "drop-in (GPU)"

    def __init__(self, model, cmodel, batched=True):
        from pyroboplan_b200.core import utils as cu
        from pyroboplan_b200.planning import rrt_batched, utils as pu

        self.m, self.c, self.cu, self.pu, self.rb, self.batched = model, cmodel, cu, pu, rrt_batched, batched

    def check_state(self, q):
        return bool(self.cu.check_collisions_at_state(self.m, self.c, q))

    def free_path(self, q1, q2, step):
        return bool(self.pu.has_collision_free_path(q1, q2, step, self.m, self.c))

    def path_collides(self, q1, q2, step):
        return bool(self.cu.check_collisions_along_path(self.m, self.c, self.pu.discretize_joint_space_path([q1, q2], step)))

    def rewire(self, qs, costs, q_new, parent, max_dist, step):
        if self.batched:   # all candidate segments in one device call, the reference's sequential rule on the host
            return self.rb.rewire_choice(self.m, self.c, qs, costs, q_new, parent, max_dist, step)
        return sequential_rewire(self, qs, costs, q_new, parent, max_dist, step)

    def free_paths(self, Q1, Q2, step):
        if self.batched:
            return np.asarray(self.pu.has_collision_free_paths(Q1, Q2, step, self.m, self.c))
        return np.array([self.free_path(a, b, step) for a, b in zip(Q1, Q2)], dtype=bool)


class OracleBackend:
    """The float64 oracle behind the same five calls (inputs rounded to float32 like the engine's)."""

    name = "oracle (CPU)"

    def __init__(self, model, cmodel, oracle):
        from pyroboplan_b200.planning.utils import discretize_joint_space_path

        self.o, self.disc, self.npairs = oracle, discretize_joint_space_path, len(cmodel.collisionPairs)

    def check_state(self, q):
        return self.npairs > 0 and self.o.check_state(f32(q))

    def free_path(self, q1, q2, step):
        return self.npairs == 0 or not bool(self.o.check_edges(f32(q1)[None], f32(q2)[None], step, nthreads=1)[0][0])

    def path_collides(self, q1, q2, step):
        return self.npairs > 0 and bool(self.o.check_path(np.array([f32(q) for q in self.disc([q1, q2], step)])))

    def rewire(self, qs, costs, q_new, parent, max_dist, step):
        return sequential_rewire(self, qs, costs, q_new, parent, max_dist, step)

    def free_paths(self, Q1, Q2, step):
        return np.array([self.free_path(a, b, step) for a, b in zip(Q1, Q2)], dtype=bool)


def sequential_rewire(be, qs, costs, q_new, parent, max_dist, step):
    """rrt.py:376-402 as written: nodes in tree order, running minimum."""
    best, min_cost = parent, costs[parent] + np.linalg.norm(qs[parent] - q_new)
    for k in range(len(qs)):
        if k == parent or np.array_equal(qs[k], q_new):
            continue
        d = np.linalg.norm(qs[k] - q_new)
        if d > max_dist:
            continue
        if costs[k] + d < min_cost and not be.path_collides(q_new, qs[k], step):
            best, min_cost = k, costs[k] + d
    return best, min_cost


class Tree:
    def __init__(self, q):
        self.q, self.parent, self.cost = [np.asarray(q, dtype=np.float64)], [-1], [0.0]

    def nearest(self, q):
        return int(np.argmin([np.linalg.norm(x - q) for x in self.q]))

    def add(self, be, q, parent, opt):
        """add_node_to_tree (rrt.py:350-402)."""
        cost = self.cost[parent] + np.linalg.norm(self.q[parent] - q)
        if opt["rrt_star"]:
            parent, cost = be.rewire(np.array(self.q), np.array(self.cost), q, parent, opt["max_rewire_dist"], opt["max_step_size"])
        self.q.append(q), self.parent.append(int(parent)), self.cost.append(float(cost))
        return len(self.q) - 1


def rrt_plan(be, model, q_start, q_goal, opt, seed, max_iters):
    """RRTPlanner.plan with a fixed iteration budget instead of a wall-clock limit (so two backends can be compared)."""
    from pyroboplan_b200.core.utils import get_random_state
    from pyroboplan_b200.planning.utils import extend_robot_state

    np.random.seed(seed)
    trees = [Tree(q_start), Tree(q_goal)]
    out = {"trees": trees, "found": False, "iters": 0, "link": None}
    if be.check_state(q_start) or be.check_state(q_goal):
        return out
    if np.linalg.norm(q_goal - q_start) <= opt["max_connection_dist"] and not be.path_collides(q_start, q_goal, opt["max_step_size"]):
        out.update(found=True, link=(trees[0].add(be, q_goal, 0, opt), 0))
        return out
    side = 0
    for it in range(max_iters):
        out["iters"] = it + 1
        tree, other = trees[side], trees[1 - side]
        q_sample = (q_goal if side == 0 else q_start) if np.random.random() < opt["goal_biasing_probability"] else get_random_state(model)
        cur, new = tree.nearest(q_sample), None
        if not np.array_equal(tree.q[cur], q_sample):          # extend_or_connect (rrt.py:263-318)
            while True:
                q_ext = extend_robot_state(tree.q[cur], q_sample, opt["max_connection_dist"])
                if not be.free_path(tree.q[cur], q_ext, opt["max_step_size"]):
                    break
                new = cur = tree.add(be, q_ext, cur, opt)
                if not opt["rrt_connect"] or np.array_equal(q_ext, q_sample):
                    break
        if new is not None:
            j = other.nearest(tree.q[new])
            d = np.linalg.norm(tree.q[new] - other.q[j])
            if d <= opt["max_connection_dist"] and not be.path_collides(tree.q[new], other.q[j], opt["max_step_size"]):
                if d > 0:
                    new = tree.add(be, other.q[j], new, opt)
                out.update(found=True, link=(new, j) if side == 0 else (j, new))
                return out
            if opt["bidirectional_rrt"]:
                side = 1 - side
    return out


def prm_build(be, model, n_nodes, radius, k, step, seed, max_draws=100000):
    """construct_roadmap + connect_node (prm.py:115-205): nodes one by one, k nearest EARLIER nodes within radius."""
    from pyroboplan_b200.core.utils import get_random_state

    np.random.seed(seed)
    nodes, edges = [], []
    for _ in range(max_draws):
        if len(nodes) >= n_nodes:
            break
        q = get_random_state(model)
        if be.check_state(q):
            continue
        if nodes:
            d = np.linalg.norm(np.array(nodes) - q, axis=1)
            near = [j for j in np.argsort(d, kind="stable") if d[j] <= radius][:k]
            if near:
                ok = be.free_paths(np.array([nodes[j] for j in near]), np.repeat(q[None], len(near), axis=0), step)
                edges += [(int(j), len(nodes)) for j, good in zip(near, ok) if good]
        nodes.append(q)
    return np.array(nodes), edges


def same_trees(a, b):
    return all(len(x.q) == len(y.q) and x.parent == y.parent and np.allclose(x.cost, y.cost, atol=1e-12) and
               all(np.array_equal(p, q) for p, q in zip(x.q, y.q)) for x, y in zip(a["trees"], b["trees"])) and \
        a["found"] == b["found"] and a["link"] == b["link"] and a["iters"] == b["iters"]
