"""Cartesian (task-space) planning on the batched IK -- mirror of
/root/reference/src/pyroboplan/planning/cartesian_planner.py (``CartesianPlannerOptions``,
``CartesianPlanner``): a time-scaled straight-line / slerp interpolation between pose waypoints,
each sample solved by the collision-aware differential IK seeded with the previous solution.

``generate`` keeps the reference's single-trajectory API.  ``generate_batch`` runs the same
Cartesian plan from B different starting configurations at once (one ``solve_batch`` per time
sample): redundancy resolution and the collision gate depend on the seed, so racing several
seeds is the natural batched form of this sequential planner.
"""

import numpy as np
from scipy.spatial.transform import Rotation, Slerp

from ..model import SE3


class CartesianPlannerOptions:
    """Same fields, defaults and validation as the reference (cartesian_planner.py:10-52)."""

    def __init__(self, use_trapezoidal_scaling=True, max_linear_velocity=1.0, max_linear_acceleration=1.0,
                 max_angular_velocity=1.0, max_angular_acceleration=1.0):
        self.use_trapezoidal_scaling = use_trapezoidal_scaling
        for name, value in (("linear velocity", max_linear_velocity), ("linear acceleration", max_linear_acceleration),
                            ("angular velocity", max_angular_velocity), ("angular acceleration", max_angular_acceleration)):
            if value <= 0:
                raise ValueError(f"Max {name} must be positive.")
        self.max_linear_velocity = max_linear_velocity
        self.max_linear_acceleration = max_linear_acceleration
        self.max_angular_velocity = max_angular_velocity
        self.max_angular_acceleration = max_angular_acceleration


class _UnitTrapezoid:
    """Trapezoidal-velocity time scaling of the unit interval 0 -> 1 with peak speed ``v`` and
    acceleration ``a`` -- the single-segment case of the reference's TrapezoidalVelocityTrajectory
    (/root/reference/src/pyroboplan/trajectory/trapezoidal_velocity.py:60-140): triangular when
    v^2 >= a (the cruise speed is never reached), else accelerate / cruise / decelerate."""

    def __init__(self, v, a):
        self.a = a
        if not (np.isfinite(v) and np.isfinite(a)):
            self.t_ramp, self.t_cruise, self.peak = 0.0, 0.0, 0.0       # zero-length segment
        elif v * v >= a:
            self.peak = np.sqrt(a)
            self.t_ramp, self.t_cruise = self.peak / a, 0.0
        else:
            self.peak = v
            self.t_ramp = v / a
            self.t_cruise = (1.0 - v * v / a) / v
        self.duration = 2.0 * self.t_ramp + self.t_cruise

    def position(self, t):
        if self.duration == 0.0:
            return 1.0
        t = min(max(t, 0.0), self.duration)
        if t <= self.t_ramp:
            return 0.5 * self.a * t * t
        d_ramp = 0.5 * self.a * self.t_ramp**2
        if t <= self.t_ramp + self.t_cruise:
            return d_ramp + self.peak * (t - self.t_ramp)
        r = self.duration - t
        return 1.0 - 0.5 * self.a * r * r


class CartesianPlanner:
    """Cartesian motion planner (reference cartesian_planner.py:55-231)."""

    def __init__(self, model, target_frame, tforms, ik_solver, options=CartesianPlannerOptions()):
        self.model = model
        self.target_frame = target_frame
        self.tforms = tforms
        self.ik_solver = ik_solver
        self.options = options
        self.generated_tforms = []
        # (time scaling or None, end time) per segment, reference :93-152
        self.segments = []
        t_end = 0.0
        for start, end in zip(tforms[:-1], tforms[1:]):
            diff = end.actInv(start)
            lin = float(np.linalg.norm(diff.translation))
            ang = float(Rotation.from_matrix(np.asarray(diff.rotation)).magnitude())
            if options.use_trapezoidal_scaling:
                v = min(options.max_linear_velocity / lin if lin > 0 else np.inf,
                        options.max_angular_velocity / ang if ang > 0 else np.inf)
                a = min(options.max_linear_acceleration / lin if lin > 0 else np.inf,
                        options.max_angular_acceleration / ang if ang > 0 else np.inf)
                scaling = _UnitTrapezoid(v, a)
                t_end += scaling.duration
                self.segments.append((scaling, t_end))
            else:
                t_end += max(lin / options.max_linear_velocity, ang / options.max_angular_velocity)
                self.segments.append((None, t_end))

    # ------------------------------------------------------------------ sampling of the Cartesian path
    def _time_vector(self, dt):
        t_final = self.segments[-1][1]
        t_vec = np.arange(0.0, t_final, dt)
        if len(t_vec) == 0 or t_vec[-1] != t_final:
            t_vec = np.append(t_vec, t_final)
        return t_vec

    def _poses(self, t_vec):
        """Interpolated pose at every sample time (position lerp, rotation slerp, reference :186-215)."""
        out = []
        seg, t_prev = 0, 0.0
        for t in t_vec:
            while t > self.segments[seg][1] and seg + 1 < len(self.segments):
                t_prev = self.segments[seg][1]
                seg += 1
            scaling, t_seg_end = self.segments[seg]
            a, b = self.tforms[seg], self.tforms[seg + 1]
            local = t - t_prev
            if scaling is not None:
                alpha = scaling.position(min(local, scaling.duration))
            else:
                alpha = local / (t_seg_end - t_prev) if t_seg_end > t_prev else 1.0
            alpha = min(max(float(alpha), 0.0), 1.0)   # rounding of the time vector must not leave [0, 1]
            slerp = Slerp([0.0, 1.0], Rotation.concatenate([Rotation.from_matrix(np.asarray(a.rotation)),
                                                            Rotation.from_matrix(np.asarray(b.rotation))]))
            p = alpha * np.asarray(b.translation) + (1.0 - alpha) * np.asarray(a.translation)
            out.append(SE3(slerp(alpha).as_matrix(), p))
        return out

    # ------------------------------------------------------------------ reference API
    def generate(self, q_init, dt):
        """(success, t_vec, q_vec [nq, num_pts]); each sample's IK starts from the previous solution
        and the first failure aborts (reference :154-231).

        The reference passes the same ``q_cur`` array to every ``solve`` call and its IK updates
        that array in place (differential_ik.py:286-287 writes into ``init_state``), which is how
        the seed follows the path there; here the seed is advanced explicitly."""
        t_vec = self._time_vector(dt)
        self.generated_tforms = self._poses(t_vec)
        q = np.zeros([self.model.nq, len(t_vec)])
        q[:, 0] = q_init
        q_cur = np.array(q_init, dtype=np.float64)
        for idx, tform in enumerate(self.generated_tforms):
            q_sol = self.ik_solver.solve(self.target_frame, tform, init_state=q_cur)
            if q_sol is None:
                print(f"Failed to solve on point {idx + 1} out of {len(t_vec)}")
                return False, t_vec, q
            q[:, idx] = q_sol
            q_cur = np.array(q_sol, dtype=np.float64)
        return True, t_vec, q

    # ------------------------------------------------------------------ batched
    def generate_batch(self, q_inits, dt, seed=0):
        """The same Cartesian plan from B starting configurations at once.

        Returns (success [B] bool, t_vec, q_vec [B, nq, num_pts]).  Like ``generate`` every lane
        seeds a sample with its previous solution; a lane that fails a sample is dropped from the
        remaining samples."""
        import torch

        t_vec = self._time_vector(dt)
        self.generated_tforms = self._poses(t_vec)
        q_inits = np.asarray(q_inits, dtype=np.float64).reshape(-1, self.model.nq)
        B = q_inits.shape[0]
        q = np.zeros([B, self.model.nq, len(t_vec)])
        q[:, :, 0] = q_inits
        seeds = q_inits.copy()
        alive = np.ones(B, dtype=bool)
        for idx, tform in enumerate(self.generated_tforms):
            lanes = np.nonzero(alive)[0]
            if len(lanes) == 0:
                break
            target = np.concatenate([np.asarray(tform.rotation).reshape(9), np.asarray(tform.translation)])
            Q, solved, _ = self.ik_solver.solve_batch(self.target_frame, np.repeat(target[None, :], len(lanes), axis=0),
                                                      init_states=seeds[lanes], seed=seed + idx)
            ok = solved.cpu().numpy()
            Qh = Q.cpu().numpy()
            q[lanes[ok], :, idx] = Qh[ok]
            seeds[lanes[ok]] = Qh[ok]
            alive[lanes[~ok]] = False
        return alive, t_vec, q
