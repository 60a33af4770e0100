"""Collision-aware differential IK on the B200 engine -- mirror of
/root/reference/src/pyroboplan/ik/differential_ik.py (``DifferentialIkOptions``, ``DifferentialIk``).

``solve`` keeps the reference's control flow (tries, wrap, limit check, collision gate, random
restarts drawn from the global ``np.random`` stream) but runs a whole try -- up to ``max_iters``
damped-least-squares iterations -- in one device call.  ``solve_batch`` does the same for N targets
at once (BASELINE.json config 5): the first try is one kernel launch over all targets, then ALL
random restarts of every still-unsolved target run in one more launch (the earliest successful
restart wins, as in the reference's sequential loop); the collision gate is one batched
``check_states`` and restarts come from the device sampler.
Null-space components (host callables ``lambda model, q: ...`` as in the reference; batched
callables ``lambda model, Q: ...`` on CUDA tensors for ``solve_batch``) are re-evaluated every
iteration, so with them a try runs one device call per iteration instead of one per try.
"""

import numpy as np

from ..core.utils import check_collisions_at_state, get_random_collision_free_state, get_random_state
from ..engine import get_engine
from ..model import GeometryModel


class DifferentialIkOptions:
    """Options for differential IK (same fields and defaults as the reference, differential_ik.py:19-77)."""

    def __init__(self, max_iters=200, max_retries=10, max_translation_error=1e-3, max_rotation_error=1e-3, damping=1e-3,
                 min_step_size=0.1, max_step_size=0.5, ignore_joint_indices=[], joint_weights=None, rng_seed=None):
        self.max_iters = max_iters
        self.max_retries = max_retries
        self.max_translation_error = max_translation_error
        self.max_rotation_error = max_rotation_error
        self.damping = damping
        self.min_step_size = min_step_size
        self.max_step_size = max_step_size
        self.joint_weights = joint_weights
        self.ignore_joint_indices = ignore_joint_indices
        self.rng_seed = rng_seed


def _se3_rows(tform):
    """SE3-like object / 4x4 / 12-vector -> 12 floats (R row-major, then t)."""
    if hasattr(tform, "rotation"):
        return np.concatenate([np.asarray(tform.rotation, dtype=np.float64).reshape(9), np.asarray(tform.translation, dtype=np.float64)])
    a = np.asarray(tform, dtype=np.float64)
    if a.shape == (4, 4):
        return np.concatenate([a[:3, :3].reshape(9), a[:3, 3]])
    return a.reshape(12)


class DifferentialIk:
    """Damped-least-squares differential IK solver (reference differential_ik.py:81-320)."""

    def __init__(self, model, collision_model=None, data=None, collision_data=None, visualizer=None,
                 options=DifferentialIkOptions()):
        self.model = model
        self.collision_model = collision_model
        self.options = options
        self.visualizer = visualizer  # accepted for signature compatibility; nothing is drawn
        # FK / Jacobian only need the kinematic tree: use the collision model's engine when there
        # is one (shares its device tables), otherwise an engine over an empty geometry model
        if collision_model is not None:
            self._engine_model = collision_model
        else:
            from ..core.utils import fk_holder

            self._engine_model = fk_holder(model)

    # ------------------------------------------------------------------ helpers
    def _setup(self):
        opt = self.options
        active = [i for i in range(self.model.nq) if i not in opt.ignore_joint_indices]
        if opt.joint_weights is None:
            w_diag = None
        elif len(opt.joint_weights) != len(active):
            raise ValueError(f"Joint weights, if specified, must have {len(active)} elements.")
        elif np.any(np.array(opt.joint_weights) <= 0.0):
            raise ValueError("All joint weights must be strictly positive.")
        else:
            w_diag = np.ones(self.model.nq, dtype=np.float64)
            w_diag[active] = 1.0 / np.asarray(opt.joint_weights, dtype=np.float64)  # W = inv(diag(weights))
        mask = 0
        for i in active:
            mask |= 1 << i
        return mask, w_diag

    def _try(self, eng, frame_id, targets, Q, e0, mask, w_diag, nullspace_fn=None):
        """One try: up to max_iters iterations.  nullspace_fn(Q [M, nq] float64 tensor) -> [M, nq]."""
        import torch

        o = self.options
        if nullspace_fn is None:
            return eng.ik_iterate(frame_id, targets, Q, e0, o.max_iters, o.max_translation_error, o.max_rotation_error,
                                  o.damping, o.min_step_size, o.max_step_size, mask, w_diag)
        # the null-space term depends on q: one single-iteration launch per iteration over the
        # targets that have not converged yet (the reference's loop, :208-298, iteration by iteration)
        Q = torch.as_tensor(Q, device=eng.torch_device).to(torch.float64).reshape(-1, eng.nq).clone()
        n = Q.shape[0]
        e0 = torch.zeros(n, dtype=torch.float64, device=eng.torch_device) if e0 is None else e0.to(torch.float64).clone()
        status = torch.zeros(n, dtype=torch.int32, device=eng.torch_device)
        iters = torch.zeros(n, dtype=torch.int32, device=eng.torch_device)
        live = torch.arange(n, device=eng.torch_device)
        for _ in range(o.max_iters):
            if live.numel() == 0:
                break
            Ql = Q[live]
            ns = nullspace_fn(Ql)
            Qo, st, it, e0l = eng.ik_iterate(frame_id, targets[live], Ql, e0[live], 1, o.max_translation_error, o.max_rotation_error,
                                             o.damping, o.min_step_size, o.max_step_size, mask, w_diag, nullspace=ns)
            Q[live], status[live], e0[live] = Qo, st, e0l
            iters[live] += it
            finite = torch.isfinite(Qo).all(dim=1)
            live = live[((st & 1) == 0) & finite]
        return Q, status, iters, e0

    # ------------------------------------------------------------------ single target (reference API)
    def solve(self, target_frame, target_tform, init_state=None, nullspace_components=[], verbose=False):
        """Solves one IK query; returns the joint configuration or None (reference :134-320)."""
        import torch

        np.random.seed(self.options.rng_seed)
        frame_id = self.model.getFrameId(target_frame)
        if frame_id >= self.model.nframes:
            raise ValueError(f"unknown frame {target_frame!r}")
        mask, w_diag = self._setup()
        eng = get_engine(self.model, self._engine_model)
        if init_state is None:
            init_state = get_random_state(self.model)
        target = torch.as_tensor(_se3_rows(target_tform), dtype=torch.float64, device=eng.torch_device).reshape(1, 12)
        q_cur = np.array(init_state, dtype=np.float64)
        e0 = None
        n_tries = 0
        ns_fn = None
        if nullspace_components:
            def ns_fn(Qt):   # host callables of the reference's form, summed (:274-279)
                qh = Qt[0].cpu().numpy()
                return torch.as_tensor(sum(np.asarray(comp(self.model, qh), dtype=np.float64) for comp in nullspace_components),
                                       device=Qt.device).reshape(1, -1)
        while n_tries <= self.options.max_retries:
            Q = torch.as_tensor(q_cur, dtype=torch.float64, device=eng.torch_device).reshape(1, -1)
            Qo, status, _, e0 = self._try(eng, frame_id, target, Q, e0, mask, w_diag, ns_fn)
            st = int(status[0].item())
            q_cur = Qo[0].cpu().numpy()
            if st & 1:
                if st & 2:
                    if self.collision_model is not None and check_collisions_at_state(self.model, self.collision_model, q_cur):
                        if verbose:
                            print("Solved within joint limits, but in collision.")
                    else:
                        if verbose:
                            print(f"Solved in {n_tries + 1} tries.")
                        return q_cur
                elif verbose:
                    print("Solved, but outside joint limits.")
            if self.collision_model is not None:
                q_cur = get_random_collision_free_state(self.model, self.collision_model)
                if q_cur is None:
                    return None
            else:
                q_cur = get_random_state(self.model)
            n_tries += 1
            if verbose:
                print(f"Retry {n_tries}")
        return None

    # ------------------------------------------------------------------ batched
    def solve_batch(self, target_frame, target_tforms, init_states=None, seed=0, nullspace_components=[]):
        """Solves N IK queries at once.

        nullspace_components: batched callables ``lambda model, Q: tensor [M, nq]`` on CUDA tensors
        (see ``pyroboplan_b200.ik.nullspace_components.*_batch``), summed and re-evaluated every iteration.

        target_tforms: [N, 12] tensor/array (R row-major, then t) or a list of SE3-like objects.
        Returns (Q [N, nq] float64 device tensor, solved [N] bool device tensor, tries [N] int32).
        Restarts use the device sampler (Philox stream ``seed``), not ``np.random``.
        """
        import torch

        frame_id = self.model.getFrameId(target_frame)
        if frame_id >= self.model.nframes:
            raise ValueError(f"unknown frame {target_frame!r}")
        mask, w_diag = self._setup()
        eng = get_engine(self.model, self._engine_model)
        dev = eng.torch_device
        if isinstance(target_tforms, (list, tuple)):
            target_tforms = np.stack([_se3_rows(t) for t in target_tforms])
        T = torch.as_tensor(target_tforms, device=dev).to(torch.float64).reshape(-1, 12).contiguous()
        n = T.shape[0]
        has_pairs = self.collision_model is not None and len(self.collision_model.collisionPairs) > 0

        def random_states(count, salt):
            if has_pairs:
                s = eng.sample_free_states(count, seed=seed * 7919 + salt)
                if s.shape[0] < count:
                    raise RuntimeError("sampler could not find enough collision-free restart states")
                return s.to(torch.float64)
            g = torch.Generator(device=dev)
            g.manual_seed(seed * 7919 + salt)
            lo = torch.as_tensor(self.model.lowerPositionLimit, dtype=torch.float64, device=dev)
            hi = torch.as_tensor(self.model.upperPositionLimit, dtype=torch.float64, device=dev)
            return lo + (hi - lo) * torch.rand((count, self.model.nq), generator=g, device=dev, dtype=torch.float64)

        Q = random_states(n, 0) if init_states is None else torch.as_tensor(init_states, device=dev).to(torch.float64).reshape(n, -1).clone()
        e0 = torch.zeros(n, dtype=torch.float64, device=dev)
        solved = torch.zeros(n, dtype=torch.bool, device=dev)
        tries = torch.zeros(n, dtype=torch.int32, device=dev)
        ns_fn = (lambda Qt: sum(comp(self.model, Qt) for comp in nullspace_components)) if nullspace_components else None

        def attempt(idx, Qin):
            """One try for the targets idx from the states Qin -> (Q_out, accepted, e0_out)."""
            Qo, status, _, e0p = self._try(eng, frame_id, T[idx], Qin, e0[idx], mask, w_diag, ns_fn)
            ok = (status & 3) == 3
            if has_pairs and ok.any():
                hit = eng.check_states(Qo[ok].to(torch.float32))
                ok_idx = torch.nonzero(ok).flatten()
                ok[ok_idx[hit]] = False
            return Qo, ok, e0p

        # first try: every target from its initial state
        everyone = torch.arange(n, device=dev)
        Qo, ok, e0p = attempt(everyone, Q)
        e0[:] = e0p            # the initial error norm is captured once per solve and kept across retries (:200,264-265)
        Q[:] = Qo
        solved[ok] = True
        tries[:] = 1
        pending = everyone[~ok]
        R = int(self.options.max_retries)
        if pending.numel() and R > 0:
            # The retries of a target are independent random restarts, and the reference returns the
            # FIRST one that succeeds.  A try is latency-bound (up to max_iters dependent iterations,
            # ~3 ms on the device however few targets are left), so all R restarts of every pending
            # target run in ONE launch and the earliest successful one is kept: same outcome as the
            # sequential loop for the same restart draws, one launch latency instead of R.
            P = pending.numel()
            restarts = random_states(P * R, 1).reshape(P, R, -1)
            idx = pending.repeat_interleave(R)
            Qr, okr, _ = attempt(idx, restarts.reshape(P * R, -1))
            okr = okr.reshape(P, R)
            any_ok = okr.any(dim=1)
            first = torch.argmax(okr.to(torch.int8), dim=1)              # earliest successful restart
            chosen = torch.where(any_ok, first, torch.full_like(first, R - 1))   # none: report the last try's state
            Q[pending] = Qr.reshape(P, R, -1)[torch.arange(P, device=dev), chosen]
            solved[pending] = any_ok
            tries[pending] = (1 + torch.where(any_ok, first + 1, torch.full_like(first, R))).to(torch.int32)
        return Q, solved, tries
