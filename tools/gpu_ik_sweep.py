#!/usr/bin/env python
"""Config 5 across GPUs (weak scaling: the same 65 536 Cartesian targets on every rank -- targets are
independent, a batch shards without a collective).  Launch with torchrun; rank 0 prints one JSON line.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/gpu_ik_sweep.py
"""

import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from pyroboplan_b200.core.utils import _engine  # noqa: E402
from pyroboplan_b200.ik.differential_ik import DifferentialIk, DifferentialIkOptions  # noqa: E402
from pyroboplan_b200.models import panda  # noqa: E402


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    model, cmodel, vmodel = panda.load_models()
    panda.add_self_collisions(model, cmodel)
    panda.add_object_collisions(model, cmodel, vmodel)
    eng = _engine(model, cmodel)
    n = 65536
    fid = model.getFrameId("panda_hand")
    goals = eng.sample_free_states(n, seed=2)              # the same synthetic batch on every rank: fixed work per GPU
    _, _, oMf = eng.fk(goals, joints=False, geoms=False, frames=True)
    targets = oMf[:, fid].contiguous()
    ik = DifferentialIk(model, collision_model=cmodel, options=DifferentialIkOptions(ignore_joint_indices=[7, 8]))
    for rep in range(3):                                   # untimed: sizes the device buffers for these restart draws
        ik.solve_batch("panda_hand", targets, seed=5 + rep)
    walls, rates = [], []
    for rep in range(3):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        Q, solved, tries = ik.solve_batch("panda_hand", targets, seed=5 + rep)
        torch.cuda.synchronize()
        t = torch.tensor([time.perf_counter() - t0], device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        walls.append(float(t.item()))
        rates.append(float(solved.float().mean()))
    if rank == 0:
        t_med = sorted(walls)[1]
        print(json.dumps({"config": "config5: 65536 IK targets per GPU, DifferentialIkOptions() defaults, collision gate + restarts",
                          "n_gpus": world, "scaling": "weak", "wall_s_median_of_3_max_over_ranks": t_med, "wall_s_all": walls,
                          "solves_per_s": world * n / t_med, "success_rate_rank0": rates[-1]}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
