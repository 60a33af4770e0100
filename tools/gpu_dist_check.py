#!/usr/bin/env python
"""N-rank NCCL check (run under torchrun on N GPUs): the sharded PRM edge-validity mask and the
sharded state verdicts are identical on every rank and equal to the single-rank result."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from bench import panda_models  # noqa: E402
from pyroboplan_b200.distributed import check_states_sharded  # noqa: E402
from pyroboplan_b200.engine import get_engine  # noqa: E402
from pyroboplan_b200.planning.prm_batched import construct_roadmap_batched  # noqa: E402

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
model, cmodel = panda_models()
eng = get_engine(model, cmodel)
nodes = eng.sample_free_states(20000, seed=4)          # same Philox stream on every rank -> same nodes
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
rm = construct_roadmap_batched(model, cmodel, n_nodes=20000, k=15, radius=np.inf, max_step_size=0.05, nodes=nodes)   # warm-up
dist.barrier()
torch.cuda.synchronize()
ev0.record()
rm = construct_roadmap_batched(model, cmodel, n_nodes=20000, k=15, radius=np.inf, max_step_size=0.05, nodes=nodes)
ev1.record()
torch.cuda.synchronize()
ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
dist.all_reduce(ms, op=dist.ReduceOp.MAX)
valid = rm["valid"]
# every rank holds the same full mask
ref = valid.clone().to(torch.uint8)
dist.broadcast(ref, src=0)
same = bool((ref == valid.to(torch.uint8)).all())
# and it equals the unsharded computation on this rank
cand = rm["candidates"]
alone = ~eng.check_edges(nodes[cand[:, 0]].contiguous(), nodes[cand[:, 1]].contiguous(), 0.05)
eq = bool((alone == valid).all())
Q = torch.as_tensor(np.random.default_rng(0).uniform(model.lowerPositionLimit, model.upperPositionLimit, size=(200001, 9)).astype(np.float32), device=dev)
hs = check_states_sharded(eng.check_states, Q)
eq_s = bool((hs == eng.check_states(Q)).all())
flags = torch.tensor([int(same), int(eq), int(eq_s)], device=dev)
dist.all_reduce(flags, op=dist.ReduceOp.MIN)
if rank == 0:
    print({"world": world, "prm_ms_max_over_ranks": float(ms.item()), "candidates": int(cand.shape[0]), "valid": int(valid.sum().item()),
           "mask_identical_on_all_ranks": bool(flags[0].item()), "mask_equals_unsharded": bool(flags[1].item()),
           "states_equal_unsharded": bool(flags[2].item())})
dist.destroy_process_group()
