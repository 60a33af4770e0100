#!/usr/bin/env python
"""BASELINE.json config 4 on 1..8 GPUs: PRM roadmap, 20 k collision-free nodes, k = 15 nearest predecessors, every
candidate edge validated; k-NN query blocks interleaved over the ranks, edges cut by prefix-summed sample counts,
two all-gathers (k-NN rows, bit-packed validity mask).  Strong scaling: the roadmap is the same for every N.

    python tools/gpu_config4.py                                   # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/gpu_config4.py

Rank 0 prints one JSON line; the mask is checked identical on every rank and equal to rank 0's unsharded result."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from bench import panda_models  # noqa: E402
from pyroboplan_b200.engine import get_engine  # noqa: E402
from pyroboplan_b200.planning.prm_batched import construct_roadmap_batched, nearest_predecessors  # noqa: E402

rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
model, cmodel = panda_models()
eng = get_engine(model, cmodel)
N, K = 20000, 15


def timed(fn, reps=5):
    out, ts = None, []
    for _ in range(reps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ts.append(float(t.item()))
    return out, sorted(ts)[len(ts) // 2], ts


build = lambda: construct_roadmap_batched(model, cmodel, n_nodes=N, k=K, radius=np.inf, max_step_size=0.05, seed=4)
build()                                                                   # warm-up: buffers, NCCL channels
rm, ms_total, all_total = timed(build)
nodes = rm["nodes"]
_, ms_sample, _ = timed(lambda: eng.sample_free_states(N, seed=4))
_, ms_knn, _ = timed(lambda: nearest_predecessors(nodes, K, np.inf, engine=eng))
cand, valid = rm["candidates"], rm["valid"]
ref = valid.clone().to(torch.uint8)
if world > 1:
    dist.broadcast(ref, src=0)
same = bool((ref == valid.to(torch.uint8)).all())
alone = ~eng.check_edges(nodes[cand[:, 0]].contiguous(), nodes[cand[:, 1]].contiguous(), 0.05)
flags = torch.tensor([int(same), int(bool((alone == valid).all()))], device=dev)
if world > 1:
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
if rank == 0:
    samples = float((torch.ceil(rm["lengths"].double() / 0.05) + 1).sum())
    print(json.dumps({"config": "config4: PRM roadmap, 20000 free Panda nodes, k=15 nearest predecessors, step 0.05 rad", "n_gpus": world,
                      "scaling": "strong", "ms_total_median_of_5_max_over_ranks": ms_total, "ms_total_all": all_total, "ms_sampling": ms_sample,
                      "ms_knn_incl_allgather": ms_knn, "ms_edges_incl_allgather": ms_total - ms_sample - ms_knn,
                      "candidate_edges": int(cand.shape[0]), "valid_edges": int(valid.sum().item()), "edge_samples_covered": samples,
                      "edges_per_s": cand.shape[0] / (ms_total * 1e-3), "mask_bytes_on_the_wire_per_rank": int((cand.shape[0] / world + 7) // 8),
                      "mask_identical_on_all_ranks": bool(flags[0].item()), "mask_equals_unsharded": bool(flags[1].item())}))
if world > 1:
    dist.destroy_process_group()
