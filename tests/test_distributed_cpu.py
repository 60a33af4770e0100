"""World-size-2 gloo test of the N>1 host logic: block sharding + mask all-gather.

The rank-local validator here is the CPU oracle (test stand-in for ``CollisionEngine.check_edges``
-- there is no GPU in this container); what is under test is the sharding and the collective."""

import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, n_edges, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from conftest import make_panda, uniform_states
    from oracle.oracle import Oracle
    from pyroboplan_b200.distributed import check_states_sharded, shard_bounds, validate_edges_sharded

    model, cmodel, _ = make_panda()
    orc = Oracle(model, cmodel)
    a = uniform_states(model, n_edges, 1)
    b = a + np.float32(0.1) * (uniform_states(model, n_edges, 2) - a)

    def check_edges(q1, q2, step, padding):
        hit, _, _, _ = orc.check_edges(q1.astype(np.float64), q2.astype(np.float64), step, padding, nthreads=2)
        return torch.as_tensor(hit)

    def check_states(q, padding):
        return torch.as_tensor(orc.check_states(q.astype(np.float64), padding, nthreads=2)[0])

    valid = validate_edges_sharded(check_edges, a, b, 0.05)
    states = check_states_sharded(check_states, a)
    lo, hi = shard_bounds(n_edges, rank, world)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), valid=valid.numpy(), states=states.numpy(), lo=lo, hi=hi)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_edges", [301, 64])
def test_sharded_edge_mask_world2(tmp_path, n_edges):
    world = 2
    port = 29500 + (os.getpid() % 2000) + n_edges % 7
    mp.spawn(_worker, args=(world, port, n_edges, str(tmp_path)), nprocs=world, join=True)
    from conftest import make_panda, uniform_states
    from oracle.oracle import Oracle

    model, cmodel, _ = make_panda()
    orc = Oracle(model, cmodel)
    a = uniform_states(model, n_edges, 1)
    b = a + np.float32(0.1) * (uniform_states(model, n_edges, 2) - a)
    hit, _, _, _ = orc.check_edges(a.astype(np.float64), b.astype(np.float64), 0.05)
    hs = orc.check_states(a.astype(np.float64))[0]
    r0, r1 = np.load(tmp_path / "rank0.npz"), np.load(tmp_path / "rank1.npz")
    assert (r0["valid"] == ~hit.astype(bool)).all()        # full mask, identical to the single-process result
    assert (r1["valid"] == r0["valid"]).all()               # and identical on every rank
    assert (r0["states"] == hs.astype(bool)).all() and (r1["states"] == r0["states"]).all()
    assert (r0["lo"], r0["hi"], r1["lo"], r1["hi"]) == (0, (n_edges + 1) // 2, (n_edges + 1) // 2, n_edges)


def test_shard_bounds_cover_everything():
    from pyroboplan_b200.distributed import shard_bounds

    for n in (0, 1, 7, 8, 1000003):
        for world in (1, 2, 3, 8):
            blocks = [shard_bounds(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(b0[1] == b1[0] for b0, b1 in zip(blocks[:-1], blocks[1:]))
            assert all(hi - lo <= (n + world - 1) // world for lo, hi in blocks)
