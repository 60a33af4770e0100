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


# ---------------------------------------------------------------------------------------------------
# PRM construction across ranks: balanced edge cuts, bit-packed mask, interleaved k-NN (world-size-2 gloo)
# ---------------------------------------------------------------------------------------------------
def test_balanced_bounds_and_bit_packing():
    from pyroboplan_b200.distributed import balanced_bounds, pack_bits, unpack_bits

    rng = np.random.default_rng(0)
    for n in (0, 1, 5, 1000, 4097):
        w = rng.integers(1, 220, size=n).astype(np.float64)
        for world in (1, 2, 3, 8):
            cuts = balanced_bounds(w, world)
            assert cuts[0] == 0 and cuts[-1] == n and all(a <= b for a, b in zip(cuts[:-1], cuts[1:]))
            if n >= 1000:   # every block within one item's weight of the ideal share
                sums = [w[a:b].sum() for a, b in zip(cuts[:-1], cuts[1:])]
                assert max(sums) - min(sums) <= 2 * w.max()
        m = torch.as_tensor(rng.integers(0, 2, size=n).astype(bool))
        p = pack_bits(m)
        assert p.numel() == (n + 7) // 8 and p.dtype == torch.uint8
        assert (unpack_bits(p, n) == m).all()
    assert balanced_bounds(np.zeros(10), 4)[-1] == 10


def _knn_numpy(nodes, k, first, stride, block):
    """Predecessor k-NN of the interleaved query blocks, brute force (stand-in for CollisionEngine.knn)."""
    n = len(nodes)
    nblocks = (n + block - 1) // block
    rows = []
    for b in range(first, nblocks, stride):
        for i in range(b * block, (b + 1) * block):
            idx, dd = np.full(k, -1, np.int32), np.full(k, np.inf, np.float32)
            if 0 < i < n:
                d = np.linalg.norm(nodes[:i].astype(np.float64) - nodes[i].astype(np.float64), axis=1)
                o = np.argsort(d, kind="stable")[:k]
                idx[: len(o)], dd[: len(o)] = o, d[o]
            rows.append((idx, dd))
    return torch.as_tensor(np.array([r[0] for r in rows])), torch.as_tensor(np.array([r[1] for r in rows]))


def _prm_worker(rank, world, port, n_nodes, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from conftest import make_panda, uniform_states
    from oracle.oracle import Oracle
    from pyroboplan_b200.distributed import knn_interleaved, validate_edges_balanced

    model, cmodel, _ = make_panda()
    orc = Oracle(model, cmodel)
    nodes = uniform_states(model, n_nodes, 5)
    k = 4
    idx, d = knn_interleaved(lambda first, stride: _knn_numpy(nodes, k, first, stride, 16), n_nodes, k, 16, torch.device("cpu"))
    keep = idx >= 0
    i_of = torch.arange(n_nodes)[:, None].expand_as(idx)
    q1, q2 = nodes[idx[keep].numpy()], nodes[i_of[keep].numpy()]
    seen = []

    def check_edges(a, b, step, padding):
        seen.append(len(a))
        return torch.as_tensor(orc.check_edges(np.asarray(a, np.float64), np.asarray(b, np.float64), step, padding, nthreads=2)[0])

    valid = validate_edges_balanced(check_edges, q1, q2, 0.05, lengths=d[keep])
    np.savez(os.path.join(out_dir, f"prm{rank}.npz"), idx=idx.numpy(), dist=d.numpy(), valid=valid.numpy(), mine=sum(seen),
             samples=float((torch.ceil(d[keep].double() / 0.05) + 1).sum()))
    dist.barrier()
    dist.destroy_process_group()


def test_prm_sharding_world2(tmp_path):
    world, n_nodes = 2, 150
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_prm_worker, args=(world, port, n_nodes, str(tmp_path)), nprocs=world, join=True)
    from conftest import make_panda, uniform_states
    from oracle.oracle import Oracle

    model, cmodel, _ = make_panda()
    nodes = uniform_states(model, n_nodes, 5)
    idx_ref, d_ref = _knn_numpy(nodes, 4, 0, 1, 16)
    r0, r1 = np.load(tmp_path / "prm0.npz"), np.load(tmp_path / "prm1.npz")
    for r in (r0, r1):      # the interleaved rows come back in node order, identical on both ranks
        assert (r["idx"] == idx_ref.numpy()[:n_nodes]).all() and np.array_equal(r["dist"], d_ref.numpy()[:n_nodes])
    keep = r0["idx"] >= 0
    i_of = np.repeat(np.arange(n_nodes)[:, None], 4, axis=1)
    hit = Oracle(model, cmodel).check_edges(nodes[r0["idx"][keep]].astype(np.float64), nodes[i_of[keep]].astype(np.float64), 0.05)[0]
    assert (r0["valid"] == ~hit.astype(bool)).all() and (r1["valid"] == r0["valid"]).all()
    assert r0["mine"] + r1["mine"] == keep.sum() and r0["mine"] > 0 and r1["mine"] > 0   # every edge validated exactly once
