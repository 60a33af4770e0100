#!/usr/bin/env python
"""Developer smoke/parity/timing run on a real B200 (run through gpurun). Prints diagnostics."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from oracle.oracle import Oracle  # noqa: E402
from pyroboplan_b200.engine import CollisionEngine  # noqa: E402
from pyroboplan_b200.models import panda, two_dof  # noqa: E402


def panda_models():
    m, c, v = panda.load_models()
    panda.add_self_collisions(m, c)
    panda.add_object_collisions(m, c, v)
    return m, c


def classify(name, hit_gpu, hit_cpu, md_cpu):
    mism = np.nonzero(hit_gpu != hit_cpu)[0]
    print(f"[{name}] N={len(hit_cpu)} P(hit) cpu={hit_cpu.mean():.4f} gpu={hit_gpu.mean():.4f} mismatches={len(mism)}")
    if len(mism):
        d = md_cpu[mism]
        print("   mismatch oracle min-dist:", np.sort(d)[:20], " #|d|>1e-6:", int((np.abs(d) > 1e-6).sum()))
        print("   idx", mism[:10], "gpu", hit_gpu[mism[:10]], "cpu", hit_cpu[mism[:10]])
    return mism


def main():
    N = int(os.environ.get("N", 200000))
    dev = torch.device("cuda:0")
    print(torch.cuda.get_device_name(0))
    m, c = panda_models()
    t0 = time.time()
    eng = CollisionEngine(m, c)
    print("engine create s:", time.time() - t0, "npairs", eng.npairs)
    orc = Oracle(m, c)
    rng = np.random.default_rng(0)
    Q64 = rng.uniform(m.lowerPositionLimit, m.upperPositionLimit, size=(N, 9))
    Q32 = Q64.astype(np.float32)
    Qo = Q32.astype(np.float64)  # the oracle sees exactly the float32 inputs

    # 1. FK parity
    oMi, oMg, oMf = eng.fk(Q32[:2000])
    err = 0.0
    for i in range(200):
        a, b = orc.fk(Qo[i])
        err = max(err, np.abs(oMg[i] - b).max(), np.abs(oMi[i] - a).max())
    print("[fk] max abs err vs oracle (200 states):", err)

    # 2. boolean parity
    t0 = time.time()
    hit_cpu, md_cpu, fp_cpu, cnt = orc.check_states(Qo, want_all=True, use_cull=False)
    print("oracle all-pairs s:", time.time() - t0, cnt)
    hit_gpu = eng.check_states(Q32)
    classify("panda states host", hit_gpu.astype(np.uint8), hit_cpu, md_cpu)
    Qd = torch.as_tensor(Q32, device=dev)
    hit_d = eng.check_states(Qd).cpu().numpy()
    print("   device==host path:", bool((hit_d == hit_gpu).all()))

    # 3. distance + pair parity
    h2, dist, pair = eng.check_states(Q32, return_distance=True, return_pair=True)
    classify("panda states dist-mode", h2.astype(np.uint8), hit_cpu, md_cpu)
    free = md_cpu > 0
    print("[dist] max abs err (free):", np.abs(dist[free] - md_cpu[free]).max(), " penetrating gpu max:", dist[~free].max() if (~free).any() else None)
    print("[pair] first-pair mismatches:", int((pair != fp_cpu).sum()))
    bad = np.nonzero(pair != fp_cpu)[0][:5]
    for i in bad:
        print("    ", i, pair[i], fp_cpu[i], md_cpu[i])

    # padding
    pad = 0.03
    hp_cpu, mdp, _, _ = orc.check_states(Qo[:50000], padding=pad)
    hp_gpu = eng.check_states(Q32[:50000], padding=pad)
    mm = np.nonzero(hp_gpu != hp_cpu.astype(bool))[0]
    print(f"[padding {pad}] mismatches {len(mm)}  |d-pad| of mismatches:", np.abs(md_cpu[mm] - pad)[:10])

    # 4. 2-DOF
    m2, c2, v2 = two_dof.load_models()
    two_dof.add_object_collisions(m2, c2, v2)
    e2 = CollisionEngine(m2, c2)
    o2 = Oracle(m2, c2)
    q2 = np.random.default_rng(0).uniform(-3.14, 3.14, size=(10000, 2)).astype(np.float32)
    hc, mdc, _, _ = o2.check_states(q2.astype(np.float64), want_all=True, use_cull=False)
    hg, dg = e2.check_states(q2, return_distance=True)
    classify("2dof states", hg.astype(np.uint8), hc, mdc)
    fr = mdc > 0
    print("[2dof dist] max abs err:", np.abs(dg[fr] - mdc[fr]).max())
    hg_b = e2.check_states(q2)
    print("[2dof bool-mode] mismatches:", int((hg_b != hc.astype(bool)).sum()))

    # 5. edges
    E = int(os.environ.get("E", 3000))
    r1 = np.random.default_rng(1)
    A = r1.uniform(m.lowerPositionLimit, m.upperPositionLimit, size=(E, 9)).astype(np.float32)
    B = r1.uniform(m.lowerPositionLimit, m.upperPositionLimit, size=(E, 9)).astype(np.float32)
    # shorter edges too (more likely free)
    B[: E // 2] = A[: E // 2] + 0.15 * (B[: E // 2] - A[: E // 2])
    t0 = time.time()
    he_cpu, fb_cpu, evals, _ = orc.check_edges(A.astype(np.float64), B.astype(np.float64), 0.05)
    print("oracle edges s:", time.time() - t0, "states evaluated", evals, "P(edge hit)", he_cpu.mean())
    he_gpu = eng.check_edges(A, B, 0.05)
    print("[edges] mismatches:", int((he_gpu != he_cpu.astype(bool)).sum()))
    he2, fb_gpu = eng.check_edges(A, B, 0.05, return_first_bad=True)
    print("[edges first_bad] verdict mismatches:", int((he2 != he_cpu.astype(bool)).sum()))
    # oracle first_bad: -2 means q2 collides (reference checks q2 first); compare only where path sample found
    cmp = fb_cpu >= 0
    print("[edges first_bad] index mismatches (where oracle has an index):", int((fb_gpu[cmp] != fb_cpu[cmp]).sum()), "of", int(cmp.sum()))
    counts, offsets, states = eng.discretize(A[:100], B[:100], 0.05)
    ns = np.array([orc.segment_samples(A[i].astype(np.float64), B[i].astype(np.float64), 0.05) for i in range(100)])
    print("[discretize] counts equal:", bool((counts == ns).all()), "total", int(offsets[-1]))

    # 6. timing, device resident
    NT = int(os.environ.get("NT", 1 << 20))
    Qt = torch.as_tensor(rng.uniform(m.lowerPositionLimit, m.upperPositionLimit, size=(NT, 9)).astype(np.float32), device=dev)
    for _ in range(3):
        eng.check_states(Qt)
    torch.cuda.synchronize()
    eng.stats(reset=True)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    reps = 10
    for _ in range(reps):
        eng.check_states(Qt)
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / reps
    print(f"[timing] check_states N={NT}: {ms:.3f} ms -> {NT / ms * 1e3:.3e} states/s", eng.stats())
    # broadphase-only style split: time with events around internal phases is not available; print bins
    At = torch.as_tensor(A, device=dev).repeat(30, 1)
    Bt = torch.as_tensor(B, device=dev).repeat(30, 1)
    for _ in range(2):
        eng.check_edges(At, Bt, 0.05)
    torch.cuda.synchronize()
    ev0.record()
    eng.check_edges(At, Bt, 0.05)
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    print(f"[timing] check_edges E={At.shape[0]}: {ms:.3f} ms -> {At.shape[0] / ms * 1e3:.3e} edges/s")
    t0 = time.time()
    for _ in range(100):
        eng.check_states(Q32[:1])
    print("[latency] single-state host call us:", (time.time() - t0) / 100 * 1e6)
    fs = eng.sample_free_states(10000, seed=3, as_numpy=True)
    hs, _, _, _ = orc.check_states(fs.astype(np.float64))
    print("[sampler] got", len(fs), "draws", eng.last_sample_draws, "oracle says colliding:", int(hs.sum()))


if __name__ == "__main__":
    main()
