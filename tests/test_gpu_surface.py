"""Surface semantics of mesh geometries on the GPU (pbp_model_desc.planes / pair_enclose).

Pinocchio loads URDF meshes as coal BVHModels -- surfaces -- so a shape wholly inside a mesh without
touching its surface does not collide with it (reference call site: pinocchio.buildModelsFromUrdf,
src/pyroboplan/models/panda.py:27).  About 1 in 1000 uniform Panda states differs from the
convex-hull verdict that way (a finger inside link 0 / 1 / 2 / 5; the hand pairs are disabled by
the SRDF).  The oracle decides with exact triangle tests; the engine with the enclosure test of
k_surface_refine."""

import numpy as np
import pytest

from conftest import make_panda, uniform_states

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def setup():
    from oracle.oracle import Oracle
    from pyroboplan_b200.engine import CollisionEngine

    model, cmodel, _ = make_panda()
    eng = CollisionEngine(model, cmodel, device=0)
    eng_hull = CollisionEngine(model, cmodel, device=0, mesh_semantics="hull")
    yield model, cmodel, eng, eng_hull, Oracle(model, cmodel), Oracle(model, cmodel, mesh_semantics="hull")
    eng.close()
    eng_hull.close()


def _mismatch_far(h, ref, dref):
    bad = np.nonzero(np.asarray(h, dtype=bool) != ref.astype(bool))[0]
    return [int(i) for i in bad if abs(dref[i]) > 1e-6]


def test_state_verdicts_every_path(setup):
    import torch

    model, cmodel, eng, eng_hull, orc, orc_hull = setup
    q = uniform_states(model, 300000, 41)
    qd = torch.as_tensor(q, device="cuda:0")
    ref, dref, _, _ = orc.check_states(q.astype(np.float64), want_all=True)
    ref_hull = orc_hull.check_states(q.astype(np.float64))[0]
    flips = int((ref_hull.astype(bool) & ~ref.astype(bool)).sum())
    assert 150 < flips < 600                      # the enclosure cases are in the sample
    assert not _mismatch_far(eng.check_states(qd).cpu().numpy(), ref, dref)                          # three-kernel pipeline + refine
    hit_d, dist = eng.check_states(qd, return_distance=True)
    assert not _mismatch_far(hit_d.cpu().numpy(), ref, dref)                                         # distance mode
    free = (~ref.astype(bool)) & (~hit_d.cpu().numpy().astype(bool))
    assert np.abs(dist.cpu().numpy()[free] - dref[free]).max() < 1e-5                                # triangle distances, enclosed shapes included
    small = np.concatenate([eng.check_states(qd[i:i + 300]).cpu().numpy() for i in range(0, 30000, 300)])
    assert not _mismatch_far(small, ref[:30000], dref[:30000])                                       # fused small-batch kernel
    hit_p, pair = eng.check_states(qd[:100000], return_pair=True)
    assert not _mismatch_far(hit_p.cpu().numpy(), ref[:100000], dref[:100000])                       # all-pairs mode
    assert not _mismatch_far(eng.check_states(q[:50000]), ref[:50000], dref[:50000])                 # host buffers
    # hull semantics stay available and agree with the hull oracle
    assert not _mismatch_far(eng_hull.check_states(qd).cpu().numpy(), ref_hull, orc_hull.check_states(q.astype(np.float64), want_all=True)[1])


def test_enclosed_states_are_free_with_positive_clearance(setup):
    import torch

    model, cmodel, eng, eng_hull, orc, orc_hull = setup
    q = uniform_states(model, 200000, 42)
    ref, dref, _, _ = orc.check_states(q.astype(np.float64), want_all=True)
    ref_hull = orc_hull.check_states(q.astype(np.float64))[0].astype(bool)
    enclosed = np.nonzero(ref_hull & ~ref.astype(bool) & (dref > 1e-3))[0]
    assert len(enclosed) > 50
    qe = torch.as_tensor(q[enclosed], device="cuda:0")
    hit, dist = eng.check_states(qe, return_distance=True)
    assert not hit.any().item() and not eng.check_states(qe).any().item()
    assert eng_hull.check_states(qe).all().item()
    np.testing.assert_allclose(dist.cpu().numpy(), dref[enclosed], atol=5e-4)
    # one at a time through the drop-in function (fused kernel, host path)
    from pyroboplan_b200.core.utils import check_collisions_at_state

    for i in enclosed[:10]:
        assert not check_collisions_at_state(model, cmodel, q[i])


def test_padding_and_edges(setup):
    import torch

    model, cmodel, eng, eng_hull, orc, orc_hull = setup
    q = uniform_states(model, 60000, 43)
    qd = torch.as_tensor(q, device="cuda:0")
    for pad in (0.004, 0.02):
        ref, dref, _, _ = orc.check_states(q.astype(np.float64), padding=pad, want_all=True)
        h = eng.check_states(qd, padding=pad).cpu().numpy().astype(bool)
        bad = np.nonzero(h != ref.astype(bool))[0]
        # a verdict may differ only within FP32 resolution of the padding
        assert all(abs(dref[i] - pad) <= 1e-6 for i in bad), [(int(i), float(dref[i])) for i in bad]
    a = uniform_states(model, 4000, 44)
    b = (a + 0.25 * (uniform_states(model, 4000, 45) - a)).astype(np.float32)
    ref_e, _, _, _ = orc.check_edges(a.astype(np.float64), b.astype(np.float64), 0.05)
    got = eng.check_edges(torch.as_tensor(a, device="cuda:0"), torch.as_tensor(b, device="cuda:0"), 0.05).cpu().numpy().astype(bool)
    assert (got != ref_e.astype(bool)).sum() <= 1          # an edge sample within 1e-6 m of contact
    # short edges that start at an enclosed state: free to leave under surface semantics unless they cross the mesh
    ref_s = orc.check_states(q.astype(np.float64))[0].astype(bool)
    ref_h = orc_hull.check_states(q.astype(np.float64))[0].astype(bool)
    enclosed = q[ref_h & ~ref_s]
    assert len(enclosed) > 20
    nudged = (enclosed + np.float32(0.01) * (uniform_states(model, len(enclosed), 46) - enclosed)).astype(np.float32)
    ref_n, _, _, _ = orc.check_edges(nudged.astype(np.float64), enclosed.astype(np.float64), 0.05)
    got_n = eng.check_edges(torch.as_tensor(nudged, device="cuda:0"), torch.as_tensor(enclosed, device="cuda:0"), 0.05).cpu().numpy().astype(bool)
    assert (got_n == ref_n.astype(bool)).all() and not ref_n.all()
    assert eng_hull.check_edges(torch.as_tensor(nudged, device="cuda:0"), torch.as_tensor(enclosed, device="cuda:0"), 0.05).all().item()


def nested_shell_model(seed):
    """A synthetic model for the enclosure logic beyond the Panda: big convex MESH shells (exactly convex
    triangulated surfaces, no dents) and small bodies -- a mesh, a sphere, a box -- that move in and out of
    them.  Half of the pairs list the small body first, so both enclosure roles (pair_enclose bits 0 and 1)
    and a primitive as the enclosed shape are exercised."""
    from scipy.spatial import ConvexHull

    from pyroboplan_b200.model import (Box, CollisionPair, Convex, GeometryModel, GeometryObject, Model, SE3, Sphere,
                                       JOINT_PRISMATIC, JOINT_REVOLUTE)

    rng = np.random.default_rng(seed)

    def mesh(scale, n):
        pts = rng.normal(size=(n, 3))
        pts = pts / np.linalg.norm(pts, axis=1, keepdims=True) * rng.uniform(0.7, 1.0, size=(n, 1)) * scale
        pts = pts[ConvexHull(pts).vertices]
        return Convex(pts, triangles=ConvexHull(pts).simplices)

    model = Model(f"shells{seed}")
    cm = GeometryModel()
    I3 = np.eye(3)
    # body 1: three prismatic joints and a revolute one carry the small shapes through the static shell
    j = 0
    for k in range(3):
        j = model.addJoint(j, JOINT_PRISMATIC, I3[k], SE3(I3, np.zeros(3)), f"p{k}", -0.45, 0.45)
    j_small = model.addJoint(j, JOINT_REVOLUTE, rng.normal(size=3), SE3(I3, np.zeros(3)), "r_small", -3.0, 3.0)
    # body 2: a moving shell that sweeps over static small shapes
    j = 0
    for k in range(3):
        j = model.addJoint(j, JOINT_PRISMATIC, I3[k], SE3(I3, np.array([2.0, 0.0, 0.0]) if k == 0 else np.zeros(3)), f"q{k}", -0.45, 0.45)
    j_shell = model.addJoint(j, JOINT_REVOLUTE, rng.normal(size=3), SE3(I3, np.zeros(3)), "r_shell", -3.0, 3.0)

    def add(name, joint, shape, t=(0.0, 0.0, 0.0)):
        return cm.addGeometryObject(GeometryObject(name, joint, SE3(I3, np.asarray(t, dtype=float)), shape))

    shell_static = add("shell_static", 0, mesh(0.32, 60))
    small_mesh = add("small_mesh", j_small, mesh(0.06, 20))
    small_sphere = add("small_sphere", j_small, Sphere(0.04), (0.0, 0.0, 0.0))
    small_box = add("small_box", j_small, Box(0.07, 0.05, 0.04))
    shell_moving = add("shell_moving", j_shell, mesh(0.32, 60))
    fixed_mesh = add("fixed_mesh", 0, mesh(0.06, 20), (2.0, 0.0, 0.0))
    fixed_sphere = add("fixed_sphere", 0, Sphere(0.04), (2.0, 0.1, 0.0))
    for a, b in ((shell_static, small_mesh), (shell_static, small_sphere), (shell_static, small_box),   # big first: bit 0
                 (fixed_mesh, shell_moving), (fixed_sphere, shell_moving)):                                # small first: bit 1
        cm.addCollisionPair(CollisionPair(a, b))
    return model, cm


@pytest.mark.parametrize("seed", [1, 2])
def test_nested_shells_both_enclosure_roles(seed):
    import torch

    from oracle.oracle import Oracle
    from pyroboplan_b200.engine import CollisionEngine, flatten_model

    model, cm = nested_shell_model(seed)
    arrays, sizes = flatten_model(model, cm)
    assert sorted(int(f) for f in arrays["pair_enclose"]) == [1, 1, 1, 2, 2]
    eng = CollisionEngine(model, cm, device=0)
    orc, orc_hull = Oracle(model, cm), Oracle(model, cm, mesh_semantics="hull")
    q = uniform_states(model, 120000, 50 + seed)
    qd = torch.as_tensor(q, device="cuda:0")
    ref, dref, fp_ref, _ = orc.check_states(q.astype(np.float64), want_all=True, use_cull=False)
    ref_hull = orc_hull.check_states(q.astype(np.float64))[0].astype(bool)
    enclosed = ref_hull & ~ref.astype(bool)
    assert enclosed.sum() > 300                                       # plenty of states with a body wholly inside a shell
    assert not _mismatch_far(eng.check_states(qd).cpu().numpy(), ref, dref)
    hit_d, dist = eng.check_states(qd, return_distance=True)
    assert not _mismatch_far(hit_d.cpu().numpy(), ref, dref)
    free = (~ref.astype(bool)) & (~hit_d.cpu().numpy().astype(bool))
    assert np.abs(dist.cpu().numpy()[free] - dref[free]).max() < 1e-5   # exactly convex meshes: plane clearance = triangle distance
    small = np.concatenate([eng.check_states(qd[i:i + 200]).cpu().numpy() for i in range(0, 20000, 200)])
    assert not _mismatch_far(small, ref[:20000], dref[:20000])
    hit_p, pair = eng.check_states(qd[:50000], return_pair=True)
    assert not _mismatch_far(hit_p.cpu().numpy(), ref[:50000], dref[:50000])
    clear = np.abs(dref[:50000]) > 1e-6
    assert (pair.cpu().numpy()[clear] == fp_ref[:50000][clear]).all()
    # padding: an enclosed body closer than the padding to the shell's surface collides
    for pad in (0.01, 0.03):
        refp, drefp, _, _ = orc.check_states(q[:60000].astype(np.float64), padding=pad, want_all=True, use_cull=False)
        h = eng.check_states(qd[:60000], padding=pad).cpu().numpy().astype(bool)
        bad = np.nonzero(h != refp.astype(bool))[0]
        assert all(abs(drefp[i] - pad) <= 2e-6 for i in bad), [(int(i), float(drefp[i])) for i in bad[:5]]
    eng.close()


def test_sampler_accepts_enclosed_states_as_the_reference_does(setup):
    """Rejection sampling (get_random_collision_free_state's batched sibling) runs on the same verdicts:
    every accepted state is free for the triangle oracle, and a few of them are the enclosure states a
    convex-hull engine would have rejected."""
    model, cmodel, eng, eng_hull, orc, orc_hull = setup
    q = eng.sample_free_states(150000, seed=5, as_numpy=True)
    assert q.shape == (150000, model.nq)
    ref, dref, _, _ = orc.check_states(q.astype(np.float64), want_all=True)
    assert not [i for i in np.nonzero(ref)[0] if abs(dref[i]) > 1e-6]
    assert orc_hull.check_states(q.astype(np.float64))[0].sum() > 50
