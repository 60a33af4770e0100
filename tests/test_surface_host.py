"""Surface semantics on the CPU: the host build of the device headers (tests/host_emu: pbp_surface.cuh with one
lane walking every warp-level loop, the product's own model tables from pbp_host_tables.h) against the
float64 oracle's exact triangle distances, on Panda items at and around contact.

What is checked is the chain the GPU runs per (state, pair) item -- GJK on the inner cores (boolean) or hulls
(distance), the lane-level hull / core runs and enclosure tests, the plane walk, the triangle stage -- with the
north_star bar: identical verdicts outside the 1e-6 m band, distances within 1e-5 m."""

import collections
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "host_emu"))
import emu  # noqa: E402

from conftest import make_panda, uniform_states  # noqa: E402


@pytest.fixture(scope="module")
def near_items():
    """(model, cmodel, {pair: [(T, triangle distance)]}) for pairs whose hulls are within 2 mm (or overlap)."""
    from oracle.oracle import Oracle

    model, cmodel, _ = make_panda()
    orc = Oracle(model, cmodel)
    orc_hull = Oracle(model, cmodel, mesh_semantics="hull")
    pairs = [(p.first, p.second) for p in cmodel.collisionPairs]
    q = uniform_states(model, 1200, 97).astype(np.float64)
    bs = orc_hull._arrays["geom_bsphere"]
    items = collections.defaultdict(list)
    for i in range(len(q)):
        _, oMg = orc_hull.fk(q[i])
        centre = [oMg[g][:9].reshape(3, 3) @ bs[g, :3] + oMg[g][9:] for g in range(len(bs))]
        for k, (ga, gb) in enumerate(pairs):
            if np.linalg.norm(centre[ga] - centre[gb]) > bs[ga, 3] + bs[gb, 3] + 0.002:
                continue
            if orc_hull.pair_distance_placed(oMg, ga, gb)[0] > 0.002:
                continue
            Ra, ta, Rb, tb = oMg[ga][:9].reshape(3, 3), oMg[ga][9:], oMg[gb][:9].reshape(3, 3), oMg[gb][9:]
            T = np.concatenate([(Ra.T @ Rb).reshape(9), Ra.T @ (tb - ta)])
            items[k].append((T, orc.soup_pair_distance(oMg, ga, gb)))
    assert sum(len(v) for v in items.values()) > 1000
    return model, cmodel, items


@pytest.mark.parametrize("padding", [0.0, 0.002])
def test_boolean_items_match_the_triangle_oracle(near_items, padding):
    model, cmodel, items = near_items
    paths = collections.Counter()
    for k, lst in items.items():
        T = np.array([x[0] for x in lst])
        ds = np.array([x[1] for x in lst])
        hit, _, path = emu.surface_items(model, cmodel, k, T, padding=padding, mode=0)
        bad = (hit.astype(bool) != (ds <= padding)) & (np.abs(ds - padding) > 1e-6)
        assert not bad.any(), (k, ds[bad][:4], hit[bad][:4], path[bad][:4])
        paths.update(int(x) for x in path)
    # most items are decided by the core GJK alone; the lane-level runs, the plane walk (enclosures) and the
    # triangle stage all occur in the sample
    assert paths[0] > 0.8 * sum(paths.values()) and paths[1] > 0 and paths[2] > 0


def test_pair_distances_match_the_triangle_oracle(near_items):
    model, cmodel, items = near_items
    worst, paths = 0.0, collections.Counter()
    for k, lst in items.items():
        T = np.array([x[0] for x in lst])
        ds = np.array([x[1] for x in lst])
        _, dist, path = emu.surface_items(model, cmodel, k, T, mode=2)
        worst = max(worst, float(np.abs(dist - ds).max()))
        paths.update(int(x) for x in path)
    assert worst < 1e-5, worst
    assert paths[3] > 20 and paths[2] > 5      # triangle stages and enclosures (distance of the enclosed shape to the surface)


def test_closed_form_triangle_distance_against_the_oracle():
    """The triangle stage's FP32 closed form (edge-edge, vertex-face, edge-through-face; pbp_surface.cuh) on the host
    build against the oracle's float64 routine: random pairs at the Panda's scale, near-parallel and coplanar
    pairs, pairs sharing an edge direction, intersecting pairs, slivers."""
    from oracle.oracle import tri_tri_distance

    rng = np.random.default_rng(11)
    A, B = [], []
    for i in range(6000):
        a = rng.uniform(-0.15, 0.15, size=(3, 3))
        kind = i % 6
        if kind == 0:      # anywhere
            b = rng.uniform(-0.15, 0.15, size=(3, 3)) + rng.uniform(-0.2, 0.2, size=3)
        elif kind == 1:    # nearly parallel, close
            n = np.cross(a[1] - a[0], a[2] - a[0]); n /= np.linalg.norm(n)
            b = a + rng.uniform(-0.03, 0.03, size=(3, 3)) + n * rng.uniform(1e-4, 5e-3) + rng.normal(scale=1e-4, size=(3, 3))
        elif kind == 2:    # coplanar, side by side
            u, v = a[1] - a[0], a[2] - a[0]
            b = a[0] + np.outer(rng.uniform(1.0, 2.0, size=3), u) + np.outer(rng.uniform(-1.0, 2.0, size=3), v)
        elif kind == 3:    # intersecting: b's first edge runs through a's interior
            c = a.mean(axis=0)
            n = np.cross(a[1] - a[0], a[2] - a[0]); n /= np.linalg.norm(n)
            b = np.array([c + 0.05 * n + rng.normal(scale=0.005, size=3), c - 0.05 * n + rng.normal(scale=0.005, size=3), c + rng.uniform(-0.2, 0.2, size=3)])
        elif kind == 4:    # sliver against a regular triangle
            b = np.array([a[0], a[0] + (a[1] - a[0]) * 1.0, a[0] + (a[1] - a[0]) * 0.5 + rng.normal(scale=1e-5, size=3)]) + rng.uniform(-0.05, 0.05, size=3)
        else:              # vertex of b just above a's face
            w = rng.dirichlet([1, 1, 1])
            n = np.cross(a[1] - a[0], a[2] - a[0]); n /= np.linalg.norm(n)
            p = w @ a + n * rng.uniform(1e-5, 2e-3)
            b = np.array([p, p + rng.uniform(0.02, 0.1) * n + rng.normal(scale=0.05, size=3), p + rng.uniform(0.02, 0.1) * n + rng.normal(scale=0.05, size=3)])
        A.append(a.astype(np.float32)), B.append(b.astype(np.float32))
    A, B = np.array(A), np.array(B)
    got = emu.tri_tri(A, B)
    ref = np.array([tri_tri_distance(a.astype(np.float64), b.astype(np.float64)) for a, b in zip(A, B)])
    err = np.abs(got - ref)
    assert err.max() < 2e-6, (err.max(), int(err.argmax()), got[err.argmax()], ref[err.argmax()])
    assert (ref == 0).sum() > 500 and ((got == 0) == (ref == 0))[ref > 1e-5].all()


@pytest.mark.parametrize("which,padding", [(5, 0.0), (5, 0.01), (10, 0.0)])
def test_ur_items_match_the_triangle_oracle(which, padding):
    """The same chain on the NON-convex UR meshes (inner cells 5-67 cm inside the hulls, the UR10 forearm without
    one: hull hits that prove nothing): the triangle stage decides nearly every near item; verdicts and distances
    against the oracle.  (This test found the distance bound that did not hold for a small box sitting in the
    shell between a link's surface and its inner cell.)"""
    from oracle.oracle import Oracle
    from pyroboplan_b200.models import ur5, ur10

    if which == 5:
        model, cmodel, _ = ur5.load_ur5_on_base_models()
        ur5.add_ur5_on_base_self_collisions(model, cmodel)
    else:
        model, cmodel, _ = ur10.load_ur10_on_base_models()
        ur10.add_ur10_on_base_self_collisions(model, cmodel)
    orc, orc_hull = Oracle(model, cmodel), Oracle(model, cmodel, mesh_semantics="hull")
    pairs = [(p.first, p.second) for p in cmodel.collisionPairs]
    lo, hi = np.maximum(model.lowerPositionLimit, -3.2), np.minimum(model.upperPositionLimit, 3.2)
    q = np.random.default_rng(23 + which).uniform(lo, hi, size=(160, model.nq))
    items = collections.defaultdict(list)
    for i in range(len(q)):
        _, oMg = orc_hull.fk(q[i])
        for k, (ga, gb) in enumerate(pairs):
            if orc_hull.pair_distance_placed(oMg, ga, gb)[0] > padding + 0.002:
                continue
            Ra, ta, Rb, tb = oMg[ga][:9].reshape(3, 3), oMg[ga][9:], oMg[gb][:9].reshape(3, 3), oMg[gb][9:]
            items[k].append((np.concatenate([(Ra.T @ Rb).reshape(9), Ra.T @ (tb - ta)]), orc.soup_pair_distance(oMg, ga, gb)))
    total = sum(len(v) for v in items.values())
    assert total > 150
    paths, worst, apart = collections.Counter(), 0.0, 0
    for k, lst in items.items():
        T = np.array([x[0] for x in lst])
        ds = np.array([x[1] for x in lst])
        hit, _, path = emu.surface_items(model, cmodel, k, T, padding=padding, mode=0)
        bad = (hit.astype(bool) != (ds <= padding)) & (np.abs(ds - padding) > 1e-6)
        assert not bad.any(), (k, ds[bad][:4], hit[bad][:4], path[bad][:4])
        apart += int((ds > padding + 1e-6).sum())          # hull-near items whose SURFACES are apart
        paths.update(int(x) for x in path)
        if padding == 0.0:
            _, dist, _ = emu.surface_items(model, cmodel, k, T, mode=2)
            worst = max(worst, float(np.abs(dist - ds).max()))
    assert worst < 1e-5, worst
    assert paths[3] > 0.3 * total and apart > 10       # the triangles decide, and hulls and surfaces disagree often


@pytest.mark.parametrize("seed", [1, 2])
def test_nested_shells_on_the_host(seed):
    """Both enclosure roles (the enclosed body listed first or second), a mesh, a sphere and a box as the enclosed
    shape, with and without padding: the synthetic model of tests/test_gpu_surface.py through the host build."""
    from oracle.oracle import Oracle
    from test_gpu_surface import nested_shell_model

    model, cmodel = nested_shell_model(seed)
    orc = Oracle(model, cmodel)
    pairs = [(p.first, p.second) for p in cmodel.collisionPairs]
    q = uniform_states(model, 4000, 60 + seed).astype(np.float64)
    enclosed = 0
    for padding in (0.0, 0.02):
        for k, (ga, gb) in enumerate(pairs):
            Ts, ds = [], []
            for i in range(len(q)):
                _, oMg = orc.fk(q[i])
                Ra, ta, Rb, tb = oMg[ga][:9].reshape(3, 3), oMg[ga][9:], oMg[gb][:9].reshape(3, 3), oMg[gb][9:]
                Ts.append(np.concatenate([(Ra.T @ Rb).reshape(9), Ra.T @ (tb - ta)]))
                ds.append(orc.soup_pair_distance(oMg, ga, gb))
            Ts, ds = np.array(Ts), np.array(ds)
            hit, _, path = emu.surface_items(model, cmodel, k, Ts, padding=padding, mode=0)
            bad = (hit.astype(bool) != (ds <= padding)) & (np.abs(ds - padding) > 2e-6)
            assert not bad.any(), (k, padding, ds[bad][:4], hit[bad][:4], path[bad][:4])
            if padding == 0.0:
                _, dist, path = emu.surface_items(model, cmodel, k, Ts, mode=2)
                assert np.abs(dist - ds).max() < 1e-5, (k, float(np.abs(dist - ds).max()))
                enclosed += int(((path == 2) & (ds > 1e-6)).sum())
    assert enclosed >= 3      # bodies wholly inside a shell, measured by the plane walk


def _random_rotation(rng):
    q = rng.normal(size=4)
    w, x, y, z = q / np.linalg.norm(q)
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)], [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


@pytest.mark.parametrize("robot,n_pairs,n_place", [("panda", 24, 30), ("ur5", 12, 15)])
def test_random_relative_placements(robot, n_pairs, n_place):
    """Placements no kinematic chain would produce -- the second shape anywhere from the centre of the first to
    just out of reach, any orientation: deep enclosures, partial overlaps, grazing contacts.  Verdicts at two paddings
    and the pair's surface distance against the oracle's triangle distance, for a random subset of the pairs."""
    from oracle.oracle import Oracle
    from pyroboplan_b200.engine import flatten_model
    from pyroboplan_b200.models import ur5

    if robot == "panda":
        model, cmodel, _ = make_panda()
    else:
        model, cmodel, _ = ur5.load_ur5_on_base_models()
        ur5.add_ur5_on_base_self_collisions(model, cmodel)
    orc = Oracle(model, cmodel)
    arrays, _ = flatten_model(model, cmodel)
    pairs = [(p.first, p.second) for p in cmodel.collisionPairs]
    rng = np.random.default_rng(5)
    ng = len(cmodel.geometryObjects)
    worst = 0.0
    for k in rng.choice(len(pairs), size=n_pairs, replace=False):
        ga, gb = pairs[k]
        ca, cb = arrays["geom_outer"][ga, :3].astype(float), arrays["geom_outer"][gb, :3].astype(float)
        reach = float(arrays["geom_outer"][ga, 3] + arrays["geom_outer"][gb, 3])
        Ts, ds = [], []
        for _ in range(n_place):
            R = _random_rotation(rng)
            d = rng.normal(size=3)
            t = ca + d / np.linalg.norm(d) * reach * rng.uniform(0, 1) * np.sqrt(rng.uniform(0, 1)) - R @ cb
            oMg = np.zeros((ng, 12))
            oMg[:, [0, 4, 8]] = 1.0
            oMg[gb, :9], oMg[gb, 9:] = R.reshape(9), t
            Ts.append(np.concatenate([R.reshape(9), t]))
            ds.append(orc.soup_pair_distance(oMg, ga, gb))
        Ts, ds = np.array(Ts), np.array(ds)
        for pad in (0.0, 0.004):
            hit, _, path = emu.surface_items(model, cmodel, int(k), Ts, padding=pad, mode=0)
            bad = (hit.astype(bool) != (ds <= pad)) & (np.abs(ds - pad) > 1e-6)
            assert not bad.any(), (int(k), pad, ds[bad][:3], hit[bad][:3], path[bad][:3])
        _, dist, _ = emu.surface_items(model, cmodel, int(k), Ts, mode=2)
        worst = max(worst, float(np.abs(dist - ds).max()))
    assert worst < 1e-5, worst


@pytest.mark.parametrize("robot", ["panda", "ur5"])
def test_inner_ball_certificate_is_sound(robot):
    """k_item_setup publishes a hit when two inner balls are within the padding and the centre-line test rules an
    enclosure out.  Soundness on random placements biased towards overlap: wherever the certificate fires the oracle's
    exact surface distance is within the padding (never a false hit), and it fires often enough to matter."""
    from oracle.oracle import Oracle
    from pyroboplan_b200.engine import flatten_model
    from pyroboplan_b200.models import ur5

    if robot == "panda":
        model, cmodel, _ = make_panda()
    else:
        model, cmodel, _ = ur5.load_ur5_on_base_models()
        ur5.add_ur5_on_base_self_collisions(model, cmodel)
    orc = Oracle(model, cmodel)
    arrays, _ = flatten_model(model, cmodel)
    pairs = [(p.first, p.second) for p in cmodel.collisionPairs]
    rng = np.random.default_rng(9)
    ng = len(cmodel.geometryObjects)
    fired = colliding = 0
    for k in range(len(pairs)):
        ga, gb = pairs[k]
        ca, cb = arrays["geom_outer"][ga, :3].astype(float), arrays["geom_outer"][gb, :3].astype(float)
        reach = float(arrays["geom_outer"][ga, 3] + arrays["geom_outer"][gb, 3])
        Ts, oMgs = [], []
        for _ in range(40 if robot == "panda" else 25):
            R = _random_rotation(rng)
            d = rng.normal(size=3)
            t = ca + d / np.linalg.norm(d) * 0.8 * reach * rng.uniform(0, 1) - R @ cb
            oMg = np.zeros((ng, 12))
            oMg[:, [0, 4, 8]] = 1.0
            oMg[gb, :9], oMg[gb, 9:] = R.reshape(9), t
            Ts.append(np.concatenate([R.reshape(9), t]))
            oMgs.append(oMg)
        for pad in (0.0, 0.01):
            balls, pokes = emu.ball_certificate(model, cmodel, k, np.array(Ts), padding=pad)
            for i in np.nonzero(balls & pokes)[0]:
                ds = orc.soup_pair_distance(oMgs[i], ga, gb, stop_at=pad)
                assert ds <= pad + 1e-6, (k, pad, ds)
                fired += 1
            if pad == 0.0:
                colliding += sum(orc.soup_pair_distance(o, ga, gb, stop_at=0.0) <= 0.0 for o in oMgs)
    assert fired > 100 and colliding > 0
    print(f"[{robot}] certificate fired on {fired} placements; {colliding} colliding placements at padding 0")
