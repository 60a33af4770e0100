"""The C-ABI shared library: loads without a GPU, exports every symbol the header declares,
and refuses to create an engine without a CUDA device (no CPU fallback)."""

import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT, make_two_dof
from pyroboplan_b200 import engine as eng_mod


def _header_functions():
    text = open(os.path.join(ROOT, "include", "pbp_engine.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pbp_[a-z0-9_]+)\s*\(", text)))


def test_header_matches_python_binding():
    assert _header_functions() == sorted(eng_mod.EXPORTED_SYMBOLS)


def test_library_exports_every_declared_symbol():
    lib = eng_mod.load_library()
    for name in _header_functions():
        assert hasattr(lib, name), f"libpbp_engine.so does not export {name}"
    assert lib.pbp_abi_version() == eng_mod.PBP_ABI_VERSION


def test_status_codes_in_header():
    text = open(os.path.join(ROOT, "include", "pbp_engine.h")).read()
    for name, val in (("PBP_OK", 0), ("PBP_ERR_ARG", -1), ("PBP_ERR_CUDA", -2), ("PBP_ERR_CAPACITY", -3), ("PBP_ERR_ALLOC", -4)):
        assert re.search(rf"{name}\s*=\s*{val}\b", text)


def test_no_cpu_fallback_without_device():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    model, cmodel, _ = make_two_dof()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        eng_mod.CollisionEngine(model, cmodel)
    # and at the C level: creation fails with PBP_ERR_CUDA and a message, no compute is attempted
    lib = eng_mod.load_library()
    arrays, sizes = eng_mod.flatten_model(model, cmodel)
    desc = eng_mod._ModelDesc()
    for k, v in sizes.items():
        setattr(desc, k, v)
    for k, v in arrays.items():
        setattr(desc, k, v.ctypes.data)
    handle = ctypes.c_void_p()
    rc = lib.pbp_engine_create(ctypes.byref(desc), 0, 0, ctypes.byref(handle))
    assert rc == -2 and not handle.value
    assert b"no CPU fallback" in lib.pbp_last_error()


def test_argument_validation_without_device():
    lib = eng_mod.load_library()
    assert lib.pbp_engine_create(None, 0, 0, None) == -1
    assert lib.pbp_get_stats(None, None, 0) == -1
    assert lib.pbp_check_states(None, None, 1, ctypes.c_float(0.0), None, None, None, None) == -1
    assert b"NULL" in lib.pbp_last_error()
    # the entry points added for the IK / distance / k-NN rows validate before touching the device too
    assert lib.pbp_knn(None, None, 4, 3, ctypes.c_double(1.0), 0, None, None, None) == -1
    assert lib.pbp_compute_distances(None, None, 4, ctypes.c_float(0.0), None, None, None, None, None) == -1
    assert lib.pbp_collision_nullspace(None, None, 4, ctypes.c_float(0.05), ctypes.c_float(1.0), None, None) == -1
    assert lib.pbp_ik_iterate(None, 0, None, None, 4, None, ctypes.c_uint64(0), None, None, None, None, None, None) == -1
    assert lib.pbp_jit_broadphase_source(None, None, 0, None) == -1


def test_flatten_model_is_conservative():
    """Enclosing spheres really enclose and inscribed spheres are really inside, in float32."""
    from conftest import make_panda

    model, cmodel, _ = make_panda()
    arrays, sizes = eng_mod.flatten_model(model, cmodel)
    assert {k: sizes[k] for k in ("nq", "njoints", "ngeoms", "npairs", "nverts", "nmverts", "ntris")} == dict(
        nq=9, njoints=10, ngeoms=16, npairs=74, nverts=1204, nmverts=1204, ntris=2364)
    # planes of the exact inner cores: every hull face (= 2364 triangles' worth) plus the mesh triangles below the hulls
    assert sizes["nplanes"] > 2364 and int(arrays["geom_plane_hull_cnt"].sum()) == 2364
    assert 1000 < sizes["ncore"] < 1.2 * sizes["nverts"]     # the simplified cores keep the GJK tables' size
    for g, obj in enumerate(cmodel.geometryObjects):
        s = obj.geometry
        oc, orad = arrays["geom_outer"][g][:3].astype(np.float64), float(arrays["geom_outer"][g][3])
        ic, irad = arrays["geom_inner"][g][:3].astype(np.float64), float(arrays["geom_inner"][g][3])
        if type(s).__name__ == "Convex":
            off, cnt = arrays["geom_vert_off"][g], arrays["geom_vert_cnt"][g]
            v = arrays["verts"][off : off + cnt].astype(np.float64)
            assert np.all(np.linalg.norm(v - oc, axis=1) <= orad)
            assert np.all(s.planes[:, :3] @ ic + s.planes[:, 3] <= -irad)
            if arrays["geom_tri_cnt"][g]:
                # surface geometry: the "certainly colliding" ball lies inside every mesh triangle's inner half-space,
                # and so do the core's vertices (the core is inside the mesh); the hull is within core_eps of the core
                mo, to, tc = arrays["geom_mvert_off"][g], arrays["geom_tri_off"][g], arrays["geom_tri_cnt"][g]
                tri = arrays["mesh_verts"].astype(np.float64)[mo + arrays["tris"][to : to + tc]]
                n = np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0])
                n /= np.linalg.norm(n, axis=1, keepdims=True)
                w = (n * tri[:, 0]).sum(axis=1)
                flip = n @ ic - w > 0
                n[flip] *= -1
                w[flip] *= -1
                assert np.all(n @ ic - w <= -irad + 1e-9)
                co, cc = arrays["geom_core_off"][g], arrays["geom_core_cnt"][g]
                core = arrays["core_verts"][co : co + cc].astype(np.float64)
                assert cc >= 4 and (core @ n.T - w).max() < 5e-8
                assert 0 < arrays["geom_core_eps"][g] < 2e-3
        assert irad <= orad
