"""ctypes driver of the host build of the engine's FP32 GJK (test tooling, see gjk_emu.cpp)."""

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libgjk_emu.so")
_SRC = os.path.join(_HERE, "gjk_emu.cpp")
_HDRS = [os.path.join(_HERE, "..", "..", "pyroboplan_b200", "csrc", f) for f in ("pbp_gjk.cuh", "pbp_model.cuh", "pbp_supmap.h", "pbp_witness.cuh", "pbp_surface.cuh", "pbp_host_tables.h")]
_lib = None
_flat_cache = {}


def cuda_include():
    for d in ("/usr/local/cuda/include", os.path.join(os.environ.get("CUDA_HOME", ""), "include")):
        if os.path.exists(os.path.join(d, "cuda_runtime.h")):
            return d
    return None


def load():
    global _lib
    if _lib is None:
        newest = max(os.path.getmtime(p) for p in [_SRC] + _HDRS)
        if not os.path.exists(_LIB) or os.path.getmtime(_LIB) < newest:
            inc = cuda_include()
            if inc is None:
                raise RuntimeError("CUDA headers not found")
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", f"-I{inc}", "-o", _LIB, _SRC])
        _lib = ctypes.CDLL(_LIB)
        vp = ctypes.c_void_p
        _lib.emu_gjk_batch.argtypes = [ctypes.c_int, vp, vp, ctypes.c_int, ctypes.c_int, vp, vp, ctypes.c_int, vp, vp,
                                       ctypes.c_int64, ctypes.c_float, ctypes.c_float, ctypes.c_int, ctypes.c_int, vp, vp, vp]
        _lib.emu_gjk_batch.restype = None
        _lib.emu_gjk_witness_batch.argtypes = [ctypes.c_int, vp, vp, ctypes.c_int, ctypes.c_int, vp, vp, ctypes.c_int, vp, vp,
                                               ctypes.c_int64, ctypes.c_float, ctypes.c_float, ctypes.c_float, vp]
        _lib.emu_gjk_witness_batch.restype = None
        _lib.emu_supmap_check.argtypes = [vp, ctypes.c_int, vp, ctypes.c_int64, vp]
        _lib.emu_supmap_check.restype = None
        _lib.emu_surface_items.argtypes = [vp, ctypes.c_int, vp, ctypes.c_int64, ctypes.c_float, ctypes.c_int, vp, vp, vp]
        _lib.emu_last_error.restype = ctypes.c_char_p
        _lib.emu_ball_certificate.argtypes = [vp, ctypes.c_int, vp, ctypes.c_int64, ctypes.c_float, vp]
        _lib.emu_tri_tri.argtypes = [vp, vp, ctypes.c_int64, vp]
        _lib.emu_tri_tri.restype = None
    return _lib


def padded_verts(arrays, g):
    """float4-padded vertex table of geometry g, as the engine builds it (last vertex repeated)."""
    off, cnt = int(arrays["geom_vert_off"][g]), int(arrays["geom_vert_cnt"][g])
    if cnt == 0:
        return np.zeros((4, 4), dtype=np.float32)
    v = np.zeros((cnt, 4), dtype=np.float32)
    v[:, :3] = arrays["verts"][off : off + cnt]
    while len(v) % 4:
        v = np.vstack([v, v[-1:]])
    return np.ascontiguousarray(v)


def supmap_check(arrays, g, dirs):
    """(list entries, max candidates per cell, #directions where the map disagrees with a full scan)."""
    lib = load()
    v = padded_verts(arrays, g)
    dirs = np.ascontiguousarray(dirs, dtype=np.float32).reshape(-1, 3)
    out = np.zeros(3, dtype=np.int64)
    lib.emu_supmap_check(ctypes.c_void_p(v.ctypes.data), len(v), ctypes.c_void_p(dirs.ctypes.data), len(dirs), ctypes.c_void_p(out.ctypes.data))
    return tuple(int(x) for x in out)


def gjk_batch(arrays, ga, gb, T, v0, margin, want_dist=False, bound=np.inf, use_maps=True):
    """Run the FP32 GJK on n items of the pair (ga, gb). Returns (hit, dist, iters)."""
    lib = load()
    T = np.ascontiguousarray(T, dtype=np.float32).reshape(-1, 12)
    v0 = np.ascontiguousarray(v0, dtype=np.float32).reshape(-1, 3)
    n = len(T)
    va, vb = padded_verts(arrays, ga), padded_verts(arrays, gb)
    pa = np.ascontiguousarray(arrays["geom_params"][ga][:3], dtype=np.float32)
    pb = np.ascontiguousarray(arrays["geom_params"][gb][:3], dtype=np.float32)
    hit = np.zeros(n, dtype=np.int32)
    dist = np.zeros(n, dtype=np.float32)
    iters = np.zeros(n, dtype=np.int32)
    p = lambda a: ctypes.c_void_p(a.ctypes.data)
    lib.emu_gjk_batch(int(arrays["geom_shape"][ga]), p(pa), p(va), len(va) if arrays["geom_vert_cnt"][ga] else 0,
                      int(arrays["geom_shape"][gb]), p(pb), p(vb), len(vb) if arrays["geom_vert_cnt"][gb] else 0,
                      p(T), p(v0), n, float(margin), float(bound), int(want_dist), int(use_maps), p(hit), p(dist), p(iters))
    return hit, dist, iters


def relative_placements(oMg, ga, gb):
    """T = oMg[a]^-1 * oMg[b] for a batch of placements oMg [n, ngeoms, 12] (float64 in)."""
    Ra = oMg[:, ga, :9].reshape(-1, 3, 3)
    ta = oMg[:, ga, 9:]
    Rb = oMg[:, gb, :9].reshape(-1, 3, 3)
    tb = oMg[:, gb, 9:]
    R = np.einsum("nji,njk->nik", Ra, Rb)
    t = np.einsum("nji,nj->ni", Ra, tb - ta)
    return np.concatenate([R.reshape(-1, 9), t], axis=1)


def gjk_witness_batch(arrays, ga, gb, T, v0, cutoff=np.inf):
    """FP32 GJK with witness tracking on n items of the pair (ga, gb).

    Returns dict(dist, hit, far, pa, pb, n): witness points / normal in geometry ga's frame."""
    lib = load()
    T = np.ascontiguousarray(T, dtype=np.float32).reshape(-1, 12)
    v0 = np.ascontiguousarray(v0, dtype=np.float32).reshape(-1, 3)
    n = len(T)
    va, vb = padded_verts(arrays, ga), padded_verts(arrays, gb)
    pa = np.ascontiguousarray(arrays["geom_params"][ga][:3], dtype=np.float32)
    pb = np.ascontiguousarray(arrays["geom_params"][gb][:3], dtype=np.float32)
    ra = float(arrays["geom_params"][ga][0]) if arrays["geom_shape"][ga] in (0, 3) else 0.0
    rb = float(arrays["geom_params"][gb][0]) if arrays["geom_shape"][gb] in (0, 3) else 0.0
    out = np.zeros((n, 12), dtype=np.float32)
    p = lambda a: ctypes.c_void_p(a.ctypes.data)
    lib.emu_gjk_witness_batch(int(arrays["geom_shape"][ga]), p(pa), p(va), len(va) if arrays["geom_vert_cnt"][ga] else 0,
                              int(arrays["geom_shape"][gb]), p(pb), p(vb), len(vb) if arrays["geom_vert_cnt"][gb] else 0,
                              p(T), p(v0), n, ra, rb, float(cutoff), p(out))
    return dict(dist=out[:, 0], hit=out[:, 1].astype(bool), far=out[:, 2].astype(bool), pa=out[:, 3:6], pb=out[:, 6:9], n=out[:, 9:12])


def surface_items(model, collision_model, pair, T, padding=0.0, mode=0):
    """The engine's answer for placements T [n, 12] (b in a's frame) of the caller's pair index `pair`,
    evaluated by the host build of pbp_surface.cuh.  Returns (hit [n], dist [n], path [n])."""
    from pyroboplan_b200 import engine

    lib = load()
    key = (id(model), id(collision_model), len(collision_model.collisionPairs))
    if _flat_cache.get("key") != key:
        arrays, sizes = engine.flatten_model(model, collision_model)
        desc = engine._ModelDesc()
        for k, v in sizes.items():
            setattr(desc, k, v)
        for k, v in arrays.items():
            setattr(desc, k, v.ctypes.data)
        _flat_cache.update(key=key, arrays=arrays, desc=desc)
    desc = _flat_cache["desc"]
    T = np.ascontiguousarray(T, dtype=np.float32).reshape(-1, 12)
    n = len(T)
    hit, dist, path = np.zeros(n, np.int32), np.zeros(n, np.float32), np.zeros(n, np.int32)
    p = lambda a: ctypes.c_void_p(a.ctypes.data)
    rc = lib.emu_surface_items(ctypes.byref(desc), int(pair), p(T), n, float(padding), int(mode), p(hit), p(dist), p(path))
    if rc:
        raise RuntimeError(lib.emu_last_error().decode())
    return hit, dist, path


def _desc_for(model, collision_model):
    from pyroboplan_b200 import engine

    key = (id(model), id(collision_model), len(collision_model.collisionPairs))
    if _flat_cache.get("key") != key:
        arrays, sizes = engine.flatten_model(model, collision_model)
        desc = engine._ModelDesc()
        for k, v in sizes.items():
            setattr(desc, k, v)
        for k, v in arrays.items():
            setattr(desc, k, v.ctypes.data)
        _flat_cache.update(key=key, arrays=arrays, desc=desc)
    return _flat_cache["desc"]


def ball_certificate(model, collision_model, pair, T, padding=0.0):
    """The item setup's hit certificate (inner balls within the padding, enclosure ruled out along the centre line)
    for placements T [n, 12] of pair `pair`: (balls [n] bool, pokes [n] bool)."""
    lib = load()
    desc = _desc_for(model, collision_model)
    T = np.ascontiguousarray(T, dtype=np.float32).reshape(-1, 12)
    out = np.zeros(len(T), np.int32)
    rc = lib.emu_ball_certificate(ctypes.byref(desc), int(pair), ctypes.c_void_p(T.ctypes.data), len(T), float(padding), ctypes.c_void_p(out.ctypes.data))
    if rc:
        raise RuntimeError(lib.emu_last_error().decode())
    return (out & 1).astype(bool), (out & 2).astype(bool)


def tri_tri(A, B):
    """The triangle stage's closed-form triangle - triangle distance (FP32), A / B: [n, 3, 3]."""
    lib = load()
    A = np.ascontiguousarray(A, dtype=np.float32).reshape(-1, 9)
    B = np.ascontiguousarray(B, dtype=np.float32).reshape(-1, 9)
    out = np.zeros(len(A), dtype=np.float32)
    lib.emu_tri_tri(ctypes.c_void_p(A.ctypes.data), ctypes.c_void_p(B.ctypes.data), len(A), ctypes.c_void_p(out.ctypes.data))
    return out
