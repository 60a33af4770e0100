#!/usr/bin/env python
"""Generate (and NVRTC-compile for sm_100a, no GPU needed) the model-specialised broadphase of a
bundled model; prints the source size, the cubin size and, with --sass, the instruction count."""
import ctypes
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyroboplan_b200 import engine as E  # noqa: E402


def desc_for(model, cmodel):
    arrays, sizes = E.flatten_model(model, cmodel)
    desc = E._ModelDesc()
    for k, v in sizes.items():
        setattr(desc, k, v)
    for k, v in arrays.items():
        setattr(desc, k, v.ctypes.data)
    return desc, arrays


def jit_source(model, cmodel):
    lib = E.load_library()
    lib.pbp_jit_broadphase_source.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_size_t)]
    desc, keep = desc_for(model, cmodel)
    need = ctypes.c_size_t(0)
    E._check(lib.pbp_jit_broadphase_source(ctypes.byref(desc), None, 0, ctypes.byref(need)))
    buf = ctypes.create_string_buffer(need.value)
    E._check(lib.pbp_jit_broadphase_source(ctypes.byref(desc), buf, need.value, ctypes.byref(need)))
    return buf.value.decode()


def jit_compile(model, cmodel, cc=(10, 0)):
    lib = E.load_library()
    lib.pbp_jit_compile_check.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_size_t), ctypes.c_char_p, ctypes.c_size_t]
    desc, keep = desc_for(model, cmodel)
    nbytes = ctypes.c_size_t(0)
    log = ctypes.create_string_buffer(1 << 16)
    E._check(lib.pbp_jit_compile_check(ctypes.byref(desc), cc[0], cc[1], ctypes.byref(nbytes), log, len(log)))
    return nbytes.value, log.value.decode()


if __name__ == "__main__":
    from pyroboplan_b200.models import panda

    m, c, v = panda.load_models()
    panda.add_self_collisions(m, c)
    panda.add_object_collisions(m, c, v)
    src = jit_source(m, c)
    open("/tmp/pbp_bp_spec.cu", "w").write(src)
    print("source bytes", len(src), "-> /tmp/pbp_bp_spec.cu")
    os.environ["PBP_JIT_DUMP_CUBIN"] = "/tmp/pbp_bp_spec.cubin"
    n, log = jit_compile(m, c)
    print("cubin bytes", n, log[:500])
    if "--sass" in sys.argv:
        sass = subprocess.run(["cuobjdump", "-sass", "/tmp/pbp_bp_spec.cubin"], capture_output=True, text=True).stdout
        lines = [l for l in sass.splitlines() if l.strip().startswith("/*") and ";" in l]
        print("SASS instructions:", len(lines))
        res = subprocess.run(["cuobjdump", "-res-usage", "/tmp/pbp_bp_spec.cubin"], capture_output=True, text=True).stdout
        print(res.strip()[-300:])
