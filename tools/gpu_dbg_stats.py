"""Developer counters of one 1 Mi-state boolean round (debug build of the library, see tools/build_dbg.sh):

    PBP_ENGINE_LIB=pyroboplan_b200/libpbp_engine_dbg.so python tools/gpu_dbg_stats.py
"""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyroboplan_b200 import engine
from pyroboplan_b200.engine import CollisionEngine
from pyroboplan_b200.models import panda

NAMES = {0: "sr lane items", 1: "sr ST_NONE", 2: "sr ST_HIT", 3: "sr ST_WALK", 4: "sr ST_TRIANGLES", 5: "sr ST_VALUE", 6: "tri-stage calls",
         7: "tri pairs sum", 8: "tri pairs max", 9: "tri cycles sum", 10: "tri cycles max", 11: "sr lane-level cycles (warp sum)", 12: "walk cycles sum",
         13: "sr kernel cycles max warp", 14: "sr kernel cycles warp sum", 15: "tri ca_n sum", 16: "tri cb_n sum", 17: "walk POKES", 18: "walk ENCLOSED",
         19: "walk UNSURE", 20: "tri have_n", 21: "tri hits", 22: "sr bin chunks", 23: "sr list items", 24: "sr parked lane items",
         30: "np items", 31: "np trips (warp)", 32: "np support lanes sum", 33: "np cycles warp sum", 34: "np cycles max warp", 35: "np parked",
         36: "np doubts", 37: "np hits", 38: "np free", 40: "tri: cycles to end of A cull (sum)", 41: "tri: cycles to end of B cull (sum)", 42: "tri: cycles to first hit (sum)", 43: "sr max cycles of one list chunk", 44: "sr max cycles of one bin chunk"}
model, cmodel, vmodel = panda.load_models()
panda.add_self_collisions(model, cmodel); panda.add_object_collisions(model, cmodel, vmodel)
n = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 1 << 20
q = np.random.default_rng(int(os.environ.get("PBP_DBG_SEED", 0))).uniform(model.lowerPositionLimit, model.upperPositionLimit, size=(n, model.nq)).astype(np.float32)
qd = torch.as_tensor(q, device="cuda:0")
eng = CollisionEngine(model, cmodel, device=0)
lib = engine.load_library()
buf = (ctypes.c_ulonglong * 48)()
hit = torch.empty(n, dtype=torch.bool, device="cuda:0")

for _ in range(2):
    eng.check_states(qd)
if hasattr(lib, "pbp_debug_stats"):
    lib.pbp_debug_stats(buf, 1)
    eng.check_states(qd)
    lib.pbp_debug_stats(buf, 1)
    for i in range(48):
        if buf[i] or i in NAMES:
            print(f"{i:3d} {NAMES.get(i, ''):36s} {buf[i]}")
eng.set_profiling(True); eng.profile(reset=True)
for _ in range(5):
    eng.check_states(qd)
torch.cuda.synchronize()
pr = eng.profile(reset=True)
print({k: (round(v / 5 * 1e3, 1) if k.endswith("_ms") else v) for k, v in pr.items()}, "(us per round)")
