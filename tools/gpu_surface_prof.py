"""Three 1 Mi-state boolean checks per mesh semantics ("hull", then "surface"), for an ncu launch list:

    ncu --metrics gpu__time_duration.sum --clock-control none --csv python tools/gpu_surface_prof.py
"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyroboplan_b200.engine import CollisionEngine
from pyroboplan_b200.models import panda
model, cmodel, vmodel = panda.load_models()
panda.add_self_collisions(model, cmodel); panda.add_object_collisions(model, cmodel, vmodel)
q = np.random.default_rng(0).uniform(model.lowerPositionLimit, model.upperPositionLimit, size=(1 << 20, model.nq)).astype(np.float32)
qd = torch.as_tensor(q, device="cuda:0")
for sem in ("hull", "surface"):
    eng = CollisionEngine(model, cmodel, device=0, mesh_semantics=sem)
    for _ in range(3):
        eng.check_states(qd)
    torch.cuda.synchronize()
    eng.close()
