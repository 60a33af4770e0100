#!/usr/bin/env python
"""Lower the reference's bundled robot descriptions to neutral JSON descriptions.

Run in the authoring container (needs /root/reference):

    python tools/make_model_assets.py

Reads URDF + SRDF + collision STLs under
/root/reference/src/pyroboplan/models/{panda_description,two_dof_description} and writes
``pyroboplan_b200/models/assets/{panda,panda_spheres,two_dof,marble}.json``: link/joint tables, the
unique vertices of every collision mesh (all ten Panda meshes are exactly convex: volume
ratio 1.000, every vertex on the hull -- recorded per mesh), and the SRDF disabled pairs.
The JSON is what travels to the GPU box; /root/reference does not exist there.
"""

import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from pyroboplan_b200.model.urdf import srdf_disabled_pairs, urdf_to_description  # noqa: E402

REF_MODELS = "/root/reference/src/pyroboplan/models"
OUT = os.path.join(os.path.dirname(HERE), "pyroboplan_b200", "models", "assets")


def emit(name, urdf_rel, srdf_rel=None):
    desc = urdf_to_description(os.path.join(REF_MODELS, urdf_rel), package_dirs=[REF_MODELS])
    desc["source"] = f"pyroboplan models/{urdf_rel}"
    if srdf_rel:
        desc["disabled_collisions"] = [list(p) for p in srdf_disabled_pairs(os.path.join(REF_MODELS, srdf_rel))]
    for rec in desc["meshes"].values():
        rec["vertices"] = [[float(f"{v:.9g}") for v in p] for p in rec["vertices"]]
    path = os.path.join(OUT, f"{name}.json")
    with open(path, "w") as f:
        json.dump(desc, f, separators=(",", ":"))
    nverts = sum(len(m["vertices"]) for m in desc["meshes"].values())
    print(f"{path}: {len(desc['links'])} links, {len(desc['joints'])} joints, "
          f"{len(desc['meshes'])} meshes ({nverts} vertices), {os.path.getsize(path)} bytes")
    for fn, m in desc["meshes"].items():
        print(f"   {fn}: {len(m['vertices'])} verts, {m['n_triangles']} tris, "
              f"vol ratio {m['volume_ratio']:.6f}, all on hull {m['all_vertices_on_hull']}")


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    emit("panda", "panda_description/urdf/panda.urdf", "panda_description/srdf/panda.srdf")
    emit("panda_spheres", "panda_description/urdf/panda_spheres.urdf", "panda_description/srdf/panda.srdf")
    emit("two_dof", "two_dof_description/two_dof.urdf")
    emit("marble", "marble_description/marble.urdf")
    emit("ur5_on_base", "ur_description/urdf/ur5_on_base.urdf", "ur_description/srdf/ur5_on_base.srdf")
    emit("ur10_on_base", "ur_description/urdf/ur10_on_base.urdf", "ur_description/srdf/ur10_on_base.srdf")
