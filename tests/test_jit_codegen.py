"""The model-specialised broadphase (pyroboplan_b200/csrc/pbp_jit.h): source generation and an offline
NVRTC build for sm_100a for every bundled model -- no GPU needed.  What the kernel computes is
covered by the GPU parity tests (the engine uses it for every boolean query)."""

import os
import sys

import numpy as np
import pytest

from conftest import ROOT, make_panda, make_two_dof

sys.path.insert(0, os.path.join(ROOT, "tools"))
import jit_dump  # noqa: E402


def _marble():
    from pyroboplan_b200.models import marble

    m, c, v = marble.load_models()
    marble.add_object_collisions(m, c, v)
    return m, c


def _sphere_panda():
    from pyroboplan_b200.models import panda

    m, c, v = panda.load_models(use_sphere_collisions=True)
    panda.add_self_collisions(m, c)
    panda.add_object_collisions(m, c, v)
    return m, c


def _generic_axis_arm():
    """Revolute joints about non-axis-aligned axes + a static-vs-static pair: the generator's general paths."""
    from pyroboplan_b200.model import Box, CollisionPair, GeometryModel, GeometryObject, Model, SE3, Sphere, JOINT_REVOLUTE

    model = Model("skew")
    ax = np.array([1.0, 2.0, 2.0]) / 3.0
    j1 = model.addJoint(0, JOINT_REVOLUTE, ax, SE3(np.eye(3), np.array([0.0, 0.0, 0.1])), "a", -2.0, 2.0)
    model.addJoint(j1, JOINT_REVOLUTE, np.array([0.0, 0.6, 0.8]), SE3(np.eye(3), np.array([0.3, 0.0, 0.0])), "b", -2.0, 2.0)
    cm = GeometryModel()
    cm.addGeometryObject(GeometryObject("l1", 1, SE3(np.eye(3), np.array([0.15, 0.0, 0.0])), Box(0.3, 0.05, 0.05)))
    cm.addGeometryObject(GeometryObject("l2", 2, SE3(np.eye(3), np.array([0.15, 0.0, 0.0])), Sphere(0.06)))
    cm.addGeometryObject(GeometryObject("wall", 0, SE3(np.eye(3), np.array([0.5, 0.0, 0.0])), Box(0.1, 1.0, 1.0)))
    cm.addGeometryObject(GeometryObject("ball", 0, SE3(np.eye(3), np.array([0.0, 0.4, 0.2])), Sphere(0.1)))
    for a, b in ((2, 0), (2, 1), (3, 0), (3, 1), (2, 3)):
        cm.addCollisionPair(CollisionPair(a, b))
    return model, cm


@pytest.mark.parametrize("which", ["panda", "two_dof", "marble", "sphere_panda", "generic_axis"])
def test_specialised_broadphase_builds(which):
    make = {"panda": lambda: make_panda()[:2], "two_dof": lambda: make_two_dof()[:2], "marble": _marble,
            "sphere_panda": _sphere_panda, "generic_axis": _generic_axis_arm}[which]
    model, cmodel = make()
    src = jit_dump.jit_source(model, cmodel)
    assert "k_bp_spec" in src and f"#define NP {len(cmodel.collisionPairs)}" in src
    assert src.count("const bool sure") == len(cmodel.collisionPairs)    # one unrolled test per pair
    nbytes, log = jit_dump.jit_compile(model, cmodel)
    assert nbytes > 1000, log


def test_zero_pair_model_is_not_specialised():
    model, cmodel, _ = make_panda(self_collisions=False, objects=False)
    assert jit_dump.jit_source(model, cmodel) == ""
