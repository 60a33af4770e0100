"""The Python boundary mirrors the reference's hot-path signatures -- CPU-side checks."""

import inspect

import numpy as np
import pytest

from conftest import make_two_dof
from pyroboplan_b200.core import utils as core
from pyroboplan_b200.planning import utils as planning

# (function, parameter names) exactly as in the reference:
#   /root/reference/src/pyroboplan/core/utils.py:8-15, 54-56, 90-92, 174, 195-197, 379, 416, 449
#   /root/reference/src/pyroboplan/planning/utils.py:15, 51-60, 114
REFERENCE_SIGNATURES = {
    core.check_collisions_at_state: ["model", "collision_model", "q", "data", "collision_data", "distance_padding"],
    core.get_minimum_distance_at_state: ["model", "collision_model", "q", "data", "collision_data"],
    core.check_collisions_along_path: ["model", "collision_model", "q_path", "data", "collision_data", "distance_padding"],
    core.get_random_state: ["model", "padding"],
    core.get_random_collision_free_state: ["model", "collision_model", "joint_padding", "distance_padding", "max_tries"],
    core.get_collision_geometry_ids: ["model", "collision_model", "body"],
    core.get_collision_pair_indices_from_bodies: ["model", "collision_model", "body_list"],
    core.set_collisions: ["model", "collision_model", "body1", "body2", "enable"],
    core.configuration_distance: ["q_start", "q_end"],
    core.get_path_length: ["q_path"],
    core.check_within_limits: ["model", "q"],
    core.extract_cartesian_pose: ["model", "target_frame", "q", "data"],
    core.extract_cartesian_poses: ["model", "target_frame", "q_vec", "data"],
    planning.extend_robot_state: ["q_parent", "q_sample", "max_connection_distance"],
    planning.has_collision_free_path: ["q1", "q2", "max_step_size", "model", "collision_model", "data", "collision_data", "distance_padding"],
    planning.discretize_joint_space_path: ["q_path", "max_angle_distance"],
}


def _model_module_signatures():
    # /root/reference/src/pyroboplan/models/panda.py:13,57,82; two_dof.py:12,28; ur5.py:9,25; ur10.py:9,25
    from pyroboplan_b200.models import panda, two_dof, ur5, ur10

    return {
        panda.load_models: ["use_sphere_collisions"],
        panda.add_self_collisions: ["model", "collision_model", "srdf_filename"],
        panda.add_object_collisions: ["model", "collision_model", "visual_model", "inflation_radius"],
        two_dof.load_models: [],
        two_dof.add_object_collisions: ["model", "collision_model", "visual_model"],
        ur5.load_ur5_on_base_models: [],
        ur5.add_ur5_on_base_self_collisions: ["model", "collision_model", "srdf_filename"],
        ur10.load_ur10_on_base_models: [],
        ur10.add_ur10_on_base_self_collisions: ["model", "collision_model", "srdf_filename"],
    }


def test_model_modules_mirror_the_reference_signatures():
    for fn, params in _model_module_signatures().items():
        assert list(inspect.signature(fn).parameters) == params, fn.__name__


def test_get_engine_cache_and_invalidate_without_a_device():
    """The engine cache key is cheap and explicit: identities, revision counters, counts; the
    strict mode adds the content fingerprint.  (No device here: only the key logic is exercised.)"""
    import os

    from conftest import make_two_dof
    from pyroboplan_b200 import engine

    model, cmodel, _ = make_two_dof()
    k0 = engine._cheap_key(model, cmodel, None)
    assert k0 == engine._cheap_key(model, cmodel, None)
    cmodel.removeCollisionPair(cmodel.collisionPairs[0])
    assert engine._cheap_key(model, cmodel, None) != k0
    os.environ["PBP_STRICT_CACHE"] = "1"
    try:
        k1 = engine._cheap_key(model, cmodel, None)
        cmodel.geometryObjects[-1].placement.translation[0] += 0.01      # an in-place edit the cheap key cannot see
        assert engine._cheap_key(model, cmodel, None) != k1
    finally:
        del os.environ["PBP_STRICT_CACHE"]
    engine.invalidate_engine(cmodel)      # nothing cached: a no-op, must not raise


@pytest.mark.parametrize("fn", list(REFERENCE_SIGNATURES), ids=lambda f: f.__name__)
def test_signature_matches_reference(fn):
    assert list(inspect.signature(fn).parameters) == REFERENCE_SIGNATURES[fn]


def test_defaults_match_reference():
    sig = inspect.signature(core.get_random_collision_free_state).parameters
    assert (sig["joint_padding"].default, sig["distance_padding"].default, sig["max_tries"].default) == (0.0, 0.0, 100)
    assert inspect.signature(core.check_collisions_at_state).parameters["distance_padding"].default == 0.0


def test_extend_robot_state_vectors():
    # reference test/planning/test_planning_utils.py:12-23
    q_parent = np.array([0.0, 0.0])
    q_sample = np.array([1.0, 0.0])
    np.testing.assert_allclose(planning.extend_robot_state(q_parent, q_sample, 0.25), [0.25, 0.0])
    assert planning.extend_robot_state(q_parent, q_sample, 2.0) is q_sample
    assert planning.extend_robot_state(q_sample, q_sample, 0.1) is q_sample


def test_discretize_vectors():
    # reference test/planning/test_planning_utils.py:26-43
    path = planning.discretize_joint_space_path([np.array([0.0, 0.0]), np.array([2.0, 0.0])], 1.0)
    np.testing.assert_array_equal(np.array(path), [[0.0, 0.0], [1.0, 0.0], [2.0, 0.0]])
    assert len(planning.discretize_joint_space_path([np.array([1.0, 1.0]), np.array([1.0, 1.0])], 0.1)) == 1


def test_zero_pair_model_needs_no_device():
    """2-DOF without obstacles has no active pairs (reference test/planning/test_prm.py:11):
    every query answers "free" on the host without touching the GPU."""
    model, cmodel, _ = make_two_dof(objects=False)
    q = np.array([0.3, -0.2])
    assert core.check_collisions_at_state(model, cmodel, q) == False  # noqa: E712  (np.bool_)
    assert isinstance(core.check_collisions_at_state(model, cmodel, q), np.bool_)
    assert core.check_collisions_along_path(model, cmodel, [q, q + 0.1]) is False
    assert core.get_minimum_distance_at_state(model, cmodel, q) == np.inf
    assert planning.has_collision_free_path(q, q + 1.0, 0.05, model, cmodel) is True
    assert planning.has_collision_free_paths(np.zeros((5, 2)), np.ones((5, 2)), 0.05, model, cmodel).all()
    np.random.seed(1234)
    s = core.get_random_collision_free_state(model, cmodel)
    np.random.seed(1234)
    np.testing.assert_array_equal(s, np.random.uniform(model.lowerPositionLimit, model.upperPositionLimit))


def test_host_helpers():
    assert core.configuration_distance(np.zeros(5), np.array([0.3, 0.0, -0.4, 0.0, 0.0])) == pytest.approx(0.5)
    path = [np.zeros(5), np.array([0.3, 0, -0.4, 0, 0]), np.array([0.3, 1.0, -0.4, 0, 0]), np.array([0.3, 1.0, -0.4, 0, 10.0])]
    assert core.get_path_length(path) == pytest.approx(11.5)  # reference test/core/test_utils.py:36-43
    model, _, _ = make_two_dof(False)
    assert core.check_within_limits(model, np.zeros(2))
    assert not core.check_within_limits(model, model.upperPositionLimit + 0.2)


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure; nothing under pyroboplan_b200/ may reference it."""
    import os
    import re

    from conftest import ROOT

    for dirpath, _, files in os.walk(os.path.join(ROOT, "pyroboplan_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "pbp_oracle" not in text, f
