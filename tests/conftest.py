import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "host_emu"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


# Known-answer states of the reference's own tests
# (/root/reference/test/core/test_utils.py:129-158 and :165-212).
COLLISION_STATE = np.array([1.98103111, 0.27084068, -1.60819541, -3.07282607, -0.56607161, 2.2466373, 2.80244523, 0.01154539, 0.00453211])
FREE_STATE_1 = np.array([0.5990507, -1.46131209, 0.90076761, -2.70060109, 0.77612686, 1.65669724, -1.72831715, 0.02858002, 0.01910153])
FREE_STATE_2 = np.array([1.61357261, 0.81296483, -2.15473055, -1.38469138, -0.20183876, 3.72713384, 0.34544908, 0.02216068, 0.0398552])


def make_panda(self_collisions=True, objects=True):
    from pyroboplan_b200.models import panda

    m, c, v = panda.load_models()
    if self_collisions:
        panda.add_self_collisions(m, c)
    if objects:
        panda.add_object_collisions(m, c, v)
    return m, c, v


def make_two_dof(objects=True):
    from pyroboplan_b200.models import two_dof

    m, c, v = two_dof.load_models()
    if objects:
        two_dof.add_object_collisions(m, c, v)
    return m, c, v


@pytest.fixture(scope="session")
def panda_full():
    return make_panda()


@pytest.fixture(scope="session")
def panda_self():
    return make_panda(objects=False)


@pytest.fixture(scope="session")
def two_dof_full():
    return make_two_dof()


@pytest.fixture(scope="session")
def panda_oracle(panda_full):
    from oracle.oracle import Oracle

    return Oracle(panda_full[0], panda_full[1])


@pytest.fixture(scope="session")
def two_dof_oracle(two_dof_full):
    from oracle.oracle import Oracle

    return Oracle(two_dof_full[0], two_dof_full[1])


def uniform_states(model, n, seed):
    """float32-representable uniform samples (what both the oracle and the engine are fed)."""
    rng = np.random.default_rng(seed)
    q = rng.uniform(model.lowerPositionLimit, model.upperPositionLimit, size=(n, model.nq))
    return q.astype(np.float32)
