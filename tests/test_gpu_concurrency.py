"""One engine shared by host threads and CUDA streams (include/pbp_engine.h, conventions): calls serialise inside
the engine, so concurrent callers get the verdicts a single caller gets."""

import threading

import numpy as np
import pytest

from conftest import make_panda, uniform_states

pytestmark = pytest.mark.gpu


def test_threads_and_streams_share_one_engine():
    import torch

    from pyroboplan_b200.engine import CollisionEngine

    model, cmodel, _ = make_panda()
    eng = CollisionEngine(model, cmodel, device=0)
    batches = [torch.as_tensor(uniform_states(model, 60000 + 1000 * k, 40 + k), device="cuda:0") for k in range(4)]
    ref = [eng.check_states(b).cpu().numpy() for b in batches]
    refd = [eng.check_states(b, return_distance=True)[1].cpu().numpy() for b in batches]
    out, errors = {}, []

    def work(k):
        try:
            st = torch.cuda.Stream(device="cuda:0")
            with torch.cuda.stream(st):
                for rep in range(6):
                    h = eng.check_states(batches[k])
                    _, d = eng.check_states(batches[k], return_distance=True)
                    small = eng.check_states(batches[k][:64])
                st.synchronize()
            out[k] = (h.cpu().numpy(), d.cpu().numpy(), small.cpu().numpy())
        except Exception as exc:   # noqa: BLE001
            errors.append(repr(exc))

    threads = [threading.Thread(target=work, args=(k,)) for k in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    for k in range(4):
        assert (out[k][0] == ref[k]).all() and np.array_equal(out[k][1], refd[k]) and (out[k][2] == ref[k][:64]).all()


def test_engine_on_another_current_device_keeps_the_callers_device():
    import torch

    from pyroboplan_b200.engine import CollisionEngine

    model, cmodel, _ = make_panda()
    eng = CollisionEngine(model, cmodel, device="cuda:0")      # torch.device / "cuda:N" strings are accepted
    assert eng.device == 0 and torch.cuda.current_device() == 0
    q = uniform_states(model, 1000, 3)
    assert eng.check_states(q).dtype == np.bool_
    assert torch.cuda.current_device() == 0
