"""Host-only checks of the scheduling work around the executor: the two-queue task classes of option hi_shared and
the timed executor model (csrc/device/model.cpp, tools/model.py).  No GPU."""
import ctypes
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, write_case_mtx

sys.path.insert(0, os.path.join(ROOT, "tools"))


@pytest.fixture(scope="module")
def prob(sg, tmp_path_factory):
    return sg.Problem.from_mtx(write_case_mtx("lap3d_24", tmp_path_factory.mktemp("m")))


@pytest.mark.parametrize("slack,max_slots", [(1, 0), (50, 0), (1000, 0), (10 ** 7, 0), (200, 7000)])
def test_priority_classes_are_consistent(sg, prob, slack, max_slots):
    L = sg.lib()
    L.soglu_debug_check_queues.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p]
    out = (ctypes.c_int64 * 4)()
    assert L.soglu_debug_check_queues(prob.h, slack, max_slots, out) == 0, L.soglu_last_error().decode()
    tasks, nhi, nseg, bad = list(out)
    assert bad == 0
    assert 0 < nhi <= tasks
    if slack >= 10 ** 7:
        assert nhi == tasks          # everything is within the slack: one (high-priority) queue
    if max_slots:
        assert nseg > 1


def test_threshold_orders_the_classes(sg, prob):
    L = sg.lib()
    L.soglu_debug_check_queues.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p]
    n = []
    for slack in (1, 30, 300, 3000):
        out = (ctypes.c_int64 * 4)()
        assert L.soglu_debug_check_queues(prob.h, slack, 0, out) == 0
        n.append(out[1])
    assert n == sorted(n) and n[0] < n[-1]


def test_model_runs_every_task_and_respects_bounds(sg, prob):
    import model as m
    base = m.model(prob)
    busy_bound = base["busy_ms_per_cta"]
    assert base["makespan_ms"] >= base["critical_ms"] * 0.999      # never faster than the dependent chain
    assert base["makespan_ms"] >= busy_bound                       # nor than the work divided by the CTAs
    ideal = m.model(prob, policy=1)
    two = m.model(prob, policy=2, hi_slack_us=100)
    assert ideal["makespan_ms"] <= base["makespan_ms"] * 1.02
    assert two["hi"] > 0 and two["makespan_ms"] >= ideal["critical_ms"] * 0.999
    # a faster diagonal kernel shortens the chain; one CTA serialises everything
    assert m.model(prob, t_lu_fused=5.0)["critical_ms"] < base["critical_ms"]
    one = m.model(prob, n_ctas=1)
    assert one["makespan_ms"] >= one["busy_ms_per_cta"] >= base["busy_ms_per_cta"] * 147


def test_model_uses_the_compilers_classes(sg, prob):
    import model as m
    a = m.model(prob, policy=2, hi_slack_us=300)
    b = m.model(prob, policy=2, compile_hi_slack=300)
    assert a["hi"] == b["hi"] and abs(a["makespan_ms"] - b["makespan_ms"]) < 1e-9
