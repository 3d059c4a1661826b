"""Task compiler (op list -> task graph of the persistent executor) on the host, no GPU:
fusion / aliasing / row-split / segment bookkeeping must stay consistent."""
import ctypes

import numpy as np
import pytest

from conftest import write_case_mtx

NAMES = "tasks pairs succ initial slots levels fused_subs fused_invs aliased_invs split_tasks segments deps maxdeps gemm lu usec".split()


def compile_stats(sg, p, fuse_sub=1, fuse_inv=1, split=1, max_slots=0):
    L = sg.lib()
    L.soglu_debug_compile.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int]
    out = (ctypes.c_int64 * 16)()
    rc = L.soglu_debug_compile(p.h, fuse_sub, fuse_inv, split, max_slots, out, 16)
    if rc:
        raise sg.SogluError(L.soglu_last_error().decode())
    return dict(zip(NAMES, list(out)))


@pytest.fixture(scope="module")
def prob(sg, tmp_path_factory):
    return sg.Problem.from_mtx(write_case_mtx("lap3d_24", tmp_path_factory.mktemp("c")))


def test_plain_compile_counts(sg, prob):
    ops = prob.i32("ops")
    s = compile_stats(sg, prob, 0, 0, 0)
    n_results = len(np.unique(ops[:, 3]))               # one task per result block (lu: two blocks, one task)
    assert s["tasks"] == n_results
    assert s["pairs"] == (np.isin(ops[:, 0], (8, 9, 11))).sum() + (~np.isin(ops[:, 0], (8, 9, 11))).sum()
    assert s["deps"] == s["succ"] and s["segments"] == 1
    assert s["lu"] == (ops[:, 0] == 1).sum()
    assert s["slots"] == 1 + prob.size("n_input") + len(np.unique(np.concatenate([ops[:, 3], ops[ops[:, 0] == 1, 4]])))
    assert s["levels"] == 1422                           # true dependency depth, SURVEY.md Appendix E


def test_fusions_reduce_tasks_and_depth(sg, prob):
    base = compile_stats(sg, prob, 0, 0, 0)
    fs = compile_stats(sg, prob, 1, 0, 0)
    fi = compile_stats(sg, prob, 0, 1, 0)
    both = compile_stats(sg, prob, 1, 1, 0)
    assert fs["fused_subs"] > 0 and fs["tasks"] == base["tasks"] - fs["fused_subs"]
    assert fi["fused_invs"] > 0 and fi["tasks"] == base["tasks"] - fi["fused_invs"] - fi["aliased_invs"]
    assert both["levels"] < fs["levels"] < base["levels"]
    assert both["slots"] < base["slots"]                 # folded products and aliased inverses need no storage


def test_row_split_bookkeeping(sg, prob):
    a = compile_stats(sg, prob, 1, 1, 0)
    b = compile_stats(sg, prob, 1, 1, 1)
    assert b["split_tasks"] > 0 and b["tasks"] > a["tasks"]
    assert b["deps"] == b["succ"] and b["pairs"] == a["pairs"] and b["slots"] == a["slots"]
    assert b["levels"] == a["levels"]


def test_segments_when_pool_is_small(sg, prob):
    full = compile_stats(sg, prob)
    assert full["segments"] == 1
    small = compile_stats(sg, prob, max_slots=5200)
    assert small["segments"] > 1 and small["slots"] <= 5200
    assert small["succ"] < full["succ"]                  # cross-segment edges are dropped
    tiny = compile_stats(sg, prob, max_slots=4600)
    assert tiny["segments"] >= small["segments"]
    with pytest.raises(sg.SogluError):
        compile_stats(sg, prob, max_slots=1500)          # below inputs + factors


def compile_dist(sg, p, pr, pc, nb=1, max_slots=0):
    L = sg.lib()
    L.soglu_debug_compile_dist.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
    n = 16 + 3 * pr * pc
    out = (ctypes.c_int64 * n)()
    rc = L.soglu_debug_compile_dist(p.h, max_slots, pr, pc, nb, out, n)
    if rc:
        raise sg.SogluError(L.soglu_last_error().decode())
    v = list(out)
    d = dict(zip(NAMES, v[:16]))
    d["local_deps"], d["local_succ"], d["remote_operands"] = v[12], v[13], v[14]
    d["per_rank"] = [tuple(v[16 + 3 * r: 19 + 3 * r]) for r in range(pr * pc)]   # (tasks, slots, mirrors)
    return d


@pytest.mark.parametrize("pr,pc,nb", [(1, 2, 1), (2, 2, 1), (2, 2, 4), (2, 4, 2)])
def test_multi_gpu_sharding_bookkeeping(sg, prob, pr, pc, nb):
    """2D block-cyclic owner-computes sharding: every task lands on exactly one GPU, mirrors add one
    fetch task each, and the per-GPU pieces add up to the global graph."""
    one = compile_stats(sg, prob)
    d = compile_dist(sg, prob, pr, pc, nb)
    world = pr * pc
    mirrors = sum(m for _, _, m in d["per_rank"])
    assert mirrors > 0
    assert sum(t for t, _, _ in d["per_rank"]) == d["tasks"]           # partition of the tasks
    assert d["local_deps"] == d["deps"] == d["local_succ"] == d["succ"]  # every edge kept exactly once
    assert d["pairs"] == one["pairs"] + mirrors                           # one operand pair per fetch task
    assert d["segments"] == 1
    slots = sum(s for _, s, _ in d["per_rank"])
    assert slots == one["slots"] + mirrors + (world - 1)                  # + a zero block per extra GPU
    assert max(s for _, s, _ in d["per_rank"]) < one["slots"]             # the share of one GPU is smaller
    # mirrored blocks are read locally: far fewer remote operand reads than remote blocks read directly
    assert d["remote_operands"] <= mirrors + 2 * prob.size("n_input")  # only fetch tasks (and reads of remote inputs) cross GPUs


def test_multi_gpu_with_small_pools_uses_segments(sg, prob):
    d = compile_dist(sg, prob, 1, 2, 1)
    cap = max(s for _, s, _ in d["per_rank"]) - 400
    e = compile_dist(sg, prob, 1, 2, 1, max_slots=cap)
    assert e["segments"] > 1
    assert all(s <= cap for _, s, _ in e["per_rank"])
    assert sum(t for t, _, _ in e["per_rank"]) == e["tasks"]


@pytest.mark.parametrize("split,pr,pc,nb,seed", [(0, 1, 1, 1, 1), (1, 1, 1, 1, 2), (1, 1, 1, 1, 3), (1, 2, 1, 2, 4), (1, 2, 2, 1, 5), (1, 4, 2, 4, 6)])
def test_release_protocol_simulation(sg, prob, split, pr, pc, nb, seed):
    """Host replay of the executor's dependency protocol on the per-GPU arrays it uploads, in a random
    order: row slices of one task share their leader's counter and become ready together; every task
    runs exactly once and never before all writers of its operands are done."""
    L = sg.lib()
    L.soglu_debug_simulate.argtypes = [ctypes.c_void_p] + [ctypes.c_int] * 4 + [ctypes.c_uint64, ctypes.c_void_p]
    out = (ctypes.c_int64 * 3)()
    rc = L.soglu_debug_simulate(prob.h, split, pr, pc, nb, seed, out)
    assert rc == 0, L.soglu_last_error().decode()
    done, bad, total = list(out)
    assert done == total and bad == 0
