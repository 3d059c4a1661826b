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


@pytest.mark.parametrize("split,pr,pc,nb,seed", [(0, 1, 1, 1, 1), (1, 1, 1, 1, 2), (1, 1, 1, 1, 3), (1, 2, 1, 2, 4), (1, 2, 2, 1, 5), (1, 4, 2, 4, 6),
                                                 (1, 1, 1, 1, 0), (1, 2, 1, 2, 0), (1, 2, 2, 1, 0), (1, 4, 2, 4, 0)])
def test_release_protocol_simulation(sg, prob, split, pr, pc, nb, seed):
    """Host replay of the executor's dependency protocol on the per-GPU arrays it uploads: row slices of one task share
    their leader's counter and become ready together; every task runs exactly once and never before all writers of its
    operands are done.  seed > 0: tasks run in a random dependency-driven order.  seed 0: the executor's own discipline --
    three workers per GPU claim their GPU's tasks in task order and wait on the claimed task's counter; the run must not
    get stuck (the static order of every GPU is a filter of one global topological order)."""
    L = sg.lib()
    L.soglu_debug_simulate.argtypes = [ctypes.c_void_p] + [ctypes.c_int] * 4 + [ctypes.c_uint64, ctypes.c_void_p]
    out = (ctypes.c_int64 * 3)()
    rc = L.soglu_debug_simulate(prob.h, split, pr, pc, nb, seed, out)
    assert rc == 0, L.soglu_last_error().decode()
    done, bad, total = list(out)
    assert done == total and bad == 0


# ---- raw op lists: every validation error of the task compiler (SURVEY.md Appendix E invariants) ----------
LU, LINV, UINV, SUB, MUL, MULNEG, LLT, MULT = 1, 2, 3, 4, 8, 9, 10, 11


def compile_raw(sg, n_ids, inputs, ops, keep=(), max_slots=0):
    """ops: list of (op, src, src2, result, result2)."""
    L = sg.lib()
    L.soglu_debug_compile_raw.argtypes = [ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64] + [ctypes.c_void_p] * 5 + \
                                         [ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p]
    a = np.array(ops, dtype=np.int64).reshape(-1, 5)
    col = lambda k, dt=np.int32: np.ascontiguousarray(a[:, k], dtype=dt)
    op, src, src2, res, res2 = col(0, np.uint8), col(1), col(2), col(3), col(4)
    ins = np.array(inputs, dtype=np.int32)
    kp = np.array(keep, dtype=np.int32)
    out = (ctypes.c_int64 * 4)()
    ptr = lambda x: x.ctypes.data_as(ctypes.c_void_p)
    rc = L.soglu_debug_compile_raw(n_ids, len(ins), ptr(ins), len(a), ptr(src), ptr(src2), ptr(op), ptr(res), ptr(res2), len(kp), ptr(kp), max_slots, out)
    if rc:
        raise sg.SogluError(L.soglu_last_error().decode())
    return dict(zip(("tasks", "pairs", "slots", "segments"), out))


GOOD = [(LU, 1, 0, 3, 4), (LINV, 3, 0, 5, 0), (UINV, 4, 0, 6, 0), (MUL, 5, 2, 7, 0), (MUL, 2, 6, 8, 0), (MUL, 8, 7, 9, 0),
        (SUB, 9, 1, 10, 0), (LU, 10, 0, 11, 12)]     # inputs 1, 2; one 2x2 elimination step


def test_raw_good_list_compiles(sg):
    s = compile_raw(sg, 13, [1, 2], GOOD, keep=[3, 4, 7, 8, 11, 12])
    # lu + both inverses fused into one task; the product 9 is folded into the sub: 2 lu + 3 GEMM tasks, and in
    # these narrow levels every GEMM task is split into 4 row slices
    assert s["tasks"] == 2 + 3 * 4 and s["pairs"] == 5 and s["segments"] == 1


@pytest.mark.parametrize("ops,n_ids,msg", [
    ([(7, 1, 0, 3, 0)], 13, "unsupported op code 7"),
    ([(MUL, 1, 99, 3, 0)], 13, "block id out of range"),
    ([(MUL, 1, 2, 0, 0)], 13, "block id out of range"),
    ([(LU, 1, 0, 3, 0)], 13, "lu without second result"),
    ([(MUL, 3, 2, 3, 0)], 13, "reads its own result"),
    ([(LU, 1, 0, 3, 1)], 13, "reads its own result"),
    ([(MUL, 1, 1, 2, 0)], 13, "writes input block 2"),
    ([(LU, 1, 0, 3, 2)], 13, "writes input block 2"),
    ([(MUL, 1, 2, 3, 0), (MULNEG, 1, 2, 3, 0)], 13, "block 3 has writers of mixed or non-accumulating kinds"),
    ([(SUB, 1, 2, 3, 0), (SUB, 2, 1, 3, 0)], 13, "block 3 has writers of mixed or non-accumulating kinds"),
    ([(LU, 1, 0, 3, 4), (LU, 2, 0, 5, 4)], 13, "block 4 has writers of mixed or non-accumulating kinds"),
    ([(MUL, 3, 2, 4, 0), (MUL, 1, 2, 3, 0)], 13, "not in dependency order"),
])
def test_raw_validation_errors(sg, ops, n_ids, msg):
    with pytest.raises(sg.SogluError) as e:
        compile_raw(sg, n_ids, [1, 2], ops)
    assert msg in str(e.value)


def test_raw_error_reports_the_lowest_offending_op(sg):
    ops = [(MUL, 1, 2, 3, 0)] * 5 + [(MUL, 4, 2, 4, 0)] + [(MUL, 1, 2, 3, 0)] * 50000 + [(MUL, 5, 2, 5, 0)]
    with pytest.raises(sg.SogluError) as e:
        compile_raw(sg, 13, [1, 2], ops)
    assert "op 5 reads its own result" in str(e.value)


def test_raw_accumulation_chain_keeps_op_order(sg):
    # 3000 products into one block: the operand pairs of the chain must stay in op-list order (rounding), whatever
    # thread claimed them
    ops = [(MUL, 1, 2, 3, 0)] * 3000
    s = compile_raw(sg, 4, [1, 2], ops, keep=[3])
    assert s["tasks"] == 4 and s["pairs"] == 3000      # one chain, four row slices sharing the 3000 pairs


def test_raw_pool_limits(sg):
    with pytest.raises(sg.SogluError) as e:
        compile_raw(sg, 13, [1, 2], GOOD, keep=[3, 4, 7, 8, 11, 12], max_slots=9)
    assert "pool too small" in str(e.value)


def test_symmetric_path_fuses_cholesky_and_inverse(sg, tmp_path_factory):
    """LL^T path (BlockPlanner.cpp:941-989): the lowerInv of an llt result is folded into the llt task and
    repeated inverses alias it, as on the LU path; the release protocol still runs every task once."""
    p = sg.Problem.from_mtx(write_case_mtx("lap2d_64_sym", tmp_path_factory.mktemp("s")))
    ops = p.i32("ops")
    n_llt, n_inv = int((ops[:, 0] == 10).sum()), int((ops[:, 0] == 2).sum())
    plain = compile_stats(sg, p, 1, 0, 0)
    fused = compile_stats(sg, p, 1, 1, 0)
    assert n_llt > 0 and (ops[:, 0] == 1).sum() == 0
    assert fused["fused_invs"] + fused["aliased_invs"] == n_inv and fused["fused_invs"] <= n_llt
    assert fused["tasks"] == plain["tasks"] - n_inv and fused["levels"] < plain["levels"]
    L = sg.lib()
    L.soglu_debug_simulate.argtypes = [ctypes.c_void_p] + [ctypes.c_int] * 4 + [ctypes.c_uint64, ctypes.c_void_p]
    for cfg in ((1, 1, 1, 1), (1, 2, 2, 1)):
        out = (ctypes.c_int64 * 3)()
        assert L.soglu_debug_simulate(p.h, *cfg, 7, out) == 0
        assert out[0] == out[2] and out[1] == 0
