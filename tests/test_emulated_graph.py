"""The compiled task graph, executed on the host (tests/emu/emu_tasks.cpp, test infrastructure), must reproduce the
factors the reference's operation list defines -- checked against the oracle, no GPU.  Covers what the task compiler
does to the list: fused sub / inverse tasks, aliased inverses, row slices, slot recycling, and
pool recycling across segments."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, write_case_mtx

TOL = 1e-10


@pytest.fixture(scope="module")
def emu(sg, tmp_path_factory):
    out = str(tmp_path_factory.mktemp("emu") / "libemu_tasks.so")
    pkg = os.path.join(ROOT, "sparse-operator-graph-lu_b200")
    subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-fopenmp", "-shared", "-o", out, os.path.join(ROOT, "tests", "emu", "emu_tasks.cpp"),
                    "-L" + pkg, "-lsoglu_b200", "-Wl,-rpath," + pkg], check=True)
    L = ctypes.CDLL(out)
    vp, i64 = ctypes.c_void_p, ctypes.c_int64
    L.emu_run.argtypes = [i64, i64, vp, vp, i64, vp, vp, vp, vp, vp, i64, vp, ctypes.c_int, ctypes.c_double, i64, ctypes.c_int, ctypes.c_uint64, vp, vp, ctypes.c_char_p, ctypes.c_int, vp, vp, vp]
    return L


def run_emu(emu, p, mode=0, slack=1e9, max_slots=0, split=1, seed=12345, grid=None):
    ptr = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    ops = p.i32("ops")
    cols = [np.ascontiguousarray(ops[:, k]) for k in range(5)]
    opc = np.ascontiguousarray(ops[:, 0].astype(np.uint8))
    n_in = p.size("n_input")
    ids = np.arange(1, n_in + 1, dtype=np.int32)
    vals = p.f64("input_vals")
    keep = np.ascontiguousarray(np.concatenate([p.i32("L")[:, 0], p.i32("U")[:, 0]]).astype(np.int32))
    out = np.zeros((len(keep), 64, 64))
    stats = np.zeros(6, dtype=np.int64)
    g = np.array(grid if grid else (1, 1, 1), dtype=np.int32)
    brow, bcol = np.ascontiguousarray(p.i32("block_row")), np.ascontiguousarray(p.i32("block_col"))
    err = ctypes.create_string_buffer(256)
    rc = emu.emu_run(p.size("storage"), n_in, ptr(ids), ptr(vals), len(ops), ptr(cols[1]), ptr(cols[2]), ptr(opc), ptr(cols[3]), ptr(cols[4]),
                     len(keep), ptr(keep), mode, slack, max_slots, split, seed, ptr(out), ptr(stats), err, 256,
                     ptr(g) if grid else None, ptr(brow), ptr(bcol))
    assert rc == 0, err.value.decode()
    return keep, out, dict(zip("tasks segments slots cuts split_tasks mirrors".split(), stats.tolist()))


def check_against_oracle(oracle, p, keep, blocks):
    x_ref, h = oracle.run(p)
    num = den = 0.0
    for k, bid in enumerate(keep):
        ref = oracle.block(h, bid)
        num += float(np.sum((blocks[k] - ref) ** 2))
        den += float(np.sum(ref ** 2))
    oracle.free(h)
    assert np.sqrt(num / den) <= TOL
    # and the solution from the emulated factors (block substitution by the oracle)
    ptr = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    h2 = oracle.L.oracle_create(p.size("storage"))
    oracle.L.oracle_set_inputs(h2, len(keep), ptr(keep), ptr(np.ascontiguousarray(blocks)))
    Lf, Uf = np.ascontiguousarray(p.i32("L")), np.ascontiguousarray(p.i32("U"))
    b = p.f64("b_perm")
    x = np.zeros_like(b)
    rc = oracle.L.oracle_solve(h2, len(Lf), ptr(Lf), len(Uf), ptr(Uf) if len(Uf) else None, p.size("block_rows"), p.size("symmetric"), ptr(b), ptr(x))
    oracle.free(h2)
    assert rc == 0
    assert np.linalg.norm(x - x_ref) / np.linalg.norm(x_ref) <= TOL


CASES = ["lap2d_64", "lap2d_64_sym", "nine2d_40", "lap3d_13x11x9", "lap3d_24", "lap3d_16_sym", "banded_3000"]


@pytest.mark.parametrize("name", CASES)
def test_compiled_graph_reproduces_the_factors(sg, emu, oracle, tmp_path, name):
    p = sg.Problem.from_mtx(write_case_mtx(name, tmp_path))
    keep, blocks, st = run_emu(emu, p)
    assert st["cuts"] == 0 and st["segments"] == 1
    check_against_oracle(oracle, p, keep, blocks)


@pytest.mark.parametrize("name,max_slots", [("lap3d_24", 5200), ("lap2d_64", 800)])
def test_slot_recycling_keeps_the_factors(sg, emu, oracle, tmp_path, name, max_slots):
    p = sg.Problem.from_mtx(write_case_mtx(name, tmp_path))
    keep, blocks, st = run_emu(emu, p, max_slots=max_slots)
    assert st["segments"] > 1 and st["slots"] <= max_slots
    check_against_oracle(oracle, p, keep, blocks)


@pytest.mark.parametrize("name", ["lap3d_24", "banded_3000"])
def test_slack_based_row_split_is_bitwise_neutral(sg, emu, oracle, tmp_path, name):
    """Option split_slack: near-critical GEMM tasks are cut into row slices in wide levels too.  A slice computes its rows
    exactly as the whole task would, so the factors do not change by a bit."""
    p = sg.Problem.from_mtx(write_case_mtx(name, tmp_path))
    keep, b0, st0 = run_emu(emu, p, split=5)      # 8 SMs: most levels of the small case count as wide
    _, b1, st1 = run_emu(emu, p, split=7)
    assert st1["tasks"] > st0["tasks"] and st1["split_tasks"] > st0["split_tasks"]
    np.testing.assert_array_equal(b1, b0)
    check_against_oracle(oracle, p, keep, b1)


def test_execution_order_does_not_matter(sg, emu, tmp_path):
    """Task order and two random dependency-driven orders give bitwise the same factors (each task's arithmetic is fixed)."""
    p = sg.Problem.from_mtx(write_case_mtx("lap3d_24", tmp_path))
    _, b0, _ = run_emu(emu, p, seed=0)
    for seed in (1, 2):
        _, b, _ = run_emu(emu, p, seed=seed)
        np.testing.assert_array_equal(b, b0)


@pytest.mark.parametrize("name,grid,max_slots", [("lap3d_24", None, 0), ("banded_3000", None, 0), ("nine2d_40", None, 0), ("lap3d_24", (2, 2, 2), 0),
                                                 ("lap3d_24", None, 5200), ("lap3d_24", (2, 2, 2), 2500)])
def test_static_order_is_a_valid_schedule(sg, emu, oracle, tmp_path, name, grid, max_slots):
    """The executor claims tasks in task order and waits on the claimed task's counter, so the compiled order (tasks sorted
    by latest start time, Compiler::static_order) must be topological: executing the tasks one after the other in that order
    gives the factors, bit for bit those of the operation-list order."""
    p = sg.Problem.from_mtx(write_case_mtx(name, tmp_path))
    keep, b_static, st = run_emu(emu, p, seed=0, grid=grid, max_slots=max_slots)
    _, b_oplist, st2 = run_emu(emu, p, seed=0, split=1 | 8, grid=grid, max_slots=max_slots)
    assert st["tasks"] == st2["tasks"]
    np.testing.assert_array_equal(b_static, b_oplist)
    check_against_oracle(oracle, p, keep, b_static)
    # the two ends of the order key (option order_alpha): earliest start only, latest start only
    for alpha in (0, 100):
        _, b, _ = run_emu(emu, p, mode=alpha + 1, seed=0, grid=grid, max_slots=max_slots)
        np.testing.assert_array_equal(b, b_static)


@pytest.mark.parametrize("grid,max_slots", [((2, 1, 2), 0), ((2, 2, 1), 0), ((4, 2, 4), 0), ((2, 2, 2), 0), ((2, 1, 4), 4000), ((2, 2, 2), 2500)])
def test_sharded_graph_reproduces_the_factors(sg, emu, oracle, tmp_path, grid, max_slots):
    """The graph compiled for several GPUs (owner-computes, remote blocks mirrored by fetch tasks, per-owner pools), with
    and without pool recycling, run in a dependency-driven random order: same factors as the oracle."""
    p = sg.Problem.from_mtx(write_case_mtx("lap3d_24", tmp_path))
    _, _, st0 = run_emu(emu, p)
    keep, blocks, st = run_emu(emu, p, max_slots=max_slots, grid=grid)
    assert st["mirrors"] > 0 and st["tasks"] > st0["tasks"]
    if max_slots:
        assert st["segments"] > 1
    check_against_oracle(oracle, p, keep, blocks)


# ---- random operation lists (not from the planner): the compiled graph must still mean what the list says ------------------
LU, LINV, UINV, SUB, MUL, MULNEG, LLT, MULT = 1, 2, 3, 4, 8, 9, 10, 11


def random_op_list(seed, n_steps=60):
    """A random valid list in the sense of SURVEY.md App. E: topological, one kind of writers per block, chains complete
    before they are read.  Blocks are kept small in norm (products of [-1,1]/64 entries) and lu / llt only see diagonally
    dominant / SPD blocks, so the comparison is not dominated by conditioning."""
    rng = np.random.default_rng(seed)
    blocks, kinds = [None], ["none"]                 # id 0 = none
    def new_input(kind):
        a = rng.uniform(-1, 1, (64, 64)) / 64
        if kind == "dom":
            a += np.eye(64) * rng.uniform(1.5, 3.0)
        if kind == "spd":
            a = a @ a.T + np.eye(64) * rng.uniform(1.0, 2.0)
        blocks.append(a); kinds.append(kind)
        return len(blocks) - 1
    inputs = [new_input(k) for k in ["dom", "dom", "gen", "gen", "gen", "spd", "gen", "dom"]]
    n_input = len(inputs)
    ops = []
    def new_id(kind):
        blocks.append(None); kinds.append(kind)
        return len(blocks) - 1
    def pick(*ks):
        c = [i for i in range(1, len(kinds)) if kinds[i] in ks]
        return int(rng.choice(c)) if c else 0
    for _ in range(n_steps):
        what = rng.choice(["chain", "chain", "chain", "sub", "sub", "lu", "inv", "llt", "negcopy"])
        if what == "chain":
            code = int(rng.choice([MUL, MUL, MULNEG, MULT]))
            pairs = [(pick("gen", "prod", "sub", "linv", "uinv", "L", "U"), pick("gen", "prod", "sub", "L", "U", "linv")) for _k in range(int(rng.integers(1, 6)))]
            r = new_id("prod")
            for a_, b_ in pairs:
                ops.append((code, a_, b_, r, 0))
        elif what == "sub":
            s1 = pick("prod", "gen")
            s2 = pick("dom", "gen", "sub", "domsub")
            if s1 == s2:
                continue
            r = new_id("domsub" if kinds[s2] in ("dom", "domsub") else "sub")
            ops.append((SUB, s1, s2, r, 0))
        elif what == "negcopy":
            s = pick("prod", "gen", "sub")
            r = new_id("sub")
            ops.append((SUB, s, 0, r, 0) if rng.random() < 0.5 else (SUB, 0, s, r, 0))
        elif what == "lu":
            s = pick("dom", "domsub")
            l, u = new_id("L"), new_id("U")
            ops.append((LU, s, 0, l, u))
        elif what == "llt":
            s = pick("spd")
            ops.append((LLT, s, 0, new_id("L"), 0))
        elif what == "inv":
            s = pick("L", "U")
            if s:
                ops.append((LINV if kinds[s] == "L" else UINV, s, 0, new_id("linv" if kinds[s] == "L" else "uinv"), 0))
    produced = sorted({o[3] for o in ops} | {o[4] for o in ops if o[4] > 0})
    keep = [i for i in produced if kinds[i] in ("L", "U") or rng.random() < 0.3]
    dense = np.ascontiguousarray(np.stack([blocks[i] for i in inputs]))
    return len(blocks), np.array(inputs, dtype=np.int32), dense, np.array(ops, dtype=np.int64), np.array(keep, dtype=np.int32), n_input


@pytest.mark.parametrize("seed", range(12))
@pytest.mark.parametrize("variant", ["default", "small_pool"])
def test_random_operation_lists(emu, oracle, seed, variant):
    n_ids, inputs, dense, ops, keep, n_input = random_op_list(seed)
    ptr = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    col = lambda k, dt=np.int32: np.ascontiguousarray(ops[:, k], dtype=dt)
    opc, src, src2, res, res2 = col(0, np.uint8), col(1), col(2), col(3), col(4)
    # the list as written (oracle: one op after the other)
    h = oracle.L.oracle_create(n_ids)
    oracle.L.oracle_set_inputs(h, len(inputs), ptr(inputs), ptr(dense))
    rc = oracle.L.oracle_factor(h, len(ops), ptr(src), ptr(src2), ptr(opc), ptr(res), ptr(res2))
    assert rc == 0
    ref = np.stack([oracle.block(h, int(i)) for i in keep])
    oracle.free(h)
    # the compiled graph
    out = np.zeros((len(keep), 64, 64))
    stats = np.zeros(6, dtype=np.int64)
    err = ctypes.create_string_buffer(256)
    mode, slack, max_slots = 0, 0.0, 0
    if variant == "small_pool":
        max_slots = n_input + len(keep) + 14
    rc = emu.emu_run(n_ids, len(inputs), ptr(inputs), ptr(dense), len(ops), ptr(src), ptr(src2), ptr(opc), ptr(res), ptr(res2), len(keep), ptr(keep),
                     mode, slack, max_slots, 1, 1000 + seed, ptr(out), ptr(stats), err, 256, None, None, None)
    if rc != 0 and variant == "small_pool" and b"pool too small" in err.value:
        pytest.skip("live set larger than the tiny pool")
    assert rc == 0, err.value.decode()
    assert np.isfinite(ref).all() and np.isfinite(out).all()
    assert np.linalg.norm(out - ref) / np.linalg.norm(ref) <= TOL
