"""GPU parity: the CUDA hot path through the C ABI versus (a) golden x recorded from the
unmodified reference, (b) the oracle port on the same inputs, block by block for the factors,
and (c) size-independent properties at larger sizes.

Tolerances (BASELINE.json north_star): relative solution difference <= 1e-10; residual
||Ax-b||/||b|| <= 1e-12 on the well-conditioned 3D cases."""
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN_CASES, load_golden, ref_harness_path, unpermute, write_case_mtx

pytestmark = pytest.mark.gpu
TOL_X = 1e-10


def _rel(a, b):
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_solution_matches_reference_golden(sg, tmp_path, name):
    g = load_golden(name)
    p = sg.Problem.from_mtx(write_case_mtx(name, tmp_path))
    ctx = sg.Context(0)
    ctx.load(p)
    fs = ctx.factor()
    x, ss = ctx.solve(p)
    assert fs["kernel_launches"] >= 1 and ss["kernel_launches"] >= 1
    assert not np.isnan(x).any()
    assert _rel(x, g["x"]) <= TOL_X
    ctx.close()


@pytest.mark.parametrize("name", ["lap2d_64", "lap3d_13x11x9", "banded_3000", "lap3d_16_sym", "lap2d_64_sym"])
def test_factor_blocks_match_oracle(sg, oracle, tmp_path, name):
    p = sg.Problem.from_mtx(write_case_mtx(name, tmp_path))
    ctx = sg.Context(0)
    ctx.load(p)
    ctx.factor()
    x_ext_o, h = oracle.run(p)
    worst = 0.0
    for what in ("L", "U"):
        ids = p.i32(what)
        for k in range(len(ids)):
            mine = ctx.get_block(ids[k, 0])
            ref = oracle.block(h, ids[k, 0])
            worst = max(worst, np.abs(mine - ref).max() / max(1.0, np.abs(ref).max()))
    oracle.free(h)
    assert worst <= 1e-12, worst
    x_ext, _ = ctx.solve_ext(p.f64("b_perm"))
    assert _rel(x_ext, x_ext_o) <= TOL_X
    ctx.close()


@pytest.mark.parametrize("fuse_sub,fuse_inv", [(0, 0), (1, 0), (0, 1), (1, 1)])
def test_executor_modes_agree(sg, tmp_path, fuse_sub, fuse_inv):
    """Persistent DAG executor vs one-launch-per-level debug executor, with and without the
    task fusions: same kernels, so the results agree to rounding of the fused epilogues."""
    g = load_golden("lap3d_24")
    p = sg.Problem.from_mtx(write_case_mtx("lap3d_24", tmp_path))
    xs = []
    for mode in (0, 1):
        ctx = sg.Context(0)
        ctx.set_option("exec_mode", mode)
        ctx.set_option("fuse_sub", fuse_sub)
        ctx.set_option("fuse_inv", fuse_inv)
        ctx.load(p)
        fs = ctx.factor()
        if mode == 1:
            assert fs["kernel_launches"] > 100      # one launch per dependency level
        x, _ = ctx.solve(p)
        xs.append(x)
        ctx.close()
    assert _rel(xs[0], xs[1]) <= 1e-13
    assert _rel(xs[0], g["x"]) <= TOL_X


def test_refactor_and_multiple_rhs(sg, tmp_path):
    """Factor twice (same pattern), solve several right-hand sides: idempotence + linearity."""
    import gen_mtx
    n, r, c, v = gen_mtx.generate("lap3d", 14)
    b1 = gen_mtx.rhs(n)
    p = sg.Problem.from_coo(n, r, c, v, b1)
    ctx = sg.Context(0)
    ctx.load(p)
    ctx.factor()
    x1, _ = ctx.solve(p)
    ctx.factor()
    x1b, _ = ctx.solve(p)
    np.testing.assert_array_equal(x1, x1b)          # deterministic re-factorisation
    rng = np.random.default_rng(3)
    b2 = rng.standard_normal(n)
    x2, _ = ctx.solve(p, b2)
    x12, _ = ctx.solve(p, b1 + 2.0 * b2)
    assert _rel(x12, x1 + 2.0 * x2) <= 1e-12          # linearity of the solve
    ax = np.zeros(n)
    np.add.at(ax, r, v * x2[c])
    assert np.linalg.norm(ax - b2) / np.linalg.norm(b2) <= 1e-12
    ctx.close()


def test_array_abi_without_problem_helper(sg, tmp_path):
    """Drive soglu_set_blocks / set_graph / set_factors directly with plain arrays (what a
    reference-side binding would pass), not through soglu_load_problem."""
    g = load_golden("lap2d_50x37")
    p = sg.Problem.from_mtx(write_case_mtx("lap2d_50x37", tmp_path))
    ops = p.i32("ops")
    ctx = sg.Context(0)
    n_in = p.size("n_input")
    ctx.set_blocks(p.size("storage"), np.arange(1, n_in + 1, dtype=np.int32), p.f64("input_vals"))
    ctx.set_graph({"op": ops[:, 0], "src": ops[:, 1], "src2": ops[:, 2], "result": ops[:, 3], "result2": ops[:, 4]}, stage=ops[:, 5])
    ctx.set_factors(p.i32("L"), p.i32("U"), p.size("block_rows"))
    ctx.factor()
    x_ext, _ = ctx.solve_ext(p.f64("b_perm"))
    assert _rel(unpermute(p, x_ext), g["x"]) <= TOL_X
    ctx.close()


def test_error_paths(sg):
    ctx = sg.Context(0)
    with pytest.raises(sg.SogluError):
        ctx.factor()                                   # nothing loaded
    # an op list that writes an input block violates the invariants -> SOGLU_ERR_GRAPH
    ctx.set_blocks(4, np.array([1], dtype=np.int32), np.eye(64).reshape(1, -1))
    ctx.set_graph({"op": [4], "src": [0], "src2": [2], "result": [1], "result2": [0]})
    ctx.set_factors(np.array([[1, 0, 0]], dtype=np.int32), np.array([[1, 0, 0]], dtype=np.int32), 1)
    with pytest.raises(sg.SogluError) as e:
        ctx.factor()
    assert "writes input block" in str(e.value)
    ctx.close()


def test_solve_cli(sg, tmp_path):
    """./solve <file.mtx>: reference console lines + a correct <base>_x.mtx."""
    g = load_golden("lap3d_13x11x9")
    path = write_case_mtx("lap3d_13x11x9", tmp_path)
    out = subprocess.run([sg.SOLVE_PATH, path], capture_output=True, text=True, check=True).stdout
    for token in ("GGPS reorder: levels:", "re Order time:", "reduced ops to:", "plan time:", "kernel time:", "solve triangled", "max rhs error:"):
        assert token in out, out
    err = float(out.split("max rhs error:")[1].split()[0])
    assert err <= 1e-11
    xs = [float(t) for t in open(path.replace(".mtx", "_x.mtx")).read().split("\n")[2:] if t.strip()]
    assert _rel(np.array(xs), g["x"]) <= TOL_X


@pytest.mark.parametrize("name", ["lap3d_24", "nine2d_40", "banded_3000", "lap2d_64_sym"])
def test_reference_side_adapter(sg, tmp_path, name):
    """The boundary seen from the reference: oracle/_ref/ref_adapter is the UNMODIFIED reference (reader, GPS ordering,
    both planner passes, iniBlockStorage, result un-permutation) linked with oracle/soglu_adapter.cpp -- the binding of
    INTEGRATION.md section 2 -- so that BlockPlanner::calculate / ::solve (solver.cpp:106, 115) run in libsoglu_b200.so
    through the array ABI (dense 64x64 input blocks, op list from data::graph, factor leaves from the quadtrees).
    Its x must be the reference's own x (golden vector)."""
    from conftest import ROOT
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_adapter")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/ref_adapter not built (needs the reference sources at build time)")
    g = load_golden(name)
    path = write_case_mtx(name, tmp_path)
    out = str(tmp_path / "x.f64")
    r = subprocess.run([exe, path, out], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "ADAPTER factor_s" in r.stdout and "ADAPTER solve_s" in r.stdout
    x = np.fromfile(out)
    assert _rel(x, g["x"]) <= TOL_X


def test_solve_lu_dropin(sg):
    import gen_mtx
    n, r, c, v = gen_mtx.generate("lap2d", 40, 33)
    b = gen_mtx.rhs(n)
    x = sg.solve_lu(n, r, c, v, b)
    ax = np.zeros(n)
    np.add.at(ax, r, v * x[c])
    assert np.linalg.norm(ax - b) / np.linalg.norm(b) <= 1e-11


def test_midsize_against_live_reference(sg, tmp_path):
    """3D 40^3 (n=64 000, 340 k ops): GPU x vs the unmodified reference run on the host."""
    harness = ref_harness_path()
    if harness is None:
        pytest.skip("oracle/_ref not present on this box")
    path = str(tmp_path / "l3d40.mtx")
    sg.write_stencil_mtx("lap3d", path, 40)
    out = tmp_path / "ref"
    out.mkdir()
    subprocess.run([harness, path, str(out)], check=True, capture_output=True, env=dict(os.environ, OMP_NUM_THREADS="16"))
    xref = np.fromfile(out / "x.f64")
    p = sg.Problem.from_mtx(path)
    np.testing.assert_array_equal(p.i32("ops"), np.fromfile(out / "ops_fine.i32", dtype=np.int32).reshape(-1, 8))
    ctx = sg.Context(0)
    ctx.load(p)
    ctx.factor()
    x, _ = ctx.solve(p)
    assert _rel(x, xref) <= TOL_X
    ctx.close()


@pytest.mark.timeout(600)
def test_config2_against_live_reference(sg, tmp_path):
    """BASELINE config 2 (3D 7-pt 64^3) at full size against the UNMODIFIED reference run here on the host (about 13 s on
    16 threads; the largest BASELINE stencil config whose reference run fits this box -- 100^3 is OOM-killed at 205 GB,
    profiles/r02_reference_100_oom.md): relative solution difference <= 1e-10 (north-star gate), for the raw solve and
    for the refined one the benchmark times."""
    harness = ref_harness_path()
    if harness is None:
        pytest.skip("oracle/_ref not present on this box")
    path = str(tmp_path / "l3d64.mtx")
    sg.write_stencil_mtx("lap3d", path, 64)
    out = tmp_path / "ref"
    out.mkdir()
    subprocess.run([harness, path, str(out), "--lean"], check=True, capture_output=True, env=dict(os.environ, OMP_NUM_THREADS="16"))
    xref = np.fromfile(out / "x.f64")
    p = sg.Problem.from_mtx(path)
    ctx = sg.Context(0)
    ctx.load(p)
    ctx.factor()
    x, _ = ctx.solve(p)
    xr, _ = ctx.solve(p, refine=1)
    assert _rel(x, xref) <= TOL_X
    assert _rel(xr, xref) <= TOL_X
    ctx.close()


def test_config2_properties(sg, tmp_path):
    """BASELINE config 2 (3D 7-pt 64^3, n = 262 144, 7.1 M ops) at full size: residual gate,
    padding rows, and agreement with a second factorisation."""
    path = str(tmp_path / "l3d64.mtx")
    sg.write_stencil_mtx("lap3d", path, 64)
    p = sg.Problem.from_mtx(path)
    assert p.size("n_ops") == 7106089 and p.size("storage") == 1529367
    ctx = sg.Context(0)
    ctx.load(p)
    fs = ctx.factor()
    x, ss = ctx.solve(p)
    n = p.size("dim")
    # 7-point Laplacian residual without materialising A on the host twice
    X = x.reshape(64, 64, 64)
    ax = 6.0 * X
    ax[1:, :, :] -= X[:-1, :, :]; ax[:-1, :, :] -= X[1:, :, :]
    ax[:, 1:, :] -= X[:, :-1, :]; ax[:, :-1, :] -= X[:, 1:, :]
    ax[:, :, 1:] -= X[:, :, :-1]; ax[:, :, :-1] -= X[:, :, 1:]
    b = 1.0 + 0.25 * (np.arange(n) % 7)
    res = np.linalg.norm(ax.ravel() - b) / np.linalg.norm(b)
    assert res <= 1e-12, res
    assert abs(x[0] - 1.0377) < 1e-4 and abs(x[1] - 1.74207) < 1e-4 and abs(x[2] - 2.25873) < 1e-4
    assert fs["flops"] > 3.5e12
    ctx.close()


@pytest.mark.parametrize("name,max_slots", [("lap3d_24", 5200), ("lap3d_24", 7000), ("lap2d_64", 800), ("banded_3000", 1670)])
def test_pool_recycling_segments(sg, tmp_path, name, max_slots):
    """A block pool smaller than the number of blocks: the factorisation runs as several executor
    launches with slots recycled at the boundaries; results are unchanged."""
    g = load_golden(name)
    p = sg.Problem.from_mtx(write_case_mtx(name, tmp_path))
    ref = sg.Context(0)
    ref.load(p)
    f0 = ref.factor()
    x0, _ = ref.solve(p)
    assert f0["pool_blocks"] > max_slots           # the cap really binds
    ctx = sg.Context(0)
    ctx.set_option("max_slots", max_slots)
    ctx.load(p)
    fs = ctx.factor()
    assert fs["pool_blocks"] <= max_slots
    assert fs["kernel_launches"] >= 2              # more than one segment (+ input packing)
    x, _ = ctx.solve(p)
    np.testing.assert_array_equal(x, x0)           # same kernels, same order of accumulation
    assert _rel(x, g["x"]) <= TOL_X
    fs2 = ctx.factor()                             # re-factorisation reuses the recycled pool correctly
    x2, _ = ctx.solve(p)
    np.testing.assert_array_equal(x2, x0)
    # L/U survive, temporaries do not
    ctx.get_block(p.i32("L")[0, 0])
    ops = p.i32("ops")
    with pytest.raises(sg.SogluError):
        for o in ops[ops[:, 0] == 4][:200]:        # some early `sub` results must have been recycled
            ctx.get_block(o[3])
    ctx.close()
    ref.close()


def test_pool_too_small_is_an_error(sg, tmp_path):
    p = sg.Problem.from_mtx(write_case_mtx("lap3d_24", tmp_path))
    ctx = sg.Context(0)
    ctx.set_option("max_slots", 2000)              # below inputs + factors
    ctx.load(p)
    with pytest.raises(sg.SogluError) as e:
        ctx.factor()
    assert "pool too small" in str(e.value)
    ctx.close()


def test_two_gpu_sharded_factorisation():
    """One factorisation sharded over 2 GPUs (one process per GPU, CUDA IPC peer pools, NVLink pulls):
    bitwise equal to the single-GPU result, with and without slot recycling (segments)."""
    import json
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from conftest import ROOT
    for extra, port in (([], 29541), (["max_slots=4000"], 29542), (["dist_nb=1", "mirror_min=2"], 29543)):
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
               "--master-port", str(port), os.path.join(ROOT, "tests", "_dist_worker.py"), "lap3d", "24"] + extra
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
        assert out.returncode == 0, out.stderr[-3000:]
        rec = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
        assert rec["world"] == 2 and rec["bitwise_equal"], rec
        assert all(r["tasks"] > 0 for r in rec["per_rank"])
        assert sum(r["mirrored"] for r in rec["per_rank"]) > 0


@pytest.mark.parametrize("name,opts", [("lap3d_24", {}), ("lap3d_24", {"max_slots": 4000}), ("lap2d_64_sym", {}), ("banded_3000", {"dist_nb": 2})])
def test_in_process_two_gpus(sg, tmp_path, name, opts):
    """soglu_create(n_gpus = 2): one process, one compilation, two GPUs wired with cudaDeviceEnablePeerAccess -- the
    entry point SOGLU::solveLU / ./solve use with SOGLU_GPUS.  Bitwise the single-GPU solution (every block is produced
    by the same task with the same operand order), with and without pool recycling (segments), and refinement on top."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    p = sg.Problem.from_mtx(write_case_mtx(name, tmp_path))
    one = sg.Context(0)
    one.load(p)
    one.factor()
    x1, _ = one.solve(p)
    two = sg.Context(n_gpus=2)
    for k, v in opts.items():
        two.set_option(k, v)
    two.load(p)
    for _ in range(2):
        fs = two.factor()
        x2, _ = two.solve(p)
        np.testing.assert_array_equal(x2, x1)
    assert fs["tasks"] > 0 and fs["seconds"] > 0
    if opts.get("max_slots"):
        assert two.segments() > 1
    xr1, _ = one.solve(p, refine=1)
    xr2, _ = two.solve(p, refine=1)
    np.testing.assert_array_equal(xr2, xr1)
    lid = int(p.i32("L")[5, 0])
    np.testing.assert_array_equal(two.get_block(lid), one.get_block(lid))
    two.close()
    one.close()


def test_solve_cli_two_gpus(sg, tmp_path):
    """SOGLU_GPUS=2 ./solve file.mtx (main.cpp:34-66 with the sharded context behind SOGLU::solveLU): same _x.mtx as on one GPU."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    g = load_golden("lap3d_24")
    xs = []
    for gpus in ("1", "2"):
        d = tmp_path / ("g" + gpus)
        d.mkdir()
        path = write_case_mtx("lap3d_24", d)
        r = subprocess.run([sg.SOLVE_PATH, path], capture_output=True, text=True, timeout=300, env=dict(os.environ, SOGLU_GPUS=gpus))
        assert r.returncode == 0, r.stdout + r.stderr
        if gpus == "2":
            assert "sharded over 2 GPUs" in r.stdout
        vals = [float(l) for l in open(path[: path.find(".mtx")] + "_x.mtx").read().split("\n")[2:] if l.strip()]
        xs.append(np.array(vals))
    np.testing.assert_array_equal(xs[0], xs[1])
    assert _rel(xs[1], g["x"]) <= TOL_X


def test_iterative_refinement_lowers_residual(sg, tmp_path):
    """Device-side refinement (r = b - A x in FP64, re-solve, update) drives the residual to the FP64 floor
    (SURVEY.md 0.9): on the 2D Laplacian, where the raw solve sits near 1e-12, one step must not be worse
    and must stay within the solution-parity tolerance of the reference."""
    g = load_golden("lap2d_64")
    p = sg.Problem.from_mtx(write_case_mtx("lap2d_64", tmp_path))
    import gen_mtx
    n, r, c, v = gen_mtx.generate("lap2d", 64)
    b = gen_mtx.rhs(n)

    def resid(x):
        ax = np.zeros(n)
        np.add.at(ax, r, v * x[c])
        return np.linalg.norm(ax - b) / np.linalg.norm(b)
    ctx = sg.Context(0)
    ctx.load(p)
    ctx.factor()
    x0, s0 = ctx.solve(p)
    x1, s1 = ctx.solve(p, refine=1)
    x2, _ = ctx.solve(p, refine=2)
    assert s1["kernel_launches"] > s0["kernel_launches"]
    assert resid(x1) <= resid(x0) * 1.05 and resid(x2) <= resid(x0) * 1.05
    assert resid(x1) <= 1e-12
    assert _rel(x1, g["x"]) <= TOL_X and _rel(x2, g["x"]) <= TOL_X
    # a perturbed factorisation-free check: refinement must fix a deliberately bad start? (not exposed) -- instead
    # check a second right-hand side goes through the same path
    rng = np.random.default_rng(5)
    b2 = rng.standard_normal(n)
    y, _ = ctx.solve(p, b2, refine=1)
    ay = np.zeros(n)
    np.add.at(ay, r, v * y[c])
    assert np.linalg.norm(ay - b2) / np.linalg.norm(b2) <= 1e-12
    ctx.close()


@pytest.mark.parametrize("n,w,k,seed", [(64, 8, 3, 1), (65, 10, 4, 2), (129, 40, 6, 5), (200, 199, 9, 6), (513, 25, 3, 8)])
def test_edge_sizes_match_oracle(sg, oracle, n, w, k, seed):
    """One block row, just over one, nearly dense band: CUDA path vs the oracle on the same planned problem."""
    import gen_mtx
    n, r, c, v = gen_mtx.banded(n, w, k, seed=seed)
    p = sg.Problem.from_coo(n, r, c, v, gen_mtx.rhs(n))
    ctx = sg.Context(0)
    ctx.load(p)
    ctx.factor()
    x_ext, _ = ctx.solve_ext(p.f64("b_perm"))
    xo, h = oracle.run(p)
    oracle.free(h)
    assert _rel(x_ext, xo) <= TOL_X
    ctx.close()


def test_symmetric_disconnected_matrix(sg, oracle):
    """LL^T path (llt, mult, transposed back-substitution) on a matrix with two connected components."""
    import gen_mtx
    n1, r1, c1, v1 = gen_mtx.generate("lap2d", 9, 8)
    n2, r2, c2, v2 = gen_mtx.generate("lap2d", 7, 6)
    n = n1 + n2
    r = np.concatenate([r1, r2 + n1]); c = np.concatenate([c1, c2 + n1]); v = np.concatenate([v1, v2])
    keep = r >= c
    p = sg.Problem.from_coo(n, r[keep], c[keep], v[keep], gen_mtx.rhs(n), symmetric=True)
    assert p.size("n_U") == 0
    ctx = sg.Context(0)
    ctx.load(p)
    ctx.factor()
    x, _ = ctx.solve(p)
    xo, h = oracle.run(p)
    oracle.free(h)
    assert _rel(x, unpermute(p, xo)) <= TOL_X
    ax = np.zeros(n)
    np.add.at(ax, r, v * x[c])
    b = gen_mtx.rhs(n)
    assert np.linalg.norm(ax - b) / np.linalg.norm(b) <= 1e-12
    ctx.close()


def test_sparse_input_upload_equals_dense_upload(sg, tmp_path):
    """soglu_set_blocks_sparse (entry list scattered on the device) must leave exactly the blocks that
    soglu_set_blocks (dense 64x64 arrays) leaves: same factors bit for bit, also when the values change on
    the same pattern (refactorisation) and when the entries arrive in another order."""
    p = sg.Problem.from_mtx(write_case_mtx("lap3d_13x11x9", tmp_path))
    ops = p.i32("ops")
    n_in = p.size("n_input")
    ids = np.arange(1, n_in + 1, dtype=np.int32)
    graph = {"op": ops[:, 0], "src": ops[:, 1], "src2": ops[:, 2], "result": ops[:, 3], "result2": ops[:, 4]}
    ent_in, ent_pos, ent_val = p.i32("entry_block") - 1, p.i32("entry_pos"), p.f64("entry_val")

    def run(upload):
        ctx = sg.Context(0)
        upload(ctx, 1.0)
        ctx.set_graph(graph)
        ctx.set_factors(p.i32("L"), p.i32("U"), p.size("block_rows"))
        ctx.factor()
        first = [ctx.get_block(int(i)) for i in p.i32("U")[:40, 0]]
        upload(ctx, 3.0)                       # new values, same pattern
        ctx.factor()
        second = [ctx.get_block(int(i)) for i in p.i32("U")[:40, 0]]
        x_ext, _ = ctx.solve_ext(p.f64("b_perm"))
        ctx.close()
        return first, second, x_ext

    dense = p.f64("input_vals")
    rng = np.random.default_rng(5)
    perm = rng.permutation(len(ent_val))
    d1, d2, xd = run(lambda ctx, f: ctx.set_blocks(p.size("storage"), ids, np.ascontiguousarray(dense * f)))
    s1, s2, xs = run(lambda ctx, f: ctx.set_blocks_sparse(p.size("storage"), ids, np.ascontiguousarray(ent_in[perm]), np.ascontiguousarray(ent_pos[perm]),
                                                          np.ascontiguousarray(ent_val[perm] * f)))
    for a, b in zip(d1 + d2, s1 + s2):
        assert np.array_equal(a, b)
    assert np.array_equal(xd, xs)
    assert not np.array_equal(d1[0], d2[0])    # the second factorisation really saw the new values


def test_sparse_input_upload_rejects_bad_entries(sg):
    ctx = sg.Context(0)
    ids = np.array([1], dtype=np.int32)
    ok = (np.array([0], dtype=np.int32), np.array([65], dtype=np.int32), np.array([2.0]))
    ctx.set_blocks_sparse(4, ids, *ok)
    for bad_in, bad_pos in ((1, 0), (-1, 0), (0, 4096), (0, -1)):
        with pytest.raises(sg.SogluError) as e:
            ctx.set_blocks_sparse(4, ids, np.array([bad_in], dtype=np.int32), np.array([bad_pos], dtype=np.int32), np.array([1.0]))
        assert "entry outside" in str(e.value)
    ctx.close()


# ---- compiler / kernel variants ------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.timeout(300)
@pytest.mark.parametrize("name", ["lap2d_64", "lap3d_24", "nine2d_40", "banded_3000", "lap3d_13x11x9", "lap2d_64_sym", "lap3d_16_sym"])
def test_diagonal_kernel_factors(sg, oracle, tmp_path, name):
    """The blocked diagonal-block kernel (lu_blocked.cuh: 16-column panels, one warp on the pivot chain, DMMA trailing
    updates; emulated on the host by tests/test_lub_emulation.py): L, U blocks against the oracle, and a second
    factorisation in the same context (the stage buffers are its work space, the pivot barriers flip phase per task)."""
    g = load_golden(name)
    p = sg.Problem.from_mtx(write_case_mtx(name, tmp_path))
    ctx = sg.Context(0)
    ctx.load(p)
    ctx.factor()
    assert ctx.diag_warnings() == 0
    x, _ = ctx.solve(p)
    assert _rel(x, g["x"]) <= TOL_X
    x_ext, h = oracle.run(p)
    worst = 0.0
    for bid in list(p.i32("L")[:40, 0]) + list(p.i32("U")[:40, 0]):      # (U is empty on the symmetric path)
        ref = oracle.block(h, bid)
        worst = max(worst, float(np.abs(ctx.get_block(bid) - ref).max() / max(1.0, np.abs(ref).max())))
    oracle.free(h)
    assert worst <= 1e-10
    ctx.factor()
    x2, _ = ctx.solve(p)
    np.testing.assert_array_equal(x2, x)
    ctx.close()


@pytest.mark.gpu
@pytest.mark.timeout(300)
@pytest.mark.parametrize("name", ["lap3d_24", "banded_3000"])
def test_slack_split(sg, tmp_path, name):
    """split_slack (default 100 us): row slices for near-critical GEMM tasks in wide levels too; every slice computes the
    same dot products, so the solution is bitwise the one without it."""
    p = sg.Problem.from_mtx(write_case_mtx(name, tmp_path))
    ref = sg.Context(0)
    ref.set_option("split_slack", 0)
    ref.load(p)
    f0 = ref.factor()
    x0, _ = ref.solve(p)
    ctx = sg.Context(0)
    ctx.set_option("split_slack", 1000)
    ctx.load(p)
    fs = ctx.factor()
    assert fs["tasks"] >= f0["tasks"]
    x, _ = ctx.solve(p)
    np.testing.assert_array_equal(x, x0)
    ctx.close()
    ref.close()


@pytest.mark.gpu
@pytest.mark.timeout(300)
def test_watchdog_aborts_and_recovers(sg, tmp_path):
    """The executor's watchdog (executor.cu).  The test hook debug_drop_task loses the completion signal of one task, so
    its successors are never published and the claim-then-wait queue would spin forever; with a 300 ms deadline the
    scheduler lanes raise the abort word, every CTA drains and soglu_factor returns SOGLU_ERR_CUDA naming the queue slot
    instead of hanging in cudaStreamSynchronize.  The context stays usable: without the fault the next factorisation
    gives the golden solution."""
    g = load_golden("lap3d_24")
    p = sg.Problem.from_mtx(write_case_mtx("lap3d_24", tmp_path))
    ctx = sg.Context(0)
    ctx.load(p)
    ctx.factor()
    ctx.set_option("watchdog_ms", 300)
    ctx.set_option("debug_drop_task", 1000)
    with pytest.raises(sg.SogluError) as e:
        ctx.factor()
    assert "watchdog" in str(e.value) and "task (position" in str(e.value)
    ctx.set_option("debug_drop_task", -1)
    ctx.set_option("watchdog_ms", 60000)
    ctx.factor()
    x, _ = ctx.solve(p)
    assert _rel(x, g["x"]) <= TOL_X
    ctx.close()


@pytest.mark.gpu
def test_diag_warnings_flag_nan_pivots(sg):
    """soglu_diag_warnings = the reference's inv_check_diag signal (MatrixStdDouble.cpp:2871, " upper out of tolerance"):
    0 for a healthy matrix, > 0 when a NaN reaches the pivots."""
    import gen_mtx
    n, r, c, v = gen_mtx.generate("lap2d", 24)
    # the NaN goes into the FIRST diagonal block of the reordered matrix: like the reference's check (after upperInv only),
    # ours looks at the blocks whose inverse the factorisation forms -- the last diagonal block has none
    i0 = int(sg.Problem.from_coo(n, r, c, v, gen_mtx.rhs(n)).i32("perm_new2old")[0])
    v = v.copy()
    v[(r == i0) & (c == i0)] = np.nan
    p = sg.Problem.from_coo(n, r, c, v, gen_mtx.rhs(n))
    ctx = sg.Context(0)
    ctx.load(p)
    ctx.factor()
    assert ctx.diag_warnings() > 0
    ctx.close()
