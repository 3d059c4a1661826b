"""Pins the oracle (oracle/soglu_oracle.c, the CPU restatement of the hot path) against the
golden vectors recorded from the unmodified reference, and against live runs of the reference
when oracle/_ref is present."""
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN_CASES, load_golden, ref_harness_path, unpermute, write_case_mtx

TOL_X = 1e-10   # BASELINE.json north_star: relative solution difference


@pytest.mark.parametrize("name", [c for c in GOLDEN_CASES if c != "lap3d_24"])
def test_oracle_matches_reference_x(sg, oracle, tmp_path, name):
    g = load_golden(name)
    p = sg.Problem.from_mtx(write_case_mtx(name, tmp_path))
    x_ext, h = oracle.run(p)
    oracle.free(h)
    x = unpermute(p, x_ext)
    rel = np.linalg.norm(x - g["x"]) / np.linalg.norm(g["x"])
    assert rel <= TOL_X, rel
    assert rel <= 1e-12   # in practice the dense restatement agrees to ~1e-15


def test_oracle_residual(sg, oracle, tmp_path):
    import gen_mtx
    n, r, c, v = gen_mtx.generate("lap3d", 13, 11, 9)
    b = gen_mtx.rhs(n)
    p = sg.Problem.from_coo(n, r, c, v, b)
    x_ext, h = oracle.run(p)
    oracle.free(h)
    x = unpermute(p, x_ext)
    ax = np.zeros(n)
    np.add.at(ax, r, v * x[c])
    assert np.linalg.norm(ax - b) / np.linalg.norm(b) <= 1e-12   # north_star residual gate
    assert np.all(x_ext[n:] == 1.0) or np.allclose(x_ext[n:], 1.0)  # identity padding solves to b = 1


def test_oracle_vs_live_reference(sg, oracle, tmp_path):
    """Random banded unsymmetric matrix: oracle vs the unmodified reference run here."""
    harness = ref_harness_path()
    if harness is None:
        pytest.skip("oracle/_ref not built (or host CPU lacks AVX-512)")
    import gen_mtx
    n, r, c, v = gen_mtx.banded(1500, 120, 7, seed=7)
    path = str(tmp_path / "rb.mtx")
    gen_mtx.write_mtx(path, n, r, c, v)
    out = tmp_path / "ref"
    out.mkdir()
    subprocess.run([harness, path, str(out), "--blocks"], check=True, capture_output=True, env=dict(os.environ, OMP_NUM_THREADS="4"))
    xref = np.fromfile(out / "x.f64")
    p = sg.Problem.from_mtx(path)
    np.testing.assert_array_equal(p.i32("ops"), np.fromfile(out / "ops_fine.i32", dtype=np.int32).reshape(-1, 8))
    np.testing.assert_array_equal(p.i32("perm_new2old"), np.fromfile(out / "perm_new2old.i32", dtype=np.int32))
    x_ext, h = oracle.run(p)
    x = unpermute(p, x_ext)
    assert np.linalg.norm(x - xref) / np.linalg.norm(xref) <= 1e-12
    # factor blocks too
    L = np.fromfile(out / "L.i32", dtype=np.int32).reshape(-1, 3)
    Lv = np.fromfile(out / "L.f64").reshape(-1, 64, 64)
    for k in range(0, len(L), max(1, len(L) // 40)):
        mine = oracle.block(h, L[k, 0])
        assert np.abs(mine - Lv[k]).max() <= 1e-12 * max(1.0, np.abs(Lv[k]).max())
    oracle.free(h)


@pytest.mark.parametrize("n,w,k,seed", [(64, 8, 3, 1), (65, 10, 4, 2), (127, 20, 5, 3), (128, 30, 5, 4), (129, 40, 6, 5),
                                         (200, 199, 9, 6), (511, 60, 7, 7), (513, 25, 3, 8), (1000, 300, 9, 9)])
def test_planner_bit_exact_on_random_matrices(sg, tmp_path, n, w, k, seed):
    """Edge sizes (one block, just over one block, powers of two +-1, nearly dense band): reader + GPS ordering +
    planner versus a LIVE run of the unmodified reference -- permutation, op list, stages and block ids bit-exact."""
    harness = ref_harness_path()
    if harness is None:
        pytest.skip("oracle/_ref not built (or host CPU lacks AVX-512)")
    import gen_mtx
    n, r, c, v = gen_mtx.banded(n, w, k, seed=seed)
    path = str(tmp_path / "m.mtx")
    gen_mtx.write_mtx(path, n, r, c, v)
    out = tmp_path / "ref"
    out.mkdir()
    subprocess.run([harness, path, str(out)], check=True, capture_output=True, env=dict(os.environ, OMP_NUM_THREADS="2"))
    p = sg.Problem.from_mtx(path)
    rd = lambda f, dt: np.fromfile(out / f, dtype=dt)
    np.testing.assert_array_equal(p.i32("perm_new2old"), rd("perm_new2old.i32", np.int32))
    np.testing.assert_array_equal(p.i32("coarse_ops"), rd("ops_coarse.i32", np.int32).reshape(-1, 8))
    np.testing.assert_array_equal(p.i32("ops"), rd("ops_fine.i32", np.int32).reshape(-1, 8))
    np.testing.assert_array_equal(p.i32("laststage"), rd("laststage.i32", np.int32))
    np.testing.assert_array_equal(p.i32("L"), rd("L.i32", np.int32).reshape(-1, 3))
    np.testing.assert_array_equal(p.i32("U"), rd("U.i32", np.int32).reshape(-1, 3))
    np.testing.assert_array_equal(p.f64("b_perm"), rd("b_perm.f64", np.float64))


def test_symmetric_and_disconnected_inputs_vs_live_reference(sg, oracle, tmp_path):
    """A block-diagonal (two disconnected components) symmetric matrix: exercises addMissing (GPSOrder.cpp:251-277)
    and the LL^T path end to end against the reference's x."""
    harness = ref_harness_path()
    if harness is None:
        pytest.skip("oracle/_ref not built (or host CPU lacks AVX-512)")
    import gen_mtx
    n1, r1, c1, v1 = gen_mtx.generate("lap2d", 9, 8)
    n2, r2, c2, v2 = gen_mtx.generate("lap2d", 7, 6)
    n = n1 + n2
    r = np.concatenate([r1, r2 + n1]); c = np.concatenate([c1, c2 + n1]); v = np.concatenate([v1, v2])
    path = str(tmp_path / "two.mtx")
    gen_mtx.write_mtx(path, n, r, c, v, symmetric=True)
    out = tmp_path / "ref"
    out.mkdir()
    subprocess.run([harness, path, str(out)], check=True, capture_output=True, env=dict(os.environ, OMP_NUM_THREADS="2"))
    p = sg.Problem.from_mtx(path)
    np.testing.assert_array_equal(p.i32("perm_new2old"), np.fromfile(out / "perm_new2old.i32", dtype=np.int32))
    np.testing.assert_array_equal(p.i32("ops"), np.fromfile(out / "ops_fine.i32", dtype=np.int32).reshape(-1, 8))
    x_ext, h = oracle.run(p)
    oracle.free(h)
    xref = np.fromfile(out / "x.f64")
    assert np.linalg.norm(unpermute(p, x_ext) - xref) / np.linalg.norm(xref) <= 1e-12


def test_wide_stage_split_vs_live_reference(sg, tmp_path):
    """3D 48^3 is the smallest Laplacian whose widest dependency stage exceeds 8000 ops, so the reference's stage
    split (BlockPlanner.cpp:335-354) and the multi-threaded stage sort are exercised: op list (incl. stage,
    group and sequence numbers), per-block stages and factor structure bit-exact against a LIVE reference run."""
    harness = ref_harness_path()
    if harness is None:
        pytest.skip("oracle/_ref not built (or host CPU lacks AVX-512)")
    import gen_mtx
    n, r, c, v = gen_mtx.generate("lap3d", 48)
    path = str(tmp_path / "m.mtx")
    gen_mtx.write_mtx(path, n, r, c, v)
    out = tmp_path / "ref"
    out.mkdir()
    subprocess.run([harness, path, str(out)], check=True, capture_output=True, env=dict(os.environ, OMP_NUM_THREADS="8"))
    p = sg.Problem.from_mtx(path)
    rd = lambda f, dt: np.fromfile(out / f, dtype=dt)
    ops = p.i32("ops")
    assert np.bincount(ops[:, 5]).max() > 8000
    np.testing.assert_array_equal(ops, rd("ops_fine.i32", np.int32).reshape(-1, 8))
    np.testing.assert_array_equal(p.i32("stage"), rd("stage.i32", np.int32))
    np.testing.assert_array_equal(p.i32("laststage"), rd("laststage.i32", np.int32))
    np.testing.assert_array_equal(p.i32("L"), rd("L.i32", np.int32).reshape(-1, 3))
    np.testing.assert_array_equal(p.i32("U"), rd("U.i32", np.int32).reshape(-1, 3))


def test_duplicate_entries_keep_the_reference_value(sg, tmp_path):
    """A .mtx that lists some (row, col) twice: the reference scatters the sorted cells one by one, so the later
    cell wins (BlockPlanner.cpp:1498-1519).  Our duplicate-free entry list (what goes to the GPU) and its dense
    view must hold the same values as the blocks of a LIVE reference run."""
    harness = ref_harness_path()
    if harness is None:
        pytest.skip("oracle/_ref not built (or host CPU lacks AVX-512)")
    import gen_mtx
    n, r, c, v = gen_mtx.banded(300, 40, 6, seed=11)
    rng = np.random.default_rng(3)
    pick = rng.choice(len(v), 60, replace=False)
    r2, c2, v2 = np.concatenate([r, r[pick]]), np.concatenate([c, c[pick]]), np.concatenate([v, v[pick] * 1.5 + 0.25])
    order = rng.permutation(len(v2))
    path = str(tmp_path / "dup.mtx")
    gen_mtx.write_mtx(path, n, r2[order], c2[order], v2[order])
    out = tmp_path / "ref"
    out.mkdir()
    subprocess.run([harness, path, str(out), "--blocks"], check=True, capture_output=True, env=dict(os.environ, OMP_NUM_THREADS="2"))
    p = sg.Problem.from_mtx(path)
    leaves = np.fromfile(out / "inputs.i32", dtype=np.int32).reshape(-1, 3)
    ref_dense = np.fromfile(out / "inputs.f64").reshape(len(leaves), 4096)
    mine = p.f64("input_vals")
    np.testing.assert_array_equal(p.i32("inputs"), leaves)
    np.testing.assert_array_equal(mine[leaves[:, 0] - 1], ref_dense)
    # the entry list is duplicate-free and reproduces the dense view
    eb, ep, ev = p.i32("entry_block"), p.i32("entry_pos"), p.f64("entry_val")
    keys = eb.astype(np.int64) * 4096 + ep
    assert len(np.unique(keys)) == len(keys) and len(keys) < len(v2) + (p.size("n_ext") - n) + 1
    dense = np.zeros_like(mine)
    dense[eb - 1, ep] = ev
    np.testing.assert_array_equal(dense, mine)
