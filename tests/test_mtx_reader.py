"""Reader quirks of the reference (mtx.cpp:44-182) that feed the bit-exact planner."""
import numpy as np


def _write(path, text):
    with open(path, "w") as f:
        f.write(text)


def test_reader_rules(sg, tmp_path):
    n = 64
    lines = ["%%MatrixMarket matrix coordinate real general", "% comment", "%d %d %d" % (n, n, 3 * n)]
    for i in range(1, n + 1):
        lines.append("%d %d 4.0" % (i, i))
        lines.append("%d %d 0" % (i, (i % n) + 1))            # explicit zero: dropped (mtx.cpp:102)
        lines.append("%d %d -1 %% trailing comment" % (i, (i % n) + 1))  # contains '%': skipped (mtx.cpp:93)
    lines.append("%d 1 7.0" % (n + 5))                          # out of range: dropped (mtx.cpp:104)
    lines.append("1")                                           # <= 3 chars: skipped
    p = tmp_path / "q.mtx"
    _write(p, "\n".join(lines) + "\n")
    _write(tmp_path / "q_b.mtx", "%%MatrixMarket matrix array real general\n%d 1\n2.5\n3.5\n" % n)
    prob = sg.Problem.from_mtx(str(p))
    assert prob.size("dim") == n and prob.size("nnz") == n and prob.size("symmetric") == 0
    b = prob.f64("b")
    assert b[0] == 2.5 and b[1] == 3.5 and np.all(b[2:] == 1.0)   # missing rhs entries become 1.0 (mtx.cpp:177-179)


def test_symmetric_banner_and_mirroring(sg, tmp_path):
    n = 70
    lines = ["%%MatrixMarket matrix coordinate real symmetric", "%d %d %d" % (n, n, 2 * n - 1)]
    for i in range(1, n + 1):
        lines.append("%d %d 4" % (i, i))
        if i > 1:
            lines.append("%d %d -1" % (i, i - 1))
    p = tmp_path / "s.mtx"
    _write(p, "\n".join(lines) + "\n")
    prob = sg.Problem.from_mtx(str(p))
    assert prob.size("symmetric") == 1
    assert prob.size("nnz") == 2 * n - 1 and prob.size("nnz_expanded") == 3 * n - 2   # solver.cpp:136-149
    assert prob.size("n_U") == 0 and 10 in set(prob.i32("ops")[:, 0].tolist())       # llt path


def test_config_sizes(sg, tmp_path):
    """config.cpp:41-49: blockRows = 2^(floor(log2(dim/64))+1); exact powers of two double."""
    import gen_mtx
    for nx, ny, rows, rows2 in ((8, 8, 2, 2), (16, 8, 4, 4), (32, 32, 32, 8), (50, 37, 32, 8)):
        n, r, c, v = gen_mtx.generate("lap2d", nx, ny)
        prob = sg.Problem.from_coo(n, r, c, v, gen_mtx.rhs(n))
        assert prob.size("block_rows") == rows and prob.size("block_rows_l2") == rows2
        assert prob.size("n_ext") == rows * 64


def test_too_small_matrix_is_rejected(sg):
    i = np.arange(10)
    try:
        sg.Problem.from_coo(10, i, i, np.ones(10), np.ones(10))
    except sg.SogluError as e:
        assert "not supported" in str(e)
    else:
        raise AssertionError("dim < 64 must be rejected")


def test_default_process_grids(sg):
    assert [sg.default_grid(n) for n in (1, 2, 4, 8)] == [(1, 1), (2, 1), (2, 2), (4, 2)]
