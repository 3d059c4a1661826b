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


def _reference_reader(text):
    """The reference's reader rules (mtx.cpp:44-117), line by line, as the serial restatement did them."""
    import ctypes
    libc = ctypes.CDLL(None)
    libc.strtod.restype = ctypes.c_double
    libc.strtod.argtypes = [ctypes.c_char_p, ctypes.c_void_p]
    libc.strtol.restype = ctypes.c_long
    libc.strtol.argtypes = [ctypes.c_char_p, ctypes.POINTER(ctypes.c_char_p), ctypes.c_int]

    libc.strtol.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p), ctypes.c_int]
    libc.strtod.argtypes = [ctypes.c_void_p, ctypes.c_void_p]

    def fields(line, third_is_double):     # strtol, strtol, strtol / strtod on the NUL-terminated line
        buf = ctypes.create_string_buffer(line.split(b"\0")[0])
        end = ctypes.c_void_p()
        a = libc.strtol(ctypes.addressof(buf), ctypes.byref(end), 10)
        c = libc.strtol(end.value, ctypes.byref(end), 10)
        third = libc.strtod(end.value, None) if third_is_double else libc.strtol(end.value, None, 10)
        return a, c, third
    lines = text.split(b"\n")
    if text.endswith(b"\n"):
        lines = lines[:-1]
    it = iter(lines)
    rows = declared = None
    for ln in it:
        if len(ln) <= 3 or b"%" in ln:
            continue
        r, c, declared = fields(ln, False)
        rows = min(r, c) if r > c else r
        break
    if rows is None:
        return 0, []
    out = []
    for ln in it:
        if len(ln) <= 3 or b"%" in ln or len(ln) >= 1000:
            continue
        r, c, v = fields(ln, True)
        if v == 0 or r > rows or c > rows:
            continue
        if len(out) >= declared:
            break
        out.append((r - 1, c - 1, v))
    return rows, out


def test_parallel_reader_equals_the_serial_rules(sg, tmp_path):
    """The reader parses the file in parallel chunks; on a file with every oddity the rules know (CRLF, blank and short
    lines, comments, a missing value, signs and exponents, out-of-range and zero entries, more entries than declared, no
    final newline) it must return exactly what the line-by-line reader returns."""
    rng = np.random.default_rng(5)
    n = 3000
    body = []
    for k in range(260000):                       # enough lines for the chunked path
        i, j = int(rng.integers(1, n + 1)), int(rng.integers(1, n + 1))
        kind = k % 97
        if kind == 0: body.append(b"%% comment %d" % k)
        elif kind == 1: body.append(b"")
        elif kind == 2: body.append(b"%d %d" % (i, j))                      # missing value -> 0 -> dropped
        elif kind == 3: body.append(b"%d %d 0.0" % (i, j))
        elif kind == 4: body.append(b"%d %d 1.5\r" % (i, j))               # CRLF file
        elif kind == 5: body.append(b"%d %d -2.5e-3" % (i, j))
        elif kind == 6: body.append(b"%d %d +7" % (n + 3, j))               # out of range
        elif kind == 7: body.append(b"  %d\t%d\t  .125" % (i, j))
        elif kind == 8: body.append(b"7 8")                                 # <= 3 characters
        elif kind == 9: body.append(b"%d %d 3.0 " % (i, j) + b"x" * 1000)   # >= 1000 characters
        else: body.append(b"%d %d %.17g" % (i, j, rng.uniform(-1, 1)))
    declared = 200000                                                       # fewer than the ~241 k surviving entries
    text = b"%%MatrixMarket matrix coordinate real general\n% c\n" + b"%d %d %d\n" % (n, n, declared) + b"\n".join(body)   # no final newline
    p = tmp_path / "odd.mtx"
    with open(p, "wb") as f:
        f.write(text)
    rows, ref = _reference_reader(text)
    assert rows == n and len(ref) == declared
    import ctypes
    L = sg.lib()
    L.soglu_debug_read_mtx.restype = ctypes.c_int64
    L.soglu_debug_read_mtx.argtypes = [ctypes.c_char_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    i = np.zeros(declared + 10, dtype=np.int32); j = np.zeros_like(i); v = np.zeros(declared + 10); meta = np.zeros(2, dtype=np.int64)
    cnt = L.soglu_debug_read_mtx(str(p).encode(), len(i), i.ctypes.data, j.ctypes.data, v.ctypes.data, meta.ctypes.data)
    assert cnt == declared and meta[0] == n and meta[1] == 0
    ri, rj, rv = (np.array(x) for x in zip(*ref))
    np.testing.assert_array_equal(i[:cnt], ri)
    np.testing.assert_array_equal(j[:cnt], rj)
    np.testing.assert_array_equal(v[:cnt], rv)
