"""Host front-end (reader, GPS ordering, two-level planner) against golden vectors recorded
from the unmodified reference: integer outputs must be bit-exact."""
import numpy as np
import pytest

from conftest import GOLDEN_CASES, load_golden, write_case_mtx

# known answers of the unmodified reference
KNOWN = {
    # name: (gps levels, width, start, coarse emitted, coarse kept, fine emitted, fine kept, stages, storage)
    "lap2d_64": (127, 64, 4095, 183, 73, 1887, 893, 317, 2270),
    "lap2d_64_sym": (127, 64, 4095, 114, 54, 1259, 638, 317, 1514),
    "lap3d_24": (70, 432, 13823, 772, 301, 19447, 16302, 1423, 13650),
}
OPCOUNT = {  # lu, lowerInv, upperInv, sub, mul, mulneg, llt, mult
    "lap2d_64": {1: 128, 2: 63, 3: 63, 4: 450, 8: 189},
    "lap2d_64_sym": {10: 128, 2: 63, 4: 321, 11: 126},
    "lap3d_24": {1: 256, 2: 362, 3: 362, 4: 3472, 8: 10101, 9: 1749},
}


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_plan_bit_exact(sg, tmp_path, name):
    g = load_golden(name)
    p = sg.Problem.from_mtx(write_case_mtx(name, tmp_path))
    assert p.size("dim") == int(g["dim"])
    assert p.size("symmetric") == int(g["symmetric"])
    assert [p.size("gps_levels"), p.size("gps_width"), p.size("gps_last"), p.size("dim"), p.size("gps_start")] == list(g["gps"][:5])
    np.testing.assert_array_equal(p.i32("perm_new2old"), g["perm_new2old"])
    np.testing.assert_array_equal(p.i32("perm_old2new"), g["perm_old2new"])
    np.testing.assert_array_equal(p.i32("coarse_ops"), g["coarse_ops"])
    np.testing.assert_array_equal(p.i32("ops"), g["ops"])          # op src src2 result result2 stage group seq
    assert p.size("storage") == int(g["storage"]) and p.size("coarse_storage") == int(g["coarse_storage"])
    assert [p.size("coarse_emitted"), p.size("fine_emitted")] == list(g["ops_emitted"])
    np.testing.assert_array_equal(p.i32("stage"), g["stage"])
    np.testing.assert_array_equal(p.i32("laststage"), g["laststage"])
    np.testing.assert_array_equal(p.i32("inputs"), g["inputs"])
    np.testing.assert_array_equal(p.i32("L"), g["L"])
    if len(g["U"]):
        np.testing.assert_array_equal(p.i32("U"), g["U"])
    else:
        assert p.size("n_U") == 0
    np.testing.assert_array_equal(p.f64("b_perm"), g["b_perm"])    # permuted rhs padded with 1.0


@pytest.mark.parametrize("name", sorted(KNOWN))
def test_survey_known_answers(sg, tmp_path, name):
    lv, w, st, ce, ck, fe, fk, stages, storage = KNOWN[name]
    p = sg.Problem.from_mtx(write_case_mtx(name, tmp_path))
    assert (p.size("gps_levels"), p.size("gps_width"), p.size("gps_start")) == (lv, w, st)
    assert (p.size("coarse_emitted"), p.size("coarse_ops"), p.size("fine_emitted"), p.size("n_ops")) == (ce, ck, fe, fk)
    ops = p.i32("ops")
    assert len(np.unique(ops[:, 5])) + 1 == stages or ops[:, 5].max() == stages
    assert p.size("storage") == storage
    cnt = {int(k): int(v) for k, v in zip(*np.unique(ops[:, 0], return_counts=True))}
    assert cnt == OPCOUNT[name]


def test_op_list_invariants(sg, tmp_path):
    """SURVEY.md Appendix E invariants the executor relies on."""
    p = sg.Problem.from_mtx(write_case_mtx("lap3d_24", tmp_path))
    ops = p.i32("ops")
    inputs = set(p.i32("inputs")[:, 0].tolist())
    stage_of_writer = {}
    kinds = {}
    for o, s, s2, r, r2, stg, _, _ in ops.tolist():
        for res in ([r, r2] if o == 1 else [r]):
            assert res not in inputs
            kinds.setdefault(res, set()).add(o)
            stage_of_writer[res] = max(stage_of_writer.get(res, 0), stg)
        assert r not in (s, s2)
    for res, k in kinds.items():
        assert len(k) == 1
    multi = [r for r in kinds if (ops[:, 3] == r).sum() > 1]
    assert all(kinds[r] <= {8, 9, 11} for r in multi)
    # every reader runs in a later stage than all writers of what it reads
    for o, s, s2, r, r2, stg, _, _ in ops.tolist():
        for src in (s, s2):
            if src > 0 and src in stage_of_writer:
                assert stage_of_writer[src] < stg


def test_from_coo_matches_from_mtx(sg, tmp_path):
    import gen_mtx
    n, r, c, v = gen_mtx.generate("lap2d", 50, 37)
    b = gen_mtx.rhs(n)
    p1 = sg.Problem.from_coo(n, r, c, v, b)
    p2 = sg.Problem.from_mtx(write_case_mtx("lap2d_50x37", tmp_path))
    np.testing.assert_array_equal(p1.i32("ops"), p2.i32("ops"))
    np.testing.assert_array_equal(p1.i32("perm_new2old"), p2.i32("perm_new2old"))
    np.testing.assert_array_equal(p1.f64("input_vals"), p2.f64("input_vals"))


def test_stencil_writer_matches_python_generator(sg, tmp_path):
    """soglu_write_stencil_mtx (used by bench.py at full size) writes the same files."""
    import gen_mtx
    for kind, dims in (("lap2d", (9, 7)), ("nine2d", (6, 5)), ("lap3d", (5, 4, 3))):
        n, r, c, v = gen_mtx.generate(kind, *dims)
        a = str(tmp_path / ("py_%s.mtx" % kind))
        bpath = str(tmp_path / ("c_%s.mtx" % kind))
        gen_mtx.write_mtx(a, n, r, c, v)
        sg.write_stencil_mtx(kind, bpath, *dims)
        assert open(a).read() == open(bpath).read()
        assert open(a.replace(".mtx", "_b.mtx")).read() == open(bpath.replace(".mtx", "_b.mtx")).read()
