"""Shared fixtures.  GPU tests are marked `gpu`; everything else runs on a CPU-only box.

The oracle (oracle/liboracle.so, and oracle/_ref/ when prebuilt) is test infrastructure: it is
loaded here and nowhere in the product."""
import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
GOLDEN_CASES = ["lap2d_64", "lap2d_64_sym", "lap2d_50x37", "nine2d_40", "lap3d_13x11x9", "lap3d_24", "lap3d_16_sym", "banded_3000"]
GOLDEN_GEN = {
    "lap2d_64": ("lap2d", (64,), False), "lap2d_64_sym": ("lap2d", (64,), True), "lap2d_50x37": ("lap2d", (50, 37), False),
    "nine2d_40": ("nine2d", (40,), False), "lap3d_13x11x9": ("lap3d", (13, 11, 9), False), "lap3d_24": ("lap3d", (24,), False),
    "lap3d_16_sym": ("lap3d", (16,), True),
}


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _ensure_built():
    import __graft_entry__ as ge
    ge.build()


@pytest.fixture(scope="session")
def sg():
    _ensure_built()
    import soglu_b200
    return soglu_b200


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))


def write_case_mtx(name, tmpdir):
    """Write the .mtx/_b.mtx of a golden case and return its path."""
    import gen_mtx
    path = os.path.join(str(tmpdir), name + ".mtx")
    if name in GOLDEN_GEN:
        kind, dims, sym = GOLDEN_GEN[name]
        n, r, c, v = gen_mtx.generate(kind, *dims)
        gen_mtx.write_mtx(path, n, r, c, v, sym)
    else:
        g = load_golden(name)
        gen_mtx.write_mtx(path, int(g["dim"]), g["coo_i"].astype(np.int64), g["coo_j"].astype(np.int64), g["coo_v"], False)
    return path


class Oracle:
    """ctypes view of oracle/liboracle.so (the CPU restatement of the hot path)."""

    def __init__(self):
        path = os.path.join(ROOT, "oracle", "liboracle.so")
        if not os.path.exists(path):
            subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "oracle"], check=True, capture_output=True)
        L = ctypes.CDLL(path)
        vp, i64 = ctypes.c_void_p, ctypes.c_int64
        L.oracle_create.restype = vp
        L.oracle_create.argtypes = [i64]
        L.oracle_destroy.argtypes = [vp]
        L.oracle_set_inputs.argtypes = [vp, i64, vp, vp]
        L.oracle_factor.restype = i64
        L.oracle_factor.argtypes = [vp, i64, vp, vp, vp, vp, vp]
        L.oracle_get_block.argtypes = [vp, ctypes.c_int32, vp]
        L.oracle_solve.argtypes = [vp, i64, vp, i64, vp, ctypes.c_int32, ctypes.c_int, vp, vp]
        self.L = L

    def run(self, problem):
        """Factor + solve a planned problem; returns (x_ext, handle) -- caller frees handle."""
        p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
        ops = problem.i32("ops")
        inputs = problem.i32("inputs")
        vals = problem.f64("input_vals")
        h = self.L.oracle_create(problem.size("storage"))
        ids = np.arange(1, len(inputs) + 1, dtype=np.int32)   # input_vals is indexed by id-1
        self.L.oracle_set_inputs(h, len(ids), p(ids), p(vals))
        cols = [np.ascontiguousarray(ops[:, k]) for k in range(5)]
        opc = np.ascontiguousarray(ops[:, 0].astype(np.uint8))
        rc = self.L.oracle_factor(h, len(ops), p(cols[1]), p(cols[2]), p(opc), p(cols[3]), p(cols[4]))
        assert rc == 0, "oracle cannot execute op %d" % (rc - 1)
        Lf = np.ascontiguousarray(problem.i32("L"))
        Uf = np.ascontiguousarray(problem.i32("U"))
        b = problem.f64("b_perm")
        x = np.zeros_like(b)
        rc = self.L.oracle_solve(h, len(Lf), p(Lf), len(Uf), p(Uf) if len(Uf) else None, problem.size("block_rows"),
                                 problem.size("symmetric"), p(b), p(x))
        assert rc == 0, "oracle solve failed rc=%d" % rc
        return x, h

    def block(self, h, block_id):
        out = np.empty((64, 64))
        rc = self.L.oracle_get_block(h, int(block_id), out.ctypes.data_as(ctypes.c_void_p))
        return out if rc == 0 else None

    def free(self, h):
        self.L.oracle_destroy(h)


@pytest.fixture(scope="session")
def oracle():
    return Oracle()


def ref_harness_path():
    p = os.path.join(ROOT, "oracle", "_ref", "ref_harness")
    if not os.path.exists(p):
        return None
    try:
        with open("/proc/cpuinfo") as f:
            if "avx512f" not in f.read():
                return None
    except OSError:
        return None
    return p


def unpermute(problem, x_ext):
    """x in the original ordering from the permuted extended solution (GPSOrder.cpp:41-53)."""
    return x_ext[problem.i32("perm_old2new")]
