#!/usr/bin/env python3
"""Generate tests/golden/*.npz from the UNMODIFIED reference.

Runs oracle/_ref/ref_harness (reference objects compiled in place from /root/reference by
oracle/Makefile, driven through SOGLU::solveLU) on small synthetic matrices and records what
the parity tests compare against: GPS permutation, coarse and fine operation lists (op, src,
src2, result, result2, stage, groupNum, sequenceNum), per-block stage/laststage, input /
L / U block coordinates, the permuted padded rhs, x (original ordering) and the reference's
own console summary (GGPS line, op counts, max rhs error).

Only runs where /root/reference exists (the build container); the fixtures are committed.
    python tools/make_golden.py
"""
import os
import re
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import gen_mtx  # noqa: E402

HARNESS = os.path.join(ROOT, "oracle", "_ref", "ref_harness")
OUT = os.path.join(ROOT, "tests", "golden")

CASES = [
    # name, kind, dims, symmetric-file
    ("lap2d_64", "lap2d", (64,), False),
    ("lap2d_64_sym", "lap2d", (64,), True),
    ("lap2d_50x37", "lap2d", (50, 37), False),
    ("nine2d_40", "nine2d", (40,), False),
    ("lap3d_13x11x9", "lap3d", (13, 11, 9), False),
    ("lap3d_24", "lap3d", (24,), False),
    ("lap3d_16_sym", "lap3d", (16,), True),
    ("banded_3000", "banded", (3000, 200, 9), False),
]


def run_case(name, kind, dims, sym):
    tmp = tempfile.mkdtemp(prefix="golden_")
    path = os.path.join(tmp, name + ".mtx")
    n, r, c, v = gen_mtx.generate(kind, *dims)
    gen_mtx.write_mtx(path, n, r, c, v, sym)
    out = os.path.join(tmp, "out")
    os.makedirs(out)
    env = dict(os.environ, OMP_NUM_THREADS="8")
    res = subprocess.run([HARNESS, path, out], env=env, capture_output=True, text=True, check=True)
    log = res.stdout
    rd = lambda f, dt: np.fromfile(os.path.join(out, f), dtype=dt)
    meta = dict(line.split() for line in open(os.path.join(out, "meta.txt")) if len(line.split()) == 2)
    m = re.search(r"GGPS reorder: levels: (\d+) bandwidth: (\d+) last level count: (\d+) total accounted: (\d+) start from (\d+)", log)
    gps = np.array([int(g) for g in m.groups()], dtype=np.int64)
    reduced = np.array([int(x) for x in re.findall(r"reduced ops to: (\d+)", log)], dtype=np.int64)
    emitted = np.array([int(x) for x in re.findall(r"op count: (\d+)", log)], dtype=np.int64)
    data = dict(
        dim=np.int64(n), symmetric=np.int64(int(sym)),
        gps=gps, ops_reduced=reduced, ops_emitted=emitted,
        storage=np.int64(int(meta["storage"])), coarse_storage=np.int64(int(meta["coarse_storage"])),
        block_rows=np.int64(int(meta["block_rows"])), max_rhs_error=np.float64(float(meta["max_rhs_error"])),
        perm_new2old=rd("perm_new2old.i32", np.int32), perm_old2new=rd("perm_old2new.i32", np.int32),
        ops=rd("ops_fine.i32", np.int32).reshape(-1, 8), coarse_ops=rd("ops_coarse.i32", np.int32).reshape(-1, 8),
        stage=rd("stage.i32", np.int32), laststage=rd("laststage.i32", np.int32),
        inputs=rd("inputs.i32", np.int32).reshape(-1, 3), L=rd("L.i32", np.int32).reshape(-1, 3),
        U=rd("U.i32", np.int32).reshape(-1, 3), b_perm=rd("b_perm.f64", np.float64), x=rd("x.f64", np.float64),
    )
    if kind == "banded":   # RNG-dependent input: keep the matrix itself
        keep = r >= c if sym else np.ones(len(r), dtype=bool)
        data.update(coo_i=r[keep].astype(np.int32), coo_j=c[keep].astype(np.int32), coo_v=v[keep])
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **data)
    print("%-16s n=%6d ops=%7d storage=%7d x[0:3]=%s err=%.3g" % (name, n, len(data["ops"]), data["storage"], data["x"][:3], data["max_rhs_error"]))


def main():
    if not os.path.exists(HARNESS):
        sys.exit("oracle/_ref/ref_harness missing: run `make -C oracle ref` where /root/reference exists")
    os.makedirs(OUT, exist_ok=True)
    for case in CASES:
        run_case(*case)


if __name__ == "__main__":
    main()
