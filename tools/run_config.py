#!/usr/bin/env python3
"""Factor + solve one BASELINE.json config on the GPU from an in-memory COO matrix and report timings,
residuals (raw and after one refinement step) and, optionally, the difference to the unmodified reference.
usage: run_config.py <kind> <dims...> [--ref] [--sym] [key=value executor options]
  kinds: lap2d NX [NY] | lap3d NX [NY NZ] | nine2d NX [NY] | banded N [W] [K]"""
import os, subprocess, sys, tempfile, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import gen_mtx
import soglu_b200 as sg

args = [a for a in sys.argv[1:] if "=" not in a and not a.startswith("--")]
opts = [a for a in sys.argv[1:] if "=" in a]
want_ref = "--ref" in sys.argv
want_sym = "--sym" in sys.argv          # symmetric storage (lower triangle) -> the planner's LL^T path
kind, dims = args[0], [int(a) for a in args[1:]]
t = time.time(); n, r, c, v = gen_mtx.generate(kind, *dims); b = gen_mtx.rhs(n); print("generate %.1f s  n=%d nnz=%d" % (time.time() - t, n, len(v)), flush=True)
t = time.time()
if want_sym:
    low = r >= c
    p = sg.Problem.from_coo(n, r[low], c[low], v[low], b, symmetric=True)
else:
    p = sg.Problem.from_coo(n, r, c, v, b)
print("reorder+plan %.1f s%s" % (time.time() - t, "  (symmetric storage, LL^T path)" if want_sym else ""), flush=True)
print(p.log.strip().splitlines()[0])
print("ops %d storage %d L %d U %d stages %d flops %.4e" % (p.size("n_ops"), p.size("storage"), p.size("n_L"), p.size("n_U"), p.size("max_stage"), p.f64("flops")[0]), flush=True)
ctx = sg.Context(0)
for kv in opts:
    k, val = kv.split("="); ctx.set_option(k, int(val))
t = time.time(); ctx.load(p); fs = ctx.factor(); print("load + first factor wall %.1f s" % (time.time() - t), flush=True)
fs = ctx.factor()
print("factor %.4f s  %.1f GFLOP/s  tasks %d launches %d pool %.1f GB" % (fs["seconds"], fs["flops"] / fs["seconds"] * 1e-9, fs["tasks"], fs["kernel_launches"], fs["pool_blocks"] * 34816e-9), flush=True)
x, ss = ctx.solve(p); print("solve %.4f s  %.1f GB/s" % (ss["seconds"], ss["bytes"] / ss["seconds"] * 1e-9), flush=True)
xr, _ = ctx.solve(p, refine=1)
def resid(xv):
    ax = np.zeros(n); np.add.at(ax, r, v * xv[c]); return np.linalg.norm(ax - b) / np.linalg.norm(b)
print("residual ||Ax-b||/||b||: raw %.3e, after 1 refinement %.3e   nan %d   x[0:3] %s" % (resid(x), resid(xr), int(np.isnan(x).sum()), x[:3]), flush=True)
if want_ref:
    tmp = tempfile.mkdtemp(); path = os.path.join(tmp, "a.mtx"); gen_mtx.write_mtx(path, n, r, c, v)
    out = os.path.join(tmp, "ref"); os.makedirs(out)
    res = subprocess.run([os.path.join(ROOT, "oracle", "_ref", "ref_harness"), path, out], capture_output=True, text=True, env=dict(os.environ, OMP_NUM_THREADS="16"))
    print(res.stdout.strip().splitlines()[-1])
    xref = np.fromfile(out + "/x.f64")
    print("reference: rel diff of x %.3e   reference residual %.3e   op list identical: %s" % (
        np.linalg.norm(x - xref) / np.linalg.norm(xref), resid(xref),
        np.array_equal(p.i32("ops"), np.fromfile(out + "/ops_fine.i32", dtype=np.int32).reshape(-1, 8))))
