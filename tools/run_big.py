#!/usr/bin/env python3
"""Factor + solve a large stencil problem on the GPU and report timings + residual.
usage: run_big.py lap3d 100 [key=value options]"""
import os, sys, tempfile, time, resource
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import soglu_b200 as sg

kind = sys.argv[1]; dims = [int(a) for a in sys.argv[2].split("x")]
tmp = tempfile.mkdtemp(); path = os.path.join(tmp, "a.mtx")
t = time.time(); sg.write_stencil_mtx(kind, path, *dims); print("write %.1f s" % (time.time() - t), flush=True)
t = time.time(); p = sg.Problem.from_mtx(path); print("reorder+plan %.1f s" % (time.time() - t), flush=True)
print(p.log.strip().splitlines()[0]); print("ops %d storage %d L %d stages %d flops %.4e" % (p.size("n_ops"), p.size("storage"), p.size("n_L"), p.size("max_stage"), p.f64("flops")[0]), flush=True)
ctx = sg.Context(0)
for kv in sys.argv[3:]:
    k, v = kv.split("="); ctx.set_option(k, int(v))
t = time.time(); ctx.load(p); print("load %.1f s" % (time.time() - t), flush=True)
t = time.time(); fs = ctx.factor(); print("first factor wall %.1f s (compile + upload + run)" % (time.time() - t), fs, flush=True)
fs = ctx.factor(); print("factor %.4f s  %.1f GFLOP/s  launches %d pool %.1f GB" % (fs["seconds"], fs["flops"] / fs["seconds"] * 1e-9, fs["kernel_launches"], fs["pool_blocks"] * 34816e-9), flush=True)
x, ss = ctx.solve(p); print("solve %.4f s  %.1f GB/s" % (ss["seconds"], ss["bytes"] / ss["seconds"] * 1e-9), flush=True)
xr, sr = ctx.solve(p, refine=1); print("solve+1 refinement %.4f s" % sr["seconds"], flush=True)
n = p.size("dim")
b = 1.0 + 0.25 * (np.arange(n) % 7)
if kind == "lap3d":
    nx = dims[0]; ny = dims[1] if len(dims) > 1 else nx; nz = dims[2] if len(dims) > 2 else nx
    X = x.reshape(nz, ny, nx); ax = 6.0 * X
    ax[1:] -= X[:-1]; ax[:-1] -= X[1:]; ax[:, 1:] -= X[:, :-1]; ax[:, :-1] -= X[:, 1:]; ax[:, :, 1:] -= X[:, :, :-1]; ax[:, :, :-1] -= X[:, :, 1:]
    print("residual ||Ax-b||/||b|| = %.3e  nan %d  x[0:3] %s" % (np.linalg.norm(ax.ravel() - b) / np.linalg.norm(b), int(np.isnan(x).sum()), x[:3]))
    X = xr.reshape(nz, ny, nx); ax = 6.0 * X
    ax[1:] -= X[:-1]; ax[:-1] -= X[1:]; ax[:, 1:] -= X[:, :-1]; ax[:, :-1] -= X[:, 1:]; ax[:, :, 1:] -= X[:, :, :-1]; ax[:, :, :-1] -= X[:, :, 1:]
    print("residual after 1 refinement step = %.3e   rel change of x %.3e" % (np.linalg.norm(ax.ravel() - b) / np.linalg.norm(b), np.linalg.norm(xr - x) / np.linalg.norm(x)))
print("host maxrss %.1f GB" % (resource.getrusage(resource.RUSAGE_SELF).ru_maxrss / 1e6))
