#!/usr/bin/env python3
"""One BASELINE config, planned once, factorised with every listed set of executor options (fresh context per variant);
prints the factor time and the deviation of x from the default run.
usage: option_sweep.py <kind> <dims...>      e.g.  option_sweep.py lap3d 64"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import gen_mtx
import soglu_b200 as sg

NGPU = int(os.environ.get("SWEEP_GPUS", "1"))          # > 1: the in-process group of soglu_create
VARIANTS = [("default", {})]
for spec in os.environ.get("SWEEP_VARIANTS", "").split(";"):      # e.g. "order_alpha=30;dist_nb=8,mirror_min=2"
    if spec:
        VARIANTS.append((spec, {kv.split("=")[0]: int(kv.split("=")[1]) for kv in spec.split(",")}))

kind, dims = sys.argv[1], [int(a) for a in sys.argv[2:]]
n, r, c, v = gen_mtx.generate(kind, *dims)
t = time.time()
p = sg.Problem.from_coo(n, r, c, v, gen_mtx.rhs(n))
print("%s %s: n=%d ops=%d planned in %.1f s" % (kind, dims, n, p.size("n_ops"), time.time() - t), flush=True)
x0 = None
for name, opts in VARIANTS:
    ctx = sg.Context(0) if NGPU == 1 else sg.Context(n_gpus=NGPU)
    for k, val in opts.items():
        ctx.set_option(k, val)
    t = time.time()
    ctx.load(p)
    ctx.factor()
    first = time.time() - t
    best = min(ctx.factor()["seconds"] for _ in range(3))
    fs = ctx.factor()
    x, ss = ctx.solve(p)
    if x0 is None:
        x0 = x
    dev = float(np.linalg.norm(x - x0) / np.linalg.norm(x0))
    print("%-55s factor %8.2f ms  (%.1f GFLOP/s, %d tasks, first call %.1f s)  solve %.2f ms  |x - x_default|/|x| = %.1e  nan %d"
          % (name, best * 1e3, fs["flops"] / best * 1e-9, fs["tasks"], first, ss["seconds"] * 1e3, dev, int(np.isnan(x).sum())), flush=True)
    ctx.close()
