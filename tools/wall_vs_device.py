#!/usr/bin/env python3
"""Wall time of the ABI calls next to the device time they report (host overhead per call)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np, gen_mtx, soglu_b200 as sg
kind, dims = sys.argv[1], [int(a) for a in sys.argv[2:]]
n, r, c, v = gen_mtx.generate(kind, *dims)
p = sg.Problem.from_coo(n, r, c, v, gen_mtx.rhs(n))
ctx = sg.Context(0); ctx.load(p); ctx.factor(); ctx.solve(p, refine=1)
for it in range(3):
    t0 = time.perf_counter(); fs = ctx.factor(); t1 = time.perf_counter()
    x, ss = ctx.solve(p); t2 = time.perf_counter()
    x, sr = ctx.solve(p, refine=1); t3 = time.perf_counter()
    print("factor wall %.1f ms device %.1f ms | solve wall %.1f device %.1f | solve refine=1 wall %.1f device %.1f" % (
        (t1 - t0) * 1e3, fs["seconds"] * 1e3, (t2 - t1) * 1e3, ss["seconds"] * 1e3, (t3 - t2) * 1e3, sr["seconds"] * 1e3), flush=True)
