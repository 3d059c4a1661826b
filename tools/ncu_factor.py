#!/usr/bin/env python3
"""One factorisation of a bench workload, nothing else (target of the ncu captures; run under `ncu --replay-mode application`).
usage: ncu_factor.py <workload of bench.py>"""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import soglu_b200 as sg
name = sys.argv[1] if len(sys.argv) > 1 else "lap3d_100"
tmp = tempfile.mkdtemp(prefix="soglu_ncu_")
p = sg.Problem.from_mtx(bench.write_workload(sg, name, tmp))
ctx = sg.Context(0)
ctx.set_option("watchdog_ms", 0)      # a profiled pass can be arbitrarily slow
ctx.load(p)
fs = ctx.factor()
print("%s: factor %.1f ms, %d launches" % (name, fs["seconds"] * 1e3, fs["kernel_launches"]))
