#!/usr/bin/env python3
"""Cycle counts of the diagonal-block kernels in isolation (needs a GPU)."""
import ctypes, os, sys, tempfile
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import soglu_b200 as sg
tmp = tempfile.mkdtemp(); path = os.path.join(tmp, "a.mtx")
sg.write_stencil_mtx("lap2d", path, 64)
p = sg.Problem.from_mtx(path); ctx = sg.Context(0); ctx.load(p); ctx.factor()
cyc = (ctypes.c_longlong * 12)()
L = sg.lib(); L.soglu_debug_diag_bench.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
rc = L.soglu_debug_diag_bench(ctx.h, 20, cyc); assert rc == 0, L.soglu_last_error()
print("cycles incl. write-out: lu+Linv+Uinv fused %d (%.2f us at 1.965 GHz)  lu only %d  standalone lowerInv %d  upperInv %d" % (cyc[0], cyc[0] / 1965.0, cyc[1], cyc[2], cyc[3]))
