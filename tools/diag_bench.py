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
print("cycles: lu+Linv+Uinv fused %d  lu only %d  lowerInv %d  upperInv %d   (per pivot: %.0f %.0f %.0f %.0f)" % (cyc[0], cyc[1], cyc[2], cyc[3], cyc[0] / 64, cyc[1] / 64, cyc[2] / 64, cyc[3] / 64))
print('blocked kernel (lu_blocked.cuh, incl. write-out): lu+Linv+Uinv %d  lu only %d cycles' % (cyc[10], cyc[11]))
print('lu-only variants (cycles): no-barrier %d | no-rcp %d | no-update %d | no-rcp+no-update %d | none of the three %d | full %d' % tuple(cyc[4:10]))
# numerical check of the bench outputs (slots 2..7) against numpy on the same block
import numpy as np
def blk(slot):
    # read by abusing get_block on ids is not possible for raw slots; use inputs: slot 1 = input id with slot 1
    return None
