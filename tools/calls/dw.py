import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, "tools")
import numpy as np, soglu_b200 as sg, gen_mtx
n, r, c, v = gen_mtx.generate("lap2d", 24)
v = v.copy(); v[(r == 100) & (c == 100)] = np.nan
p = sg.Problem.from_coo(n, r, c, v, gen_mtx.rhs(n))
ops = p.i32("ops"); print("ops:", {int(k): int((ops[:, 0] == k).sum()) for k in np.unique(ops[:, 0])})
ctx = sg.Context(0); ctx.load(p); fs = ctx.factor(); print("tasks", fs["tasks"], "warnings", ctx.diag_warnings())
x, _ = ctx.solve(p); print("nan in x:", int(np.isnan(x).sum()))
U = p.i32("U"); d = [int(i) for i, br, bc in U if br == bc]
for bid in d[:9]:
    b = ctx.get_block(bid); print("U diag block", bid, "nan", int(np.isnan(b).sum()))
