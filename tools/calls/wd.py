import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, "tools")
import soglu_b200 as sg, gen_mtx
for dims in ([24], [40]):
    n, r, c, v = gen_mtx.generate("lap3d", *dims)
    p = sg.Problem.from_coo(n, r, c, v, gen_mtx.rhs(n))
    ctx = sg.Context(0); ctx.load(p)
    print(dims, "factor ms", ctx.factor()["seconds"] * 1e3)
    ctx.set_option("watchdog_ms", 1)
    try:
        print("with 1 ms watchdog:", ctx.factor()["seconds"] * 1e3)
    except sg.SogluError as e:
        print("raised:", e)
    ctx.set_option("watchdog_ms", 60000)
    print("after:", ctx.factor()["seconds"] * 1e3)
