#!/usr/bin/env python3
"""Scratch GPU check: my factor+solve vs the unmodified reference (oracle/_ref/ref_harness)."""
import os, subprocess, sys, tempfile, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import soglu_b200 as sg

def run(kind, dims, mode=0, blocks=True, fuse=1):
    tmp = tempfile.mkdtemp()
    path = os.path.join(tmp, "a.mtx")
    sg.write_stencil_mtx(kind, path, *dims)
    out = os.path.join(tmp, "ref"); os.makedirs(out)
    env = dict(os.environ, OMP_NUM_THREADS="16")
    t = time.time()
    r = subprocess.run([os.path.join(os.path.dirname(__file__), "..", "oracle", "_ref", "ref_harness"), path, out] + (["--blocks"] if blocks else []),
                       env=env, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    print(r.stdout.strip().splitlines()[-1], "wall %.2f" % (time.time() - t))
    xref = np.fromfile(out + "/x.f64")
    t = time.time(); p = sg.Problem.from_mtx(path); print("plan wall %.2f" % (time.time() - t))
    ctx = sg.Context(0)
    ctx.set_option("exec_mode", mode); ctx.set_option("fuse_sub", fuse)
    t = time.time(); ctx.load(p); print("load wall %.2f" % (time.time() - t))
    t = time.time(); fs = ctx.factor(); print("factor wall %.2f" % (time.time() - t), fs)
    fs = ctx.factor(); print("factor again", fs["seconds"], "GFLOP/s %.1f" % (fs["flops"] / fs["seconds"] * 1e-9))
    x, ss = ctx.solve(p); print("solve", ss["seconds"], "GB/s %.1f" % (ss["bytes"] / ss["seconds"] * 1e-9))
    rel = np.linalg.norm(x - xref) / np.linalg.norm(xref)
    print("%s %s mode %d: rel diff vs reference %.3e  nan %d" % (kind, dims, mode, rel, int(np.isnan(x).sum())))
    if blocks:
        L = np.fromfile(out + "/L.i32", dtype=np.int32).reshape(-1, 3); Lv = np.fromfile(out + "/L.f64").reshape(-1, 64, 64)
        U = np.fromfile(out + "/U.i32", dtype=np.int32).reshape(-1, 3); Uv = np.fromfile(out + "/U.f64").reshape(-1, 64, 64)
        worst = 0
        for ids, vals, nm in ((L, Lv, "L"), (U, Uv, "U")):
            for k in range(min(len(ids), 400)):
                mine = ctx.get_block(ids[k, 0])
                d = np.abs(mine - vals[k]).max() / max(1e-300, np.abs(vals[k]).max())
                worst = max(worst, d)
        print("worst factor block rel diff %.3e" % worst)
    return rel

if __name__ == "__main__":
    kind = sys.argv[1]; dims = [int(a) for a in sys.argv[2].split("x")]
    mode = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    blocks = int(sys.argv[4]) if len(sys.argv) > 4 else 1
    fuse = int(sys.argv[5]) if len(sys.argv) > 5 else 1
    run(kind, dims, mode, bool(blocks), fuse)
