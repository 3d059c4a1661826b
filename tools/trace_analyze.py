#!/usr/bin/env python3
"""Per-task timeline of the persistent executor: where does the critical path spend its time?
usage: trace_analyze.py <kind> <dims>   (runs on the GPU with option trace=1)"""
import ctypes, os, sys, tempfile
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import soglu_b200 as sg

TYPES = {0: "gemm", 1: "sub", 2: "lu", 3: "llt", 4: "lowerInv", 5: "upperInv"}

def main():
    kind = sys.argv[1]; dims = [int(a) for a in sys.argv[2].split("x")]
    tmp = tempfile.mkdtemp(); path = os.path.join(tmp, "a.mtx")
    sg.write_stencil_mtx(kind, path, *dims)
    p = sg.Problem.from_mtx(path)
    ctx = sg.Context(0)
    for kv in sys.argv[3:]:
        k, v = kv.split("="); ctx.set_option(k, int(v))
    ctx.load(p)
    ctx.factor()
    ctx.set_option("trace", 1)
    fs = ctx.factor()
    L = sg.lib()
    L.soglu_debug_trace.restype = ctypes.c_int64
    L.soglu_debug_trace.argtypes = [ctypes.c_void_p] * 5
    nt = L.soglu_debug_trace(ctx.h, None, None, None, None)
    tr = np.zeros((nt, 6), dtype=np.uint64); info = np.zeros((nt, 4), dtype=np.int32); sp = np.zeros(nt + 1, dtype=np.int32)
    L.soglu_debug_trace(ctx.h, tr.ctypes.data_as(ctypes.c_void_p), info.ctypes.data_as(ctypes.c_void_p), sp.ctypes.data_as(ctypes.c_void_p), None)
    succ = np.zeros(sp[-1], dtype=np.int32)
    L.soglu_debug_trace(ctx.h, None, None, None, succ.ctypes.data_as(ctypes.c_void_p))
    t = tr[:, :5].astype(np.int64)
    t0 = t[:, 1].min()
    t = (t - t0) * 1e-3  # us
    pub, clm, lod, cmp_, sig = t.T
    pub = np.where(tr[:, 0] == 0, 0.0, pub)
    total = sig.max()
    print("traced factor: %.3f ms device (events %.3f ms), %d tasks, %d levels" % (total * 1e-3, fs["seconds"] * 1e3, nt, info[:, 2].max() + 1))
    print("%-9s %8s %7s | %8s %8s %8s %8s  (us, mean)" % ("type", "count", "pairs", "seen->go", "load", "compute", "signal"))
    for ty, nm in TYPES.items():
        m = info[:, 0] == ty
        if not m.any(): continue
        print("%-9s %8d %7.1f | %8.2f %8.2f %8.2f %8.2f" % (nm, m.sum(), info[m, 1].mean(), (clm - pub)[m].mean(), (lod - clm)[m].mean(), (cmp_ - lod)[m].mean(), (sig - cmp_)[m].mean()))
    # critical path: walk back from the last task through the predecessor that finished last
    pred_ptr = np.zeros(nt + 1, dtype=np.int64)
    src_of = np.repeat(np.arange(nt), np.diff(sp))
    order = np.argsort(succ, kind="stable")
    preds = src_of[order]
    cnt = np.bincount(succ, minlength=nt); pred_ptr[1:] = np.cumsum(cnt)
    late = np.argsort(-sig)[:5]
    print("last tasks to finish: " + "; ".join("#%d %s level %d deps %d preds %d sig %.1f us" % (c, TYPES[info[c, 0]], info[c, 2], info[c, 3], pred_ptr[c + 1] - pred_ptr[c], sig[c]) for c in late))
    print("tasks without stamps: seen %d, issued %d, loaded %d, computed %d, signalled %d" % tuple(int((tr[:, k] == 0).sum()) for k in range(5)))
    cur = int(np.argmax(sig)); chain = []
    while True:
        chain.append(cur)
        ps = preds[pred_ptr[cur]:pred_ptr[cur + 1]]
        if len(ps) == 0: break
        cur = int(ps[np.argmax(sig[ps])])
    chain = chain[::-1]
    cat = {"q-wait": 0.0, "load": 0.0, "compute": 0.0, "signal": 0.0, "publish-gap": 0.0}
    bytype = {}
    for a, b in zip(chain[:-1], chain[1:]):
        cat["publish-gap"] += max(0.0, pub[b] - sig[a]) if pub[b] > 0 else 0.0
    for c in chain:
        cat["q-wait"] += clm[c] - pub[c]; cat["load"] += lod[c] - clm[c]; cat["compute"] += cmp_[c] - lod[c]; cat["signal"] += sig[c] - cmp_[c]
        nm = TYPES[info[c, 0]]; d = bytype.setdefault(nm, [0, 0.0, 0.0, 0]); d[0] += 1; d[1] += sig[c] - pub[c]; d[2] += cmp_[c] - lod[c]; d[3] += int(info[c, 1])
    print("critical chain: %d tasks, spans %.3f ms of %.3f ms" % (len(chain), (sig[chain[-1]] - clm[chain[0]]) * 1e-3, total * 1e-3))
    print("  by phase (ms):", {k: round(v * 1e-3, 3) for k, v in cat.items()})
    print("  by type: ", {k: (v[0], "%.2f us/task total" % (v[1] / v[0]), "%.2f us compute" % (v[2] / v[0]), "%.1f pairs/task" % (v[3] / v[0])) for k, v in bytype.items()})
    # distribution of the chain's GEMM tasks by accumulation-chain length
    gl = np.array([info[c, 1] for c in chain if info[c, 0] == 0])
    if len(gl):
        print("  chain GEMM tasks by pairs: " + ", ".join("%s: %d" % (lab, int(((gl >= lo) & (gl < hi)).sum())) for lab, lo, hi in (("1", 1, 2), ("2-3", 2, 4), ("4-7", 4, 8), ("8-15", 8, 16), ("16-31", 16, 32), ("32+", 32, 10 ** 9))))
        print("  compute time of the chain's GEMM tasks by pairs (ms): " + ", ".join("%s: %.2f" % (lab, sum((cmp_[c] - lod[c]) for c in chain if info[c, 0] == 0 and lo <= info[c, 1] < hi) * 1e-3) for lab, lo, hi in (("1", 1, 2), ("2-3", 2, 4), ("4-7", 4, 8), ("8-15", 8, 16), ("16-31", 16, 32), ("32+", 32, 10 ** 9))))
    n = max(1, len(chain))
    # (static order: column 0 of the trace = the moment the waiting scheduler saw the task's counter at zero)
    print("  means along the chain (us): detection (predecessor's reds issued -> successor's scheduler sees zero) %.2f, seen -> issue %.2f, operand load %.2f, signal phase %.2f"
          % (cat["publish-gap"] / n, cat["q-wait"] / n, cat["load"] / n, cat["signal"] / n))
    # SM utilisation: busy = sum over tasks of (sig - lod) / (n_sm * total)
    busy = (sig - lod).sum()
    print("math-warp busy fraction: %.3f (148 SMs)" % (busy / (148 * total)))
    g = info[:, 0] == 0
    print("gemm compute per pair: %.3f us (ideal 2.08 at FP64 peak)" % ((cmp_ - lod)[g].sum() / info[g, 1].sum()))

if __name__ == "__main__":
    main()
