#!/usr/bin/env python3
"""Timed host model of the persistent executor (csrc/device/model.cpp) on the compiled task graph of a problem.
No GPU needed.  usage: model.py <kind> <dims...> [max_slots=N] [split=0|1] [chains=K] [grid=PRxPCxNB] [param=value ...]
  kinds as tools/run_config.py; params: the fields of ModelParams (n_ctas t_pair t_pair_half t_pair_quarter
  t_lu_fused t_lu t_llt_fused t_inv t_sub t_epilogue t_release t_poll t_poll_hit t_desc t_load t_launch t_cas hi_slack_us);
  policy=0 executor FIFO | 1 ideal list scheduling | 2 two queues, every CTA serves the high-priority one first."""
import ctypes, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import gen_mtx
import soglu_b200 as sg

PARAMS = ["n_ctas", "t_pair", "t_pair_half", "t_pair_quarter", "t_lu_fused", "t_lu", "t_llt_fused", "t_inv", "t_sub", "t_epilogue",
          "t_release", "t_poll", "t_poll_hit", "t_desc", "t_load", "t_launch", "t_cas", "hi_slack_us", "t_release_remote", "t_load_remote", "t_launch_dist"]


def model(p, split=1, max_slots=0, chains=0, policy=0, compile_hi_slack=0, grid=(1, 1, 16), split_slack=0, split_width=0, **kw):
    L = sg.lib()
    L.soglu_debug_model.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
    opts = np.array([split, max_slots, grid[0], grid[1], grid[2], chains, policy, compile_hi_slack, split_slack, split_width], dtype=np.int64)
    par = np.full(len(PARAMS), np.nan)
    for k, v in kw.items():
        par[PARAMS.index(k)] = v
    out = np.zeros(36)
    rc = L.soglu_debug_model(p.h, opts.ctypes.data, par.ctypes.data, len(par), out.ctypes.data)
    if rc:
        raise RuntimeError(L.soglu_last_error().decode())
    return dict(makespan_ms=out[0] * 1e-3, critical_ms=out[1] * 1e-3, busy_ms_per_cta=out[2] * 1e-3 / (kw.get("n_ctas", 148) * grid[0] * grid[1]),
                tasks=int(out[3]), segments=int(out[4]), pairs=int(out[5]), hi=int(out[6]),
                cp_ms=out[7] * 1e-3, cp_early_ms=out[8] * 1e-3, cuts=int(out[9]), cp_cut_ms=out[10] * 1e-3, cuts_applied=int(out[11]), remote_loads=int(out[12]), remote_releases=int(out[13]), dual_tasks=int(out[34]), dual_covered_pairs=int(out[35]),
                chain=dict(kinds="gemm64 gemm32 gemm16 lu sub inv".split(), tasks=out[14:20].astype(int).tolist(), math_ms=(out[20:26] * 1e-3).round(1).tolist(),
                           pairs=out[26:32].astype(int).tolist(), overhead_ms=round(out[32] * 1e-3, 1), remote_hops=int(out[33])))


def problem(kind, dims):
    n, r, c, v = gen_mtx.generate(kind, *dims)
    return sg.Problem.from_coo(n, r, c, v, gen_mtx.rhs(n))


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if "=" not in a]
    kv = dict(a.split("=") for a in sys.argv[1:] if "=" in a)
    p = problem(args[0], [int(a) for a in args[1:]])
    split = int(kv.pop("split", 1)); ms = int(kv.pop("max_slots", 0)); ch = int(kv.pop("chains", 0)); pol = int(kv.pop("policy", 0)); chs = int(kv.pop("compile_hi_slack", 0)); grid = tuple(int(x) for x in kv.pop("grid", "1x1x16").split("x")); ssl = int(kv.pop("split_slack", 0)); sw = int(kv.pop("split_width", 0))
    t = time.time()
    r = model(p, split, ms, ch, pol, chs, grid, ssl, sw, **{k: float(v) for k, v in kv.items()})
    print(r, "(%.1f s)" % (time.time() - t))
