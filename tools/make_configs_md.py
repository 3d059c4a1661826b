#!/usr/bin/env python3
"""profiles/r02_configs.md from the bench.py lines collected under gpurun_out/ (bench_n<N>_<workload>.json)."""
import json, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
names = {"lap3d_100": "3D 7-pt 100^3 (5)", "lap3d_64": "3D 7-pt 64^3 (2)", "nine2d_1024": "9-pt 1024^2 (4)", "banded_200k": "banded random 200 k (3)"}
out = ["# Round 2: every BASELINE config through bench.py, 1 / 2 / 8 B200 (one box)", "",
       "`python bench.py --workload W` (N = 1) and `torchrun --nproc-per-node N bench.py --gpus N --workload W` (N > 1), steps 3, warm-up 3.",
       "step = factor + solve with one refinement step (two block solves + one SpMV); x_sha256 = hash of the timed solution: equal across N",
       "for every workload, i.e. the sharded runs give bitwise the single-GPU x.", "",
       "| workload (BASELINE config) | GPUs | step ms | factor ms | solve ms (refine = 1) | e2e step ms | GFLOP/s | residual (refined / raw) | SM clock | x_sha256 |",
       "|---|---|---|---|---|---|---|---|---|---|"]
for w in names:
    first = True
    for n in (1, 2, 4, 8):
        p = os.path.join(ROOT, "gpurun_out", "bench_n%d_%s.json" % (n, w))
        try:
            d = json.loads(open(p).read().strip().splitlines()[-1])
        except (OSError, ValueError, IndexError):
            continue
        c, a, k = d["config"], d["accuracy"], d.get("clocks") or {}
        out.append("| %s | %d | %.1f | %.1f | %.1f | %.1f | %.0f | %.2e / %.2e | %s%s | %s |" % (
            names[w] if first else "", n, d["ms_per_step"], c["factor_ms"], c["solve_ms"], d["e2e"]["ms_per_step"], d["value"], a["residual_rel"],
            a["residual_rel_raw_solve"], k.get("sm_mhz"), (" " + ",".join(k.get("reasons") or [])) if k.get("reasons") else "", d["x_sha256"][:16]))
        first = False
out += ["", "Configs 2-4 are their dependency chain on one GPU already (`profiles/r02_chain_64.md`): more GPUs add NVLink hops to a chain that does",
        "not get shorter, as SURVEY 8(e) anticipated.  Config 5 is work-bound on one GPU and becomes the same chain from about two GPUs on",
        "(DESIGN.md section 7).", ""]
open(os.path.join(ROOT, "profiles", "r02_configs.md"), "w").write("\n".join(out))
print("\n".join(out))
