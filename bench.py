#!/usr/bin/env python3
"""Benchmark of the hot path: FP64 block-LU factor + triangular solve (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload lap3d_64]

One "step" = one numeric factorisation (soglu_factor: the whole operation DAG, one persistent
kernel) + one forward/back solve (soglu_solve) of the same planned problem.  The op list is
planned once on the host (bit-exact reproduction of the reference planner; not timed, like the
reference's own "plan time").  Input blocks are resident in HBM when the timed region starts;
the block pool (tens of GB) is far larger than the 126 MB L2, so no explicit L2 flush is needed.

value  = algorithmic GFLOP/s (dense-block convention of SURVEY.md 8d, from the op list) over
         the wall time of K steps, bracketed by barrier + cuda synchronize, max over ranks.
e2e    = same metric through the C ABI from HOST buffers: every step re-uploads the dense input
         blocks and the right-hand side (H2D) and reads x back (D2H) inside the timed region.
N > 1  = ONE factorisation sharded over the N GPUs (2D block-cyclic block ownership, owner computes,
         remote operands pulled over NVLink into local mirrors; one process per GPU, peers mapped with
         CUDA IPC; NCCL only for bootstrap / barriers); the solve runs on rank 0 over peer memory.
         Total work is fixed -> "scaling": "strong".  value = op-list FLOPs / max-over-ranks time.

--impl reference times the UNMODIFIED reference (oracle/_ref/ref_harness, its own OpenMP path on
the host cores, "kernel time" + "solve triangled") on the same workload; if the prebuilt
reference is unusable on this host it falls back to the oracle port on a reduced sample.
"""
import argparse
import json
import os
import re
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
if "WORLD_SIZE" in os.environ:   # torchrun pins OMP_NUM_THREADS=1; the host planner is OpenMP-parallel
    os.environ["OMP_NUM_THREADS"] = str(max(1, (os.cpu_count() or 8) // max(1, int(os.environ["WORLD_SIZE"]))))

WORKLOADS = {
    # name: (stencil kind, dims, BASELINE.json config)
    "lap2d_256": ("lap2d", (256, 256, 0), "configs[0]: 2D 5-point Laplacian 256x256 (n=65,536)"),
    "lap3d_64": ("lap3d", (64, 64, 64), "configs[1]: 3D 7-point Laplacian 64^3 (n=262,144) on 1 B200"),
    "nine2d_1024": ("nine2d", (1024, 1024, 0), "configs[3]: 2D 9-point stencil 1024x1024 (n=1,048,576)"),
    "lap3d_100": ("lap3d", (100, 100, 100), "configs[4]: 3D 7-point Laplacian 100^3 (n=1,000,000)"),
    "lap3d_24": ("lap3d", (24, 24, 24), "smoke-sized 3D 7-point Laplacian 24^3"),
    "lap3d_40": ("lap3d", (40, 40, 40), "profiling-sized 3D 7-point Laplacian 40^3 (n=64,000)"),
}
FP64_PEAK_FALLBACK_TFLOPS = 36.98   # tools/fp64_peak.cu (DMMA m8n8k4) on this pool, profiles/r01_fp64_peak.txt


def dist_env():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.proc = None
        self.lines = []
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = max(mx, float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        busy = [s for s in sm if s > 0.5 * mx] or sm
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


def write_workload(sg, name, tmp):
    kind, dims, _ = WORKLOADS[name]
    path = os.path.join(tmp, name + ".mtx")
    sg.write_stencil_mtx(kind, path, *dims)
    return path


def solve_flops_bytes(problem):
    nblk = problem.size("n_L") + (problem.size("n_U") if problem.size("n_U") else problem.size("n_L"))
    return 2.0 * 4096.0 * nblk, 32768.0 * nblk + 8.0 * 3.0 * problem.size("n_ext")


def measure_fp64_peak():
    exe = os.path.join(ROOT, "tools", "fp64_peak")
    try:
        if not os.path.exists(exe):
            subprocess.run(["nvcc", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-o", exe, exe + ".cu"], check=True, capture_output=True)
        out = subprocess.run([exe], capture_output=True, text=True, timeout=120).stdout
        m = re.search(r"DMMA m8n8k4\s+acc8\s*:\s*([0-9.]+) TFLOP/s", out)
        if m:
            return float(m.group(1)), "measured live by tools/fp64_peak.cu (DMMA m8n8k4, register resident)"
    except Exception:
        pass
    return FP64_PEAK_FALLBACK_TFLOPS, "tools/fp64_peak.cu on this pool earlier (profiles/r01_fp64_peak.txt); MEASURED_PEAKS.json has no FP64 entry"


def run_reference_harness(path, threads):
    """Unmodified reference on the host cores; returns dict with factor/solve seconds or None."""
    harness = os.path.join(ROOT, "oracle", "_ref", "ref_harness")
    if not os.path.exists(harness):
        return None
    try:
        if "avx512f" not in open("/proc/cpuinfo").read():
            return None
    except OSError:
        return None
    out = tempfile.mkdtemp(prefix="soglu_ref_")
    env = dict(os.environ, OMP_NUM_THREADS=str(threads))
    r = subprocess.run([harness, path, out], env=env, capture_output=True, text=True)
    m = re.search(r"HARNESS factor_s ([0-9.]+) solve_s ([0-9.]+) total_s ([0-9.]+) max_rhs_error (\S+)", r.stdout)
    for f in os.listdir(out):
        os.unlink(os.path.join(out, f))
    os.rmdir(out)
    if r.returncode != 0 or not m:
        return None
    return {"factor_s": float(m.group(1)), "solve_s": float(m.group(2)), "total_s": float(m.group(3)), "max_rhs_error": float(m.group(4))}


def reduce_max(seconds, device=None):
    """max over ranks of a host-measured duration (no-op without an initialised process group)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(seconds)
    t = torch.tensor([seconds], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def aggregate_gflops(flops_per_rank, seconds_max, world):
    """whole-job throughput of `world` replicas that each did flops_per_rank in <= seconds_max."""
    return world * flops_per_rank / seconds_max * 1e-9


def host_threads():
    return max(1, min(16, os.cpu_count() or 1))   # MAXTHREAD 16 is the reference's hard cap (const.h:23)


def reference_arm(args, rank, world):
    if rank != 0:
        return 0
    import soglu_b200 as sg
    tmp = tempfile.mkdtemp(prefix="soglu_bench_")
    # the reference needs ~110 GB and ~10 min for 100^3: it is timed on a bounded sample of the same
    # stencil family (default 64^3, ~12 s per run) and compared in GFLOP/s
    name = args.workload if args.workload in ("lap2d_256", "lap3d_24", "lap3d_40", "lap3d_64") else args.cpu_sample
    path = write_workload(sg, name, tmp)
    prob = sg.Problem.from_mtx(path)            # only for the FLOP count of the op list
    flops = float(prob.f64("flops")[0])
    sflops, _ = solve_flops_bytes(prob)
    threads = host_threads()
    kind, sample = "reference", "%s (%s), one full run of the unmodified reference per step; its own timers around BlockPlanner::calculate + BlockPlanner::solve" % (
        name, "the benchmarked workload" if name == args.workload else "bounded sample of the %s stencil family" % args.workload)
    times = []
    ok = True
    for it in range(args.warmup + args.steps):
        r = run_reference_harness(path, threads) if ok else None
        if r is None:
            ok = False
            break
        if it >= args.warmup:
            times.append(r["factor_s"] + r["solve_s"])
    if not ok:
        # oracle port on a reduced sample (scalar C, 1 core)
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from conftest import Oracle
        kind, threads, name2 = "port", 1, "lap3d_24"
        p2 = sg.Problem.from_mtx(write_workload(sg, name2, tmp))
        flops = float(p2.f64("flops")[0])
        sflops, _ = solve_flops_bytes(p2)
        sample = "oracle port (scalar C) on %s: oracle/_ref unusable on this host" % name2
        orc = Oracle()
        times = []
        for it in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            _, h = orc.run(p2)
            orc.free(h)
            if it >= args.warmup:
                times.append(time.perf_counter() - t0)
    t = sum(times) / len(times)
    val = (flops + sflops) / t * 1e-9
    line = {
        "impl": "reference", "metric": "fp64_lu_factor_solve_gflops", "value": val, "unit": "GFLOP/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOADS[args.workload][2], "name": args.workload},
        "cpu_baseline": {"value": val, "unit": "GFLOP/s", "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="lap3d_100", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-sample", default="lap3d_64", choices=sorted(WORKLOADS), help="workload the CPU reference is timed on (bounded sample)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--opt", action="append", default=[], metavar="KEY=VALUE",
                    help="executor option passed to soglu_set_option before the first factorisation (hi_shared, chain_cuts, lu_mode, ...); recorded in config.options")
    args = ap.parse_args()
    rank, local_rank, world = dist_env()
    if args.impl == "reference":
        return reference_arm(args, rank, world)

    import numpy as np
    import torch
    import __graft_entry__ as ge
    if rank == 0:
        ge.build()
    use_dist = world > 1
    if use_dist:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist.barrier()
    import soglu_b200 as sg
    sg.set_host_threads(max(1, (os.cpu_count() or 8) // max(1, world)))   # every rank plans the whole problem

    tmp = tempfile.mkdtemp(prefix="soglu_bench_r%d_" % rank)
    path = write_workload(sg, args.workload, tmp)
    t0 = time.perf_counter()
    prob = sg.Problem.from_mtx(path)
    t_plan = time.perf_counter() - t0
    flops = float(prob.f64("flops")[0])
    sflops, sbytes = solve_flops_bytes(prob)

    dev = local_rank if use_dist else 0

    def barrier():
        torch.cuda.synchronize(dev)
        if use_dist:
            dist.barrier()
            torch.cuda.synchronize(dev)

    ctx = sg.Context(dev, rank, world)
    for kv in args.opt:
        k, v = kv.split("=")
        ctx.set_option(k, int(v))
    t0 = time.perf_counter()
    ctx.load(prob)
    if use_dist:
        blob = torch.from_numpy(ctx.dist_export()).cuda()          # compiles + allocates this GPU's share
        allb = [torch.empty_like(blob) for _ in range(world)]
        dist.all_gather(allb, blob)
        ctx.dist_import(torch.stack(allb).cpu().numpy())
        factor = lambda: ctx.factor_dist(barrier)
    else:
        factor = ctx.factor
    first = factor()                           # includes the one-time task-graph compilation + upload
    t_first = time.perf_counter() - t0
    x = None
    if rank == 0:
        x, _ = ctx.solve(prob)
    barrier()

    # ---- device-resident steps ---------------------------------------------------------------
    def step():
        fs = factor()
        ss = {"kernel_launches": 0, "seconds": 0.0}
        if rank == 0:
            _, ss = ctx.solve(prob)
        if use_dist:
            barrier()                           # peers keep their factor blocks mapped until rank 0 has solved
        return fs, ss
    for _ in range(max(args.warmup, 3)):
        step()
    sampler = ClockSampler(dev) if rank == 0 else None
    barrier()
    t0 = time.perf_counter()
    launches, f_dev, s_dev = 0, [], []
    for _ in range(args.steps):
        fs, ss = step()
        launches += fs["kernel_launches"] * fs.get("segments", 1) + ss["kernel_launches"]
        f_dev.append(fs["seconds"])
        s_dev.append(ss["seconds"])
    barrier()
    elapsed = time.perf_counter() - t0
    clocks = sampler.stop() if sampler else None
    elapsed = reduce_max(elapsed, "cuda" if use_dist else None)
    t_factor_all = reduce_max(sum(f_dev) / len(f_dev), "cuda" if use_dist else None)   # device time of the executor launches, max over ranks
    step_s = elapsed / args.steps
    value = aggregate_gflops(flops + sflops, step_s, 1 if use_dist else world)   # sharded: the work is done once

    # ---- end to end from host buffers (pinned), H2D + D2H inside the timed region ---------------
    n_in = prob.size("n_input")
    # the matrix goes up as the planner's duplicate-free entry list (block, position, value): pinned host buffers
    ent_in = torch.from_numpy(prob.i32("entry_block") - 1).pin_memory()
    ent_pos = torch.from_numpy(prob.i32("entry_pos")).pin_memory()
    vals_host = torch.from_numpy(prob.f64("entry_val")).pin_memory()
    ent_in_np, ent_pos_np = ent_in.numpy(), ent_pos.numpy()
    ids = np.arange(1, n_in + 1, dtype=np.int32)
    b_host = torch.from_numpy(prob.f64("b_perm")).pin_memory()
    x_host = torch.empty_like(b_host).pin_memory()
    vals_np, b_np, x_np = vals_host.numpy(), b_host.numpy(), x_host.numpy()
    L = sg.lib()
    import ctypes
    st = sg.Stats()

    def e2e_step():
        ctx.set_blocks_sparse(prob.size("storage"), ids, ent_in_np, ent_pos_np, vals_np)   # H2D: matrix entries (every rank: it keeps its share)
        factor()
        if rank == 0:
            rc = L.soglu_solve(ctx.h, b_np.ctypes.data_as(ctypes.c_void_p), x_np.ctypes.data_as(ctypes.c_void_p), ctypes.byref(st))  # H2D b, D2H x
            assert rc == 0
        if use_dist:
            barrier()
    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(1, min(args.steps, 3))
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    e2e_s = reduce_max(e2e_s, "cuda" if use_dist else None)
    h2d = float(vals_np.nbytes + ent_in_np.nbytes + ent_pos_np.nbytes + b_np.nbytes)
    d2h = float(x_np.nbytes)
    # the e2e result must be the same solution
    if rank == 0:
        x_e2e = x_np[prob.i32("perm_old2new")]
        assert np.array_equal(x_e2e, x), "end-to-end path produced a different solution"

    if rank == 0:
        peak, peak_how = measure_fp64_peak()
        t_factor = t_factor_all
        t_solve = sum(s_dev) / len(s_dev)
        achieved = flops / t_factor * 1e-12
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "r01_executor_traffic.json")
        if os.path.exists(tpath):
            try:
                tj = json.load(open(tpath))
                if tj.get("workload") == args.workload:
                    traffic = tj.get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        cpu = None
        if not args.no_cpu_baseline:
            threads = host_threads()
            cname = args.workload if args.workload in ("lap2d_256", "lap3d_24", "lap3d_40", "lap3d_64") else args.cpu_sample
            cpath, cflops = path, flops + sflops
            if cname != args.workload:
                cpath = write_workload(sg, cname, tmp)
                cprob = sg.Problem.from_mtx(cpath)
                cflops = float(cprob.f64("flops")[0]) + solve_flops_bytes(cprob)[0]
                cprob.close()
            r = run_reference_harness(cpath, threads)
            if r:
                cpu = {"value": cflops / (r["factor_s"] + r["solve_s"]) * 1e-9, "unit": "GFLOP/s", "cores": threads, "kind": "reference",
                       "sample": "%s%s, one full run of the unmodified reference (factor %.2f s + solve %.2f s, max rhs error %.2e)"
                                 % (cname, "" if cname == args.workload else " = bounded sample (the reference needs ~110 GB / ~10 min for %s)" % args.workload,
                                    r["factor_s"], r["solve_s"], r["max_rhs_error"])}
            else:
                sys.path.insert(0, os.path.join(ROOT, "tests"))
                from conftest import Oracle
                p2 = sg.Problem.from_mtx(write_workload(sg, "lap3d_24", tmp))
                orc = Oracle()
                t0 = time.perf_counter()
                _, h = orc.run(p2)
                orc.free(h)
                dt = time.perf_counter() - t0
                f2 = float(p2.f64("flops")[0]) + solve_flops_bytes(p2)[0]
                cpu = {"value": f2 / dt * 1e-9, "unit": "GFLOP/s", "cores": 1, "kind": "port", "sample": "oracle port (scalar C) on lap3d_24; oracle/_ref unusable on this host"}
        # accuracy of the benchmarked solution: raw and after one device-side refinement step (not timed)
        accuracy = None
        if WORKLOADS[args.workload][0] == "lap3d":
            nx, ny, nz = WORKLOADS[args.workload][1]
            bb = 1.0 + 0.25 * (np.arange(prob.size("dim")) % 7)

            def resid(xv):
                X = xv.reshape(nz, ny, nx)
                ax = 6.0 * X
                ax[1:] -= X[:-1]; ax[:-1] -= X[1:]; ax[:, 1:] -= X[:, :-1]; ax[:, :-1] -= X[:, 1:]; ax[:, :, 1:] -= X[:, :, :-1]; ax[:, :, :-1] -= X[:, :, 1:]
                return float(np.linalg.norm(ax.ravel() - bb) / np.linalg.norm(bb))
            xr, _ = ctx.solve(prob, refine=1)
            accuracy = {"residual_rel": resid(x), "residual_rel_after_1_refinement": resid(xr), "nan": int(np.isnan(x).sum()),
                        "note": "||Ax-b||/||b||, north-star gate 1e-12; refinement = FP64 residual + re-solve on the device"}
        n_segments = max(1, ctx.segments())
        line = {
            "metric": "fp64_lu_factor_solve_gflops", "value": value, "unit": "GFLOP/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": step_s * 1e3, "higher_is_better": True, "scaling": "strong" if use_dist else "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOADS[args.workload][2], "name": args.workload, "n": prob.size("dim"), "ops": prob.size("n_ops"),
                       "tasks": int(first["tasks"]), "pool_blocks": int(first["pool_blocks"]),
                       "parallelism": "1 GPU" if world == 1 else "one factorisation sharded over %d GPUs: 2D block-cyclic (%dx%d grid, 16x16-block squares), owner computes, NVLink peer pulls; solve on rank 0" % ((world,) + sg.default_grid(world)),
                       "segments": n_segments,   # executor launches per factorisation (pool recycling)
                       "l2": "inputs_exceed_l2 (block pool %.1f GB >> 126 MB L2)" % (first["pool_blocks"] * 34816 * 1e-9),
                       "factor_ms": t_factor * 1e3, "solve_ms": t_solve * 1e3, "factor_gflops": flops / t_factor * 1e-9,
                       "solve_gbs": sbytes / t_solve * 1e-9, "host_plan_s": t_plan, "first_call_s": t_first,
                       "options": dict(kv.split("=") for kv in args.opt)},
            "e2e": {"value": aggregate_gflops(flops + sflops, e2e_s, 1 if use_dist else world), "unit": "GFLOP/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_s * 1e3},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak * world, "unit": "TFLOP/s", "frac": achieved / (peak * world),
                         "traffic": traffic if world == 1 else None,
                         "kernel": "executor_kernel (the whole factorisation DAG: %d persistent launch(es) per step per GPU; achieved = op-list FLOPs / "
                                   "CUDA-event time of those launches, max over ranks)" % n_segments,
                         "peak_source": peak_how + ("; x %d GPUs" % world if world > 1 else "")},
            "cpu_baseline": cpu,
            "accuracy": accuracy,
            "clocks": clocks,
        }
        print(json.dumps(line))
    ctx.close()
    if use_dist:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
