"""One rank of a multi-GPU factorisation (torchrun, NCCL): sharded factor, sharded solve (x on rank 0),
comparison with a single-GPU run of the same problem on rank 0.  Prints one JSON line on rank 0."""
import json
import os
import sys
import time

os.environ["OMP_NUM_THREADS"] = str(max(1, (os.cpu_count() or 8) // max(1, int(os.environ.get("WORLD_SIZE", "1")))))  # host planner threads

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import soglu_b200 as sg  # noqa: E402


def lap3d_residual(x, dims):
    nx = dims[0]; ny = dims[1] if len(dims) > 1 else nx; nz = dims[2] if len(dims) > 2 else nx
    X = x.reshape(nz, ny, nx); ax = 6.0 * X
    ax[1:] -= X[:-1]; ax[:-1] -= X[1:]; ax[:, 1:] -= X[:, :-1]; ax[:, :-1] -= X[:, 1:]; ax[:, :, 1:] -= X[:, :, :-1]; ax[:, :, :-1] -= X[:, :, 1:]
    b = 1.0 + 0.25 * (np.arange(x.size) % 7)
    return float(np.linalg.norm(ax.ravel() - b) / np.linalg.norm(b))


def run(kind, dims, steps=2):
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    path = "/tmp/soglu_dist_%s_%s.mtx" % (kind, "x".join(map(str, dims)))
    if rank == 0:
        sg.write_stencil_mtx(kind, path, *dims)
    dist.barrier()
    p = sg.Problem.from_mtx(path)
    ctx = sg.Context(local, rank, world)
    for kv in sys.argv[3:]:
        k, v = kv.split('='); ctx.set_option(k, int(v))
    ctx.load(p)
    blob = torch.from_numpy(ctx.dist_export()).cuda()
    allb = [torch.empty_like(blob) for _ in range(world)]
    dist.all_gather(allb, blob)
    ctx.dist_import(torch.stack(allb).cpu().numpy())
    info = ctx.dist_info()
    times = []
    x = None
    def barrier():
        torch.cuda.synchronize()
        dist.barrier()
    for _ in range(steps):
        t0 = time.perf_counter()
        fs = ctx.factor_dist(barrier)
        times.append(time.perf_counter() - t0)
        x, ss = ctx.solve(p)        # collective: every GPU solves the block rows it owns, rank 0 holds x
        dist.barrier()
    tt = torch.tensor([min(times)], dtype=torch.float64, device="cuda")
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    infos = [None] * world
    dist.all_gather_object(infos, info)
    if rank == 0 and os.environ.get("SOGLU_DIST_NO_REF"):
        print(json.dumps({"world": world, "kind": kind, "dims": dims, "factor_s_multi": float(tt.item()), "solve_s_multi": ss["seconds"],
                          "flops": fs["flops"], "tflops": fs["flops"] / float(tt.item()) * 1e-12, "per_rank": infos, "x0": x[:3].tolist(),
                          "residual": lap3d_residual(x, dims) if kind == "lap3d" else None}))
    elif rank == 0:
        ref = sg.Context(local)
        ref.load(p)
        f1 = ref.factor()
        f1 = ref.factor()
        x1, _ = ref.solve(p)
        rel = float(np.linalg.norm(x - x1) / np.linalg.norm(x1))
        print(json.dumps({"world": world, "kind": kind, "dims": dims, "factor_s_multi": float(tt.item()), "factor_s_single": f1["seconds"],
                          "solve_s_multi": ss["seconds"], "rel_diff_vs_single_gpu": rel, "bitwise_equal": bool(np.array_equal(x, x1)),
                          "flops": fs["flops"], "per_rank": infos}))
        ref.close()
    dist.barrier()
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    kind = sys.argv[1]
    dims = [int(a) for a in sys.argv[2].split("x")]
    run(kind, dims)
