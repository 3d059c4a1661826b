"""world_size-2 gloo worker for tests/test_bench_dist.py: exercises bench.py's multi-rank
plumbing (rank env parsing, max-over-ranks timing, replica aggregation) on CPU."""
import json
import os
import sys

import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

rank, local_rank, world = bench.dist_env()
dist.init_process_group("gloo")
dist.barrier()
mine = 0.25 * (rank + 1)                      # rank 1 is the slow one
tmax = bench.reduce_max(mine)
val = bench.aggregate_gflops(1e9, tmax, world)
dist.barrier()
for r in range(world):          # one rank at a time so the lines never interleave
    if r == rank:
        print(json.dumps({"rank": rank, "local_rank": local_rank, "world": world, "tmax": tmax, "value": val}), flush=True)
    dist.barrier()
dist.destroy_process_group()
