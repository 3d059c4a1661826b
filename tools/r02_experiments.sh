#!/bin/bash
# First GPU call of round 2: validates and measures the options that were written without a GPU in round 1
# (hi_shared, chain_cuts, lu_mode=1).  Every step runs under its own timeout so that a hang in an unmeasured
# path cannot take the box down.  usage (from the repo root):
#   gpurun --timeout 1500 -- 'bash tools/r02_experiments.sh > gpurun_out/r02_experiments.log 2>&1'
set -u
mkdir -p gpurun_out
run() { echo; echo "=== $*"; timeout "${T:-300}" "$@"; echo "--- exit $?"; }

python -c "import __graft_entry__ as g; g.build()"
# 1. correctness of the new paths (opt-in tests)
T=600 SOGLU_EXPERIMENTAL=1 run python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "shared_priority or chain_cuts or blocked_diagonal or slack_split or prefetch"
# 2a. can the L2 -> SM path deliver the operands of the DMMA loop at its peak rate?
/usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o gpurun_out/operand_bw tools/operand_bw.cu
T=120 run gpurun_out/operand_bw
# 2. the diagonal-block kernels in isolation (cycles)
T=120 run python tools/diag_bench.py
# 3. latency-bound configs: planned once, every option in a fresh context (risky variants last)
T=420 run python tools/r02_sweep.py lap3d 64
T=600 run python tools/r02_sweep.py nine2d 1024
T=420 run python tools/r02_sweep.py banded 200000
# 4. the headline (100^3, throughput-bound): shared high-priority queue, thresholds around the model's optimum
for s in 300 1000 3000; do
  T=420 run python bench.py --steps 2 --warmup 3 --no-cpu-baseline --opt hi_shared=$s
done
T=420 run python bench.py --steps 2 --warmup 3 --no-cpu-baseline --opt hi_shared=1000 --opt lu_mode=1
