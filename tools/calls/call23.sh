cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "golden or factor_blocks or executor_modes or pool_recycling or watchdog or diag_warnings or slack_split" 2>&1 | tail -5
timeout 300 python tools/trace_analyze.py lap3d 64x64x64 2>&1 | tail -16
timeout 300 python tools/option_sweep.py lap3d 64
timeout 300 python tools/option_sweep.py nine2d 1024
timeout 300 python tools/option_sweep.py banded 200000
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --trace 2>&1 | tail -12
