cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
timeout 120 python tools/diag_bench.py
timeout 300 python tools/option_sweep.py lap3d 64
timeout 300 python tools/option_sweep.py nine2d 1024
timeout 300 python tools/option_sweep.py banded 200000
