cd $GRAFT_REPO_ROOT
timeout 300 python tools/trace_analyze.py lap3d 64x64x64 2>&1 | tail -14
timeout 300 python tools/option_sweep.py lap3d 64
timeout 300 python tools/option_sweep.py nine2d 1024
timeout 300 python tools/option_sweep.py banded 200000
