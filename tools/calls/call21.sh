cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "golden or factor_blocks or executor_modes or pool_recycling or watchdog or diag_warnings or slack_split" 2>&1 | tail -5
timeout 300 python tools/option_sweep.py lap3d 64
timeout 300 python tools/option_sweep.py nine2d 1024
timeout 300 python tools/option_sweep.py banded 200000
for h in 1 0; do timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --opt static_order=$h | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); c=d['config']; print('lap3d_100 static_order=$h: factor %.1f ms solve %.1f step %.1f first_call %.1f x %s' % (c['factor_ms'], c['solve_ms'], d['ms_per_step'], c['first_call_s'], d['x_sha256'][:16]))"; done
timeout 300 python tools/trace_analyze.py lap3d 64 64 64 2>&1 | tail -12
