cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 120 python tools/diag_bench.py
timeout 300 python tools/option_sweep.py lap3d 64 | head -2
timeout 300 python tools/option_sweep.py nine2d 1024 | head -2
timeout 300 python tools/option_sweep.py banded 200000 | head -2
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); c=d['config']; print('lap3d_100 N=1: factor %.1f ms solve %.1f step %.1f e2e %.1f  value %.0f frac %.3f x %s' % (c['factor_ms'], c['solve_ms'], d['ms_per_step'], d['e2e']['ms_per_step'], d['value'], d['roofline']['frac'], d['x_sha256'][:16]))"
