cd $GRAFT_REPO_ROOT
timeout 300 python tools/option_sweep.py lap3d 64
timeout 300 python tools/option_sweep.py nine2d 1024
timeout 300 python tools/option_sweep.py banded 200000
timeout 300 python tools/trace_analyze.py lap3d 64x64x64 | head -12
for lc in 1 0; do timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --opt lazy_claim=$lc | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); c=d['config']; print('lap3d_100 lazy_claim=$lc: factor %.1f ms solve %.1f step %.1f e2e %.1f' % (c['factor_ms'], c['solve_ms'], d['ms_per_step'], d['e2e']['ms_per_step']))"; done
