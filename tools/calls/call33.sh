cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/bench_n1_lap3d_100.json 2> gpurun_out/bench_n1_lap3d_100.err; tail -1 gpurun_out/bench_n1_lap3d_100.json | cut -c1-2500
for w in lap3d_64 banded_200k nine2d_1024; do
timeout 300 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1_$w.json 2> gpurun_out/bench_n1_$w.err; tail -1 gpurun_out/bench_n1_$w.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); c=d['config']; print('$w: factor %.1f ms solve %.1f step %.1f e2e %.1f resid %.2e x %s' % (c['factor_ms'], c['solve_ms'], d['ms_per_step'], d['e2e']['ms_per_step'], d['accuracy']['residual_rel'], d['x_sha256'][:16]))"
done
timeout 300 python tools/trace_analyze.py lap3d 64x64x64 2>&1 | tail -14
