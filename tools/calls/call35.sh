cd $GRAFT_REPO_ROOT
w=nine2d_1024
timeout 200 python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1_$w.json 2> gpurun_out/bench_n1_$w.err; tail -1 gpurun_out/bench_n1_$w.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); c=d['config']; print('$w: factor %.1f ms solve %.1f step %.1f e2e %.1f resid %.2e refine %s x %s' % (c['factor_ms'], c['solve_ms'], d['ms_per_step'], d['e2e']['ms_per_step'], d['accuracy']['residual_rel'], d['accuracy'].get('refine_steps'), d['x_sha256'][:16]))" || tail -5 gpurun_out/bench_n1_$w.err
