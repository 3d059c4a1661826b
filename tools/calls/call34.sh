cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "two_gpu or two_gpus" 2>&1 | tail -3
for w in nine2d_1024 lap3d_100; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 2 --warmup 3 --workload $w --no-cpu-baseline > gpurun_out/bench_n2_$w.json 2> gpurun_out/bench_n2_$w.err
tail -1 gpurun_out/bench_n2_$w.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); c=d['config']; print('$w N=2: factor %.1f ms solve %.1f step %.1f e2e %.1f resid %.2e refine %s x %s' % (c['factor_ms'], c['solve_ms'], d['ms_per_step'], d['e2e']['ms_per_step'], d['accuracy']['residual_rel'], d['accuracy'].get('refine_steps'), d['x_sha256'][:16]))" || tail -5 gpurun_out/bench_n2_$w.err
done
