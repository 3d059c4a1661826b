cd $GRAFT_REPO_ROOT
echo skip tests
for w in lap3d_64 lap3d_100; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 3 --warmup 3 --workload $w --no-cpu-baseline --trace > gpurun_out/bench_n2_$w.json 2> gpurun_out/bench_n2_$w.err
grep "^rank" gpurun_out/bench_n2_$w.err | cut -c1-300; python - <<PY
import json
d=json.loads(open("gpurun_out/bench_n2_$w.json").read().strip().splitlines()[-1])
c=d["config"]; print("$w N=2: step %.1f ms factor %.1f ms solve %.1f ms value %.0f GFLOP/s e2e %.1f ms first_call %.1f s plan %.1f s x_sha %s resid %.2e" % (d["ms_per_step"], c["factor_ms"], c["solve_ms"], d["value"], d["e2e"]["ms_per_step"], c["first_call_s"], c["host_plan_s"], d["x_sha256"][:16], d["accuracy"]["residual_rel"]))
PY
done
