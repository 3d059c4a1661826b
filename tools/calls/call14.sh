cd $GRAFT_REPO_ROOT
nvidia-smi -L | wc -l
run() { w=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 8 --steps 3 --warmup 3 --workload $w --no-cpu-baseline "$@" > gpurun_out/bench_n8_$w.json 2> gpurun_out/bench_n8_$w.err
grep "^rank" gpurun_out/bench_n8_$w.err | cut -c1-260; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_n8_$w.json").read().strip().splitlines()[-1])
    c=d["config"]; print("$w N=8: step %.1f ms factor %.1f ms solve %.1f ms value %.0f GFLOP/s e2e %.1f ms first_call %.1f s plan %.1f s x_sha %s resid %.2e segs %d" % (d["ms_per_step"], c["factor_ms"], c["solve_ms"], d["value"], d["e2e"]["ms_per_step"], c["first_call_s"], c["host_plan_s"], d["x_sha256"][:16], d["accuracy"]["residual_rel"], c["segments"]))
except Exception as e:
    print("$w failed", e); print(open("gpurun_out/bench_n8_$w.err").read()[-1500:])
PY
}
run lap3d_100 --trace
run nine2d_1024
run banded_200k
