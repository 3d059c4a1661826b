cd $GRAFT_REPO_ROOT
for w in lap3d_64 banded_200k nine2d_1024 lap3d_100; do
  extra="--no-cpu-baseline"; [ $w = lap3d_100 ] && extra=""
  timeout 900 python bench.py --steps 3 --warmup 3 --workload $w $extra > gpurun_out/bench_n1_$w.json 2> gpurun_out/bench_n1_$w.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_n1_$w.json").read().strip().splitlines()[-1]); c=d["config"]
    print("$w N=1: step %.1f factor %.1f solve %.1f e2e %.1f value %.0f resid %.2e raw %.2e x %s" % (d["ms_per_step"], c["factor_ms"], c["solve_ms"], d["e2e"]["ms_per_step"], d["value"], d["accuracy"]["residual_rel"], d["accuracy"]["residual_rel_raw_solve"], d["x_sha256"][:16]))
except Exception as e:
    print("$w failed", e); print(open("gpurun_out/bench_n1_$w.err").read()[-800:])
PY
done
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
