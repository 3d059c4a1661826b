cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
( time timeout 1500 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/bench_ref_100.json 2> gpurun_out/bench_ref_100.err
tail -3 gpurun_out/bench_ref_100.err; cut -c1-1500 gpurun_out/bench_ref_100.json
( time timeout 1500 python bench.py --steps 5 --warmup 3 ) > gpurun_out/bench_100.json 2> gpurun_out/bench_100.err
tail -3 gpurun_out/bench_100.err; cat gpurun_out/bench_100.json
cp /tmp/soglu_ref_cache/lap3d_100_x.f64 gpurun_out/ref_lap3d_100_x.f64 2>/dev/null
