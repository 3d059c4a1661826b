cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
timeout 120 python tools/diag_bench.py
timeout 300 python tools/option_sweep.py lap3d 64 | grep -v chain_cuts
timeout 300 python tools/option_sweep.py nine2d 1024 | grep -v chain_cuts
timeout 300 python tools/option_sweep.py banded 200000 | grep -v chain_cuts
( time timeout 1500 python bench.py --steps 3 --warmup 3 ) > gpurun_out/bench_100.json 2> gpurun_out/bench_100.err
tail -3 gpurun_out/bench_100.err; cat gpurun_out/bench_100.json
