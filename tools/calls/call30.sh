cd $GRAFT_REPO_ROOT
w=lap3d_100
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 8 --steps 2 --warmup 3 --workload $w --no-cpu-baseline --trace > gpurun_out/bench_n8_$w.json 2> gpurun_out/bench_n8_$w.err
tail -2 gpurun_out/bench_n8_$w.json | cut -c1-1500
grep "^rank" gpurun_out/bench_n8_$w.err
