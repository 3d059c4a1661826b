cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "two_gpu or two_gpus" 2>&1 | tail -5
for w in lap3d_100; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 3 --warmup 3 --workload $w --no-cpu-baseline --trace > gpurun_out/bench_n2_$w.json 2> gpurun_out/bench_n2_$w.err
tail -4 gpurun_out/bench_n2_$w.json | cut -c1-1400
grep -v "^$" gpurun_out/bench_n2_$w.err | tail -5
done
