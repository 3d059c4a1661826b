set -u
cd $GRAFT_REPO_ROOT
free -g | head -2; nproc; lscpu | grep "Model name"
mkdir -p /tmp/ref100 && python - <<'PY'
import sys; sys.path.insert(0,'.')
import soglu_b200 as sg
sg.write_stencil_mtx("lap3d", "/tmp/ref100/lap3d_100.mtx", 100, 100, 100)
PY
( /usr/bin/time -v env OMP_NUM_THREADS=16 oracle/_ref/ref_harness /tmp/ref100/lap3d_100.mtx /tmp/ref100 --lean > gpurun_out/ref100.log 2>&1; cp /tmp/ref100/x.f64 gpurun_out/ref_lap3d_100_x.f64; cp /tmp/ref100/meta.txt gpurun_out/ref100_meta.txt ) &
timeout 120 tools/bin/lu_lab
wait
tail -25 gpurun_out/ref100.log
