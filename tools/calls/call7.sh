cd $GRAFT_REPO_ROOT
python tools/wall_vs_device.py lap3d 64
mkdir -p /tmp/ref100 && python - <<'PY'
import sys; sys.path.insert(0,'.')
import soglu_b200 as sg
sg.write_stencil_mtx("lap3d", "/tmp/ref100/lap3d_100.mtx", 100, 100, 100)
PY
( time env OMP_NUM_THREADS=16 oracle/_ref/ref_harness /tmp/ref100/lap3d_100.mtx /tmp/ref100 --lean ) > gpurun_out/ref100.log 2>&1 &
PID=$!
while kill -0 $PID 2>/dev/null; do sleep 20; grep -E "VmRSS|VmHWM" /proc/$(pgrep -n ref_harness)/status 2>/dev/null | tr '\n' ' '; free -g | sed -n 2p; done
tail -30 gpurun_out/ref100.log; ls -la /tmp/ref100; dmesg 2>/dev/null | tail -5
cp /tmp/ref100/x.f64 gpurun_out/ref_lap3d_100_x.f64 2>/dev/null; cp /tmp/ref100/meta.txt gpurun_out/ref100_meta.txt 2>/dev/null
