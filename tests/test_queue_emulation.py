"""CPU stress test of the shared high-priority ready queue (executor option hi_shared): tests/emu/emu_queue.cpp runs the
kernel's claim loop (csrc/device/ready_queue.cuh, the same header) with host threads as CTAs on random DAGs -- every
task exactly once, after its predecessors, and every scheduler terminates."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shared_queue_protocol(tmp_path):
    exe = str(tmp_path / "emu_queue")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.run([cxx, "-O2", "-std=c++17", "-pthread", "-o", exe, os.path.join(ROOT, "tests", "emu", "emu_queue.cpp")], check=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout + r.stderr
