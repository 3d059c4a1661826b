"""CPU check of the two-pivots-per-barrier diagonal-block kernel: tests/emu/emu_diag2.cpp compiles the device
header csrc/device/diag2.cuh for the host and runs its per-thread code for all 256 threads, interval by interval,
against a plain no-pivoting LU / Cholesky with the reference's pivot clamps (MatrixStdDouble.cpp:2745, 2640)."""
import os, subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_diag2_emulation(tmp_path):
    exe = str(tmp_path / "emu_diag2")
    subprocess.run([("/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"), "-O1", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "emu", "emu_diag2.cpp")], check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout + r.stderr
