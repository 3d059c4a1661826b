"""CPU check of the blocked diagonal-block kernel (csrc/device/lu_blocked.cuh, the executor's T_LU / T_LLT tasks):
tests/emu/emu_lub.cpp compiles the device header for the host and runs it with one host thread per CUDA thread
(pthread barriers for __syncwarp / bar.sync, emulated DMMA fragments and shuffles) against a plain no-pivoting LU
with the reference's pivot clamp (MatrixStdDouble.cpp:2745) and against L^-1 L = I, U U^-1 = I."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_blocked_lu_emulation(tmp_path):
    exe = str(tmp_path / "emu_lub")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.run([cxx, "-O1", "-std=c++17", "-pthread", "-o", exe, os.path.join(ROOT, "tests", "emu", "emu_lub.cpp")], check=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout + r.stderr
