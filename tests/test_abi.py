"""The C-ABI library loads without a GPU, exports every symbol include/soglu.h declares,
and refuses to run without a CUDA device instead of falling back to anything."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "soglu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(soglu_[a-zA-Z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(sg):
    lib = sg.lib()
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "libsoglu_b200.so does not export %s" % n
    assert sorted(sg.ABI_SYMBOLS) == names
    assert lib.soglu_abi_version() == 1


def test_no_cpu_fallback(sg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(sg.SogluError) as e:
        sg.Context(0)
    assert "no CUDA device" in str(e.value) or "CUDA" in str(e.value)


def test_product_does_not_touch_oracle():
    """Nothing under the package may import, link or mention the oracle."""
    pkg = os.path.join(ROOT, "sparse-operator-graph-lu_b200")
    for base, _, files in os.walk(pkg):
        if "build" in base:
            continue
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".h", ".cuh", "Makefile")):
                text = open(os.path.join(base, f), errors="replace").read()
                assert "liboracle" not in text and "oracle/" not in text and "oracle_" not in text, os.path.join(base, f)
    lib = os.path.join(pkg, "libsoglu_b200.so")
    if os.path.exists(lib):
        import subprocess
        out = subprocess.run(["ldd", lib], capture_output=True, text=True).stdout
        assert "oracle" not in out


def test_solve_cli_usage(sg):
    import subprocess
    out = subprocess.run([sg.SOLVE_PATH], capture_output=True, text=True)
    assert "usage: ./solve filename.mtx" in out.stdout
