"""Import shim: the package directory is named `sparse-operator-graph-lu_b200` (not a valid
Python identifier), so `import soglu_b200` loads it through importlib and re-exports it."""
import importlib
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
_pkg = importlib.import_module("sparse-operator-graph-lu_b200")
globals().update({k: getattr(_pkg, k) for k in dir(_pkg) if not k.startswith("__")})
