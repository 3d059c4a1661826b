"""Prints hashes of the planner output and of compiled task graphs for one golden case (run by
test_determinism.py under different OMP_NUM_THREADS)."""
import ctypes, hashlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import soglu_b200 as sg

p = sg.Problem.from_mtx(sys.argv[1])
L = sg.lib()
L.soglu_debug_graph_hash.restype = ctypes.c_uint64
L.soglu_debug_graph_hash.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int]
out = {"ops": hashlib.sha1(p.i32("ops").tobytes()).hexdigest(), "laststage": hashlib.sha1(p.i32("laststage").tobytes()).hexdigest()}
L.soglu_debug_compile.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int]
st = (ctypes.c_int64 * 16)()
assert L.soglu_debug_compile(p.h, 1, 1, 1, 0, st, 16) == 0
slots = st[4]                      # pool slots without recycling
for cfg in [(0, 0, 1, 1, 1), (1, 0, 1, 1, 1), (1, int(0.7 * slots), 1, 1, 1), (1, 0, 2, 2, 1), (1, int(0.3 * slots), 4, 2, 2)]:
    out[str(cfg)] = L.soglu_debug_graph_hash(p.h, *cfg)
print(json.dumps(out))
