"""soglu-b200: ctypes binding of libsoglu_b200.so (the C ABI of include/soglu.h).

This module is test/bench glue only -- the product is the shared library and the ./solve
CLI.  It contains no numerics: every number comes out of the CUDA kernels behind the ABI,
and loading fails loudly when the library has not been built (no Python/CPU fallback).
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsoglu_b200.so")
SOLVE_PATH = os.path.join(_HERE, "solve")

# every symbol include/soglu.h declares (checked by tests/test_abi.py)
ABI_SYMBOLS = [
    "soglu_last_error", "soglu_abi_version", "soglu_create", "soglu_destroy", "soglu_set_blocks", "soglu_set_graph",
    "soglu_set_factors", "soglu_factor", "soglu_solve", "soglu_get_block", "soglu_set_option", "soglu_problem_from_mtx",
    "soglu_problem_from_coo", "soglu_problem_free", "soglu_problem_size", "soglu_problem_get_i32", "soglu_problem_get_f64",
    "soglu_problem_log", "soglu_load_problem", "soglu_solve_problem", "soglu_solveLU", "soglu_free", "soglu_write_stencil_mtx",
    "soglu_create_dist", "soglu_dist_blob_bytes", "soglu_dist_export", "soglu_dist_import", "soglu_dist_reset", "soglu_dist_info",
    "soglu_dist_segments", "soglu_dist_set_segment", "soglu_set_matrix", "soglu_solve_refined", "soglu_set_host_threads", "soglu_set_blocks_sparse",
    "soglu_diag_warnings",
]

OP_NAMES = {1: "lu", 2: "lowerInv", 3: "upperInv", 4: "sub", 8: "mul", 9: "mulneg", 10: "llt", 11: "mult"}


class SogluError(RuntimeError):
    pass


class Stats(ctypes.Structure):
    _fields_ = [("seconds", ctypes.c_double), ("flops", ctypes.c_double), ("bytes", ctypes.c_double),
                ("kernel_launches", ctypes.c_int64), ("tasks", ctypes.c_int64), ("pool_blocks", ctypes.c_int64),
                ("h2d_bytes", ctypes.c_double), ("d2h_bytes", ctypes.c_double)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


_lib = None


def lib():
    """Load libsoglu_b200.so (built in-tree by `make` / __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SogluError("libsoglu_b200.so is not built (%s); run __graft_entry__.build() -- there is no fallback path" % LIB_PATH)
    L = ctypes.CDLL(LIB_PATH)
    vp, i32, i64, cp = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_char_p
    L.soglu_last_error.restype = cp
    L.soglu_problem_log.restype = cp
    L.soglu_problem_log.argtypes = [vp]
    L.soglu_problem_size.restype = i64
    L.soglu_problem_size.argtypes = [vp, cp]
    L.soglu_problem_from_mtx.argtypes = [cp, ctypes.POINTER(vp)]
    L.soglu_problem_from_coo.argtypes = [i32, i64, ctypes.c_int, vp, vp, vp, vp, ctypes.POINTER(vp)]
    L.soglu_problem_free.argtypes = [vp]
    L.soglu_problem_get_i32.argtypes = [vp, cp, vp]
    L.soglu_problem_get_f64.argtypes = [vp, cp, vp]
    L.soglu_create.argtypes = [ctypes.POINTER(vp), ctypes.c_int, vp]
    L.soglu_destroy.argtypes = [vp]
    L.soglu_destroy.restype = None
    L.soglu_set_option.argtypes = [vp, cp, i64]
    L.soglu_set_blocks.argtypes = [vp, i64, i64, vp, vp]
    L.soglu_set_blocks_sparse.argtypes = [vp, i64, i64, vp, i64, vp, vp, vp]
    L.soglu_set_graph.argtypes = [vp, i64, vp, vp, vp, vp, vp, vp, vp, vp]
    L.soglu_set_factors.argtypes = [vp, i64, vp, vp, vp, i64, vp, vp, vp, i32, ctypes.c_int]
    L.soglu_factor.argtypes = [vp, ctypes.POINTER(Stats)]
    L.soglu_solve.argtypes = [vp, vp, vp, ctypes.POINTER(Stats)]
    L.soglu_get_block.argtypes = [vp, i32, vp]
    L.soglu_load_problem.argtypes = [vp, vp]
    L.soglu_solve_problem.argtypes = [vp, vp, vp, vp, ctypes.c_int, ctypes.POINTER(Stats)]
    L.soglu_solveLU.restype = vp
    L.soglu_solveLU.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, vp, vp, vp]
    L.soglu_free.argtypes = [vp]
    L.soglu_free.restype = None
    L.soglu_write_stencil_mtx.argtypes = [cp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, cp]
    L.soglu_create_dist.argtypes = [ctypes.POINTER(vp), ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    L.soglu_dist_blob_bytes.restype = i64
    L.soglu_dist_export.argtypes = [vp, vp]
    L.soglu_dist_import.argtypes = [vp, vp]
    L.soglu_dist_reset.argtypes = [vp]
    L.soglu_dist_info.argtypes = [vp, vp]
    L.soglu_dist_segments.argtypes = [vp]
    L.soglu_dist_set_segment.argtypes = [vp, ctypes.c_int]
    L.soglu_diag_warnings.argtypes = [vp]
    L.soglu_diag_warnings.restype = i64
    _lib = L
    return L


def set_host_threads(n=0):
    """Threads of the host front-end (planner, task compiler); n <= 0 = all cores.  Overrides OMP_NUM_THREADS."""
    return int(lib().soglu_set_host_threads(int(n)))


def _check(rc):
    if rc != 0:
        raise SogluError("soglu error %d: %s" % (rc, lib().soglu_last_error().decode(errors="replace")))


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


class Problem:
    """A planned problem (host front-end output): permutation, op list, blocks, factors."""

    _I32 = {"perm_new2old": ("dim", 1), "perm_old2new": ("dim", 1), "ops": ("n_ops", 8), "coarse_ops": ("coarse_ops", 8),
            "stage": ("storage", 1), "laststage": ("storage", 1), "block_row": ("storage", 1), "block_col": ("storage", 1),
            "inputs": ("n_input", 3), "L": ("n_L", 3), "U": ("n_U", 3), "perm_i": ("nnz_expanded", 1), "perm_j": ("nnz_expanded", 1),
            "entry_block": ("n_entries", 1), "entry_pos": ("n_entries", 1)}
    _F64 = {"b": ("dim", 1), "b_perm": ("n_ext", 1), "input_vals": ("n_input", 4096), "entry_val": ("n_entries", 1), "flops": (None, 1),
            "t_reorder": (None, 1), "t_plan": (None, 1)}

    def __init__(self, handle):
        self.h = handle

    @classmethod
    def from_mtx(cls, path):
        h = ctypes.c_void_p()
        _check(lib().soglu_problem_from_mtx(str(path).encode(), ctypes.byref(h)))
        return cls(h)

    @classmethod
    def from_coo(cls, dim, i, j, v, b=None, symmetric=False):
        i = np.ascontiguousarray(i, dtype=np.int32)
        j = np.ascontiguousarray(j, dtype=np.int32)
        v = np.ascontiguousarray(v, dtype=np.float64)
        b = None if b is None else np.ascontiguousarray(b, dtype=np.float64)
        h = ctypes.c_void_p()
        _check(lib().soglu_problem_from_coo(dim, len(v), int(symmetric), _ptr(i), _ptr(j), _ptr(v), _ptr(b), ctypes.byref(h)))
        return cls(h)

    def size(self, what):
        return int(lib().soglu_problem_size(self.h, what.encode()))

    def i32(self, what):
        key, cols = self._I32[what]
        n = self.size(key)
        a = np.empty((n, cols) if cols > 1 else (n,), dtype=np.int32)
        _check(lib().soglu_problem_get_i32(self.h, what.encode(), _ptr(a)))
        return a

    def f64(self, what):
        key, cols = self._F64[what]
        n = 1 if key is None else self.size(key)
        a = np.empty((n, cols) if cols > 1 else (n,), dtype=np.float64)
        _check(lib().soglu_problem_get_f64(self.h, what.encode(), _ptr(a)))
        return a

    @property
    def log(self):
        return lib().soglu_problem_log(self.h).decode(errors="replace")

    def close(self):
        if self.h:
            lib().soglu_problem_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Context:
    """One GPU context: block pool + compiled task graph + factor/solve."""

    def __init__(self, device=0, rank=0, world=1, grid=None, n_gpus=1):
        """world > 1: this process drives one GPU of a 2D block-cyclic process grid (rows, cols).
        n_gpus > 1: ONE process shards the factorisation over GPUs 0..n_gpus-1 (soglu_create's in-process group)."""
        self.h = ctypes.c_void_p()
        self.rank, self.world = rank, world
        if n_gpus > 1:
            _check(lib().soglu_create(ctypes.byref(self.h), int(n_gpus), None))
        elif world == 1:
            dev = (ctypes.c_int * 1)(device)
            _check(lib().soglu_create(ctypes.byref(self.h), 1, dev))
        else:
            pr, pc = grid if grid else default_grid(world)
            _check(lib().soglu_create_dist(ctypes.byref(self.h), device, rank, world, pr, pc))

    # ---- multi-GPU plumbing: the caller moves the blobs between ranks (torch.distributed) ----
    def dist_export(self):
        n = int(lib().soglu_dist_blob_bytes())
        blob = np.zeros(n, dtype=np.uint8)
        _check(lib().soglu_dist_export(self.h, _ptr(blob)))
        return blob

    def dist_import(self, all_blobs):
        all_blobs = np.ascontiguousarray(all_blobs, dtype=np.uint8)
        _check(lib().soglu_dist_import(self.h, _ptr(all_blobs)))

    def dist_reset(self):
        _check(lib().soglu_dist_reset(self.h))

    def factor_dist(self, barrier):
        """One sharded factorisation: reset, then every segment on all ranks with `barrier()` in between.
        Returns the stats of the last segment with `seconds` summed over segments."""
        self.dist_reset()
        barrier()
        nseg = int(lib().soglu_dist_segments(self.h))
        total, st = 0.0, None
        for sg in range(nseg):
            _check(lib().soglu_dist_set_segment(self.h, sg))
            st = self.factor()
            total += st["seconds"]
            barrier()
        st["seconds"] = total
        st["segments"] = nseg
        return st

    def dist_info(self):
        out = np.zeros(5, dtype=np.int64)
        _check(lib().soglu_dist_info(self.h, _ptr(out)))
        return dict(zip(("tasks", "slots", "remote_edges", "remote_operands", "mirrored"), out.tolist()))

    def set_option(self, key, value):
        _check(lib().soglu_set_option(self.h, key.encode(), int(value)))

    def load(self, problem):
        _check(lib().soglu_load_problem(self.h, problem.h))

    def set_blocks(self, n_ids, input_ids, dense):
        input_ids = np.ascontiguousarray(input_ids, dtype=np.int32)
        assert dense.dtype == np.float64 and dense.flags["C_CONTIGUOUS"]
        _check(lib().soglu_set_blocks(self.h, int(n_ids), len(input_ids), _ptr(input_ids), _ptr(dense)))

    def set_blocks_sparse(self, n_ids, input_ids, entry_input, entry_pos, vals):
        """Input blocks as an entry list: vals[k] goes to element entry_pos[k] (row*64+col) of block
        input_ids[entry_input[k]]; arrays must be contiguous int32 / float64 (pinned host memory is fine)."""
        input_ids = np.ascontiguousarray(input_ids, dtype=np.int32)
        for a, dt in ((entry_input, np.int32), (entry_pos, np.int32), (vals, np.float64)):
            assert a.dtype == dt and a.flags["C_CONTIGUOUS"]
        _check(lib().soglu_set_blocks_sparse(self.h, int(n_ids), len(input_ids), _ptr(input_ids), len(vals), _ptr(entry_input), _ptr(entry_pos), _ptr(vals)))

    def set_graph(self, ops, stage=None, brow=None, bcol=None):
        """ops: dict of int32 arrays src, src2, result, result2 and uint8 op."""
        c = lambda a, dt: np.ascontiguousarray(a, dtype=dt)
        src, src2, res, res2 = (c(ops[k], np.int32) for k in ("src", "src2", "result", "result2"))
        op = c(ops["op"], np.uint8)
        stage = None if stage is None else c(stage, np.int32)
        _check(lib().soglu_set_graph(self.h, len(op), _ptr(src), _ptr(src2), _ptr(op), _ptr(res), _ptr(res2), _ptr(stage),
                                     _ptr(brow), _ptr(bcol)))

    def set_factors(self, L, U, n_block_rows, symmetric=False):
        c = lambda a: np.ascontiguousarray(a, dtype=np.int32)
        li, lr, lc = c(L[:, 0]), c(L[:, 1]), c(L[:, 2])
        if U is None or len(U) == 0:
            _check(lib().soglu_set_factors(self.h, len(li), _ptr(li), _ptr(lr), _ptr(lc), 0, None, None, None, n_block_rows, int(symmetric)))
        else:
            ui, ur, uc = c(U[:, 0]), c(U[:, 1]), c(U[:, 2])
            _check(lib().soglu_set_factors(self.h, len(li), _ptr(li), _ptr(lr), _ptr(lc), len(ui), _ptr(ui), _ptr(ur), _ptr(uc),
                                           n_block_rows, int(symmetric)))

    def diag_warnings(self):
        """Diagonal blocks of the last factorisation that fail the reference's inv_check_diag (0 when healthy)."""
        return int(lib().soglu_diag_warnings(self.h))

    def segments(self):
        """Executor launches per factorisation (more than one when pool slots are recycled); valid once compiled."""
        return int(lib().soglu_dist_segments(self.h))

    def factor(self):
        st = Stats()
        _check(lib().soglu_factor(self.h, ctypes.byref(st)))
        return st.as_dict()

    def solve_ext(self, b_ext):
        b_ext = np.ascontiguousarray(b_ext, dtype=np.float64)
        x = np.empty_like(b_ext)
        st = Stats()
        _check(lib().soglu_solve(self.h, _ptr(b_ext), _ptr(x), ctypes.byref(st)))
        return x, st.as_dict()

    def solve(self, problem, b=None, refine=0):
        """x in the original ordering; refine > 0 adds that many device-side iterative-refinement steps."""
        x = np.empty(problem.size("dim"), dtype=np.float64)
        b = None if b is None else np.ascontiguousarray(b, dtype=np.float64)
        st = Stats()
        _check(lib().soglu_solve_problem(self.h, problem.h, _ptr(b), _ptr(x), int(refine), ctypes.byref(st)))
        return x, st.as_dict()

    def get_block(self, block_id):
        out = np.empty((64, 64), dtype=np.float64)
        _check(lib().soglu_get_block(self.h, int(block_id), _ptr(out)))
        return out

    def close(self):
        if self.h:
            lib().soglu_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def default_grid(world):
    """process grid (rows, cols) of the 2D block-cyclic ownership: 1->1x1, 2->2x1, 4->2x2, 8->4x2 (the grids the scaling numbers were measured with)"""
    pr = 1
    while pr * pr * 2 <= world:
        pr *= 2
    if world % pr:
        pr = 1
    return pr, world // pr


def write_stencil_mtx(kind, path, nx, ny=0, nz=0, symmetric=False):
    _check(lib().soglu_write_stencil_mtx(kind.encode(), nx, ny, nz, int(symmetric), str(path).encode()))


def solve_lu(dim, i, j, v, b, symmetric=False):
    """Drop-in for SOGLU::solveLU: returns x in the original ordering."""
    i = np.ascontiguousarray(i, dtype=np.int32)
    j = np.ascontiguousarray(j, dtype=np.int32)
    v = np.ascontiguousarray(v, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    p = lib().soglu_solveLU(dim, len(v), int(symmetric), _ptr(i), _ptr(j), _ptr(v), _ptr(b))
    if not p:
        raise SogluError(lib().soglu_last_error().decode(errors="replace"))
    x = np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_double)), shape=(dim,)).copy()
    lib().soglu_free(p)
    return x
