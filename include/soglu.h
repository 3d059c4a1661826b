/* soglu-b200 -- C ABI of the B200-native numeric hot path of sparse-operator-graph-LU.
 *
 * The reference (hotlei/sparse-operator-graph-LU) has no plugin/FFI interface; its seams
 * for this path are three C++ entry points that communicate through global statics:
 *
 *   BlockPlanner::calculate()                  BlockPlanner.h:36, BlockPlanner.cpp:376-651
 *       consumes data::graph / data::blockstorage / data::laststage  (data.h:44-52)
 *   BlockPlanner::solve(bl, bu, b, n)          BlockPlanner.h:38, BlockPlanner.cpp:834-862
 *       consumes the L/U quadtrees + data::blockstorage, overwrites b, writes data::x
 *   SOGLU::solveLU(dim, nnz, sym, i, j, v, b)  solver.h:24, solver.cpp:121-184
 *
 * This header replaces the first two with explicit-state calls (section A) and keeps the
 * third as a drop-in (section C).  All pointers are HOST pointers to plain arrays, borrowed
 * for the duration of the call; the library owns every byte of device memory.  Every call
 * returns 0 on success and a non-zero soglu_status otherwise, with soglu_last_error()
 * giving the text.  There is no CPU fallback: without a usable CUDA device soglu_create
 * fails with SOGLU_ERR_NO_DEVICE.
 *
 * Block = dense 64x64 FP64, row-major, 32768 bytes (the reference's 64x72 layout with its
 * sub-block bitmaps, const.h:19-34, is a CPU-only device; hosts pack/unpack at this ABI).
 * Block id 0 means "none"; ids run 1..n_block_ids-1 exactly like data::blockstorage.
 */
#ifndef SOGLU_H
#define SOGLU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SOGLU_ABI_VERSION 1

typedef enum soglu_status {
    SOGLU_OK = 0,
    SOGLU_ERR_ARG = 1,          /* bad argument / call order */
    SOGLU_ERR_NO_DEVICE = 2,    /* no CUDA device / driver */
    SOGLU_ERR_CUDA = 3,         /* a CUDA call failed; text in soglu_last_error() */
    SOGLU_ERR_OOM = 4,          /* device block pool does not fit */
    SOGLU_ERR_GRAPH = 5,        /* operation list violates the invariants of SURVEY.md App. E */
    SOGLU_ERR_IO = 6,           /* file could not be read */
    SOGLU_ERR_PLAN = 7          /* host planner rejected the matrix */
} soglu_status;

/* operation codes == enum blockOp of the reference (operation.h:20) */
enum {
    SOGLU_OP_INV = 0, SOGLU_OP_LU = 1, SOGLU_OP_LOWERINV = 2, SOGLU_OP_UPPERINV = 3, SOGLU_OP_SUB = 4,
    SOGLU_OP_ADD = 5, SOGLU_OP_NEG = 6, SOGLU_OP_COPY = 7, SOGLU_OP_MUL = 8, SOGLU_OP_MULNEG = 9,
    SOGLU_OP_LLT = 10, SOGLU_OP_MULT = 11, SOGLU_OP_NOOP = 12
};

typedef struct soglu_ctx soglu_ctx;

typedef struct soglu_stats {
    double seconds;            /* device time of the call (CUDA events on the launch stream) */
    double flops;              /* algorithmic FLOPs, dense-block convention */
    double bytes;              /* algorithmic bytes (solve: factor blocks read once + vectors) */
    int64_t kernel_launches;   /* kernels of this library launched by the call */
    int64_t tasks;             /* executor tasks run (factor) / block rows processed (solve) */
    int64_t pool_blocks;       /* device block-pool slots in use (peak) */
    double h2d_bytes, d2h_bytes;
} soglu_stats;

const char* soglu_last_error(void);
int soglu_abi_version(void);

/* ---------------- A. device hot path ------------------------------------------------- */

/* n_gpus = 1: one context on one GPU (device_ids may be NULL: device 0).
 * n_gpus = 2..8: ONE process shards the factorisation and the solve over devices device_ids[0..n_gpus-1] (NULL: 0..n_gpus-1)
 * by 2D block-cyclic block ownership (the scheme of section A2) -- peer access between the GPUs is required, the
 * operation list is held and compiled once, and every call below works on the group as on a single GPU.  This is what
 * SOGLU::solveLU / ./solve use when the environment variable SOGLU_GPUS is set. */
int soglu_create(soglu_ctx** out, int n_gpus, const int* device_ids);
void soglu_destroy(soglu_ctx* ctx);

/* replaces data::blockstorage + iniBlockStorage (BlockPlanner.cpp:1470-1544):
 * n_block_ids = data::storageCount; the n_input blocks are the ones the planner filled. */
int soglu_set_blocks(soglu_ctx* ctx, int64_t n_block_ids, int64_t n_input, const int32_t* input_ids,
                     const double* input_dense_64x64_rowmajor);
/* The same with the input blocks given as a duplicate-free entry list instead of dense arrays (the
 * reference scatters the COO entries into its blocks one by one, BlockPlanner.cpp:1498-1519; a stencil
 * block holds a few hundred non-zeros of 4096, so this cuts the host->device traffic ~30x): entry k sets
 * element (entry_pos[k] / 64, entry_pos[k] % 64) of block input_ids[entry_input[k]] to vals[k]; every
 * other element of an input block is zero.  The scatter runs on the device. */
int soglu_set_blocks_sparse(soglu_ctx* ctx, int64_t n_block_ids, int64_t n_input, const int32_t* input_ids,
                            int64_t n_entries, const int32_t* entry_input, const int32_t* entry_pos, const double* vals);

/* replaces data::graph (operation.h:37-52): parallel arrays, one entry per operation, in
 * the reference's scheduled order.  stage may be NULL (dependencies are derived from
 * src/result only); block_row/block_col (per block id, may be NULL) drive multi-GPU
 * ownership. */
int soglu_set_graph(soglu_ctx* ctx, int64_t n_ops, const int32_t* src, const int32_t* src2, const uint8_t* op,
                    const int32_t* result, const int32_t* result2, const int32_t* stage,
                    const int32_t* block_row, const int32_t* block_col);

/* replaces the L/U quadtrees handed to BlockPlanner::solve (matrix.h:22-44): the factor
 * blocks with their block coordinates.  symmetric != 0: U is L^T (nU may be 0). */
int soglu_set_factors(soglu_ctx* ctx, int64_t nL, const int32_t* L_ids, const int32_t* L_brow, const int32_t* L_bcol,
                      int64_t nU, const int32_t* U_ids, const int32_t* U_brow, const int32_t* U_bcol,
                      int32_t n_block_rows, int symmetric);

/* replaces BlockPlanner::calculate() */
int soglu_factor(soglu_ctx* ctx, soglu_stats* out);

/* The reference's numeric sanity signal: number of diagonal blocks of the last soglu_factor whose U * U^-1 fails
 * MatrixStdDouble::inv_check_diag (diagonal within 1 +- 1e-3; MatrixStdDouble.cpp:2871-2937, printed as
 * " upper out of tolerance" at BlockPlanner.cpp:575-577).  0 for a healthy factorisation; in effect it counts NaN / Inf pivots. */
int64_t soglu_diag_warnings(soglu_ctx* ctx);

/* replaces BlockPlanner::solve(): b_ext = permuted rhs padded with 1.0 to n_block_rows*64
 * (NOT overwritten, unlike the reference); x_ext receives n_block_rows*64 values.
 * One process per GPU (A2): the solve is sharded like the factorisation and therefore COLLECTIVE -- every rank calls it
 * with the same b_ext; x_ext is filled on rank 0 only (may be NULL elsewhere); put a barrier across the ranks between
 * two collective calls. */
int soglu_solve(soglu_ctx* ctx, const double* b_ext, double* x_ext, soglu_stats* out);

/* Iterative refinement on the device (the reference has none): soglu_set_matrix takes the
 * CSR of the permuted system padded with the identity to n_block_rows*64 rows; soglu_solve_refined does
 * solve + `steps` x (r = b - A x in FP64, solve, x += d) and returns the refined x_ext. */
int soglu_set_matrix(soglu_ctx* ctx, int64_t n_ext, int64_t nnz, const int64_t* row_ptr, const int32_t* col, const double* val);
int soglu_solve_refined(soglu_ctx* ctx, const double* b_ext, double* x_ext, int steps, soglu_stats* out);

/* parity helpers: read back one block (dense 64x64) after soglu_factor; error if the block
 * was recycled (only inputs of later ops, L and U are guaranteed to survive). */
int soglu_get_block(soglu_ctx* ctx, int32_t id, double* out_64x64);
/* Tuning / debug knobs, to be set before the first soglu_factor.  None changes what is computed.
 *   "exec_mode" = 1   run the op list one dependency level at a time with plain per-level launches instead of the
 *                     persistent executor (cross-check path; the same kernels' math); single GPU only
 *   "max_slots"       cap of the block pool in blocks (forces slot recycling / several executor launches)
 *   "split" "split_slack" "fuse_sub" "fuse_inv"   compiler transformations (row slices of GEMM tasks in narrow levels /
 *                     within split_slack us of the critical path, default 100; folding of sub / inverse operations)
 *   "static_order" "order_alpha"   the executor claims tasks in task order; 1 (default) = the compiler sorts them by
 *                     order_alpha % latest start + (100 - order_alpha) % earliest start time under its cost model
 *                     (default 50), 0 = they stay in the order of the operation list
 *   "dist_nb" "mirror_min"   multi-GPU: side of the ownership squares in blocks (0 = automatic: 4 for
 *                     work-bound runs, 16 when the dependency chain bounds the run), reads that justify a local mirror
 *   "grid"            number of CTAs of the executor (0 = one per SM)
 *   "watchdog_ms"     a kernel whose waiters see no progress for this long aborts; the call returns SOGLU_ERR_CUDA with
 *                     the task / block row that never arrived (default 60000, 0 = off); may be set at any time
 *   "trace"           record per-task timestamps (tools/trace_analyze.py, bench.py --trace) */
int soglu_set_option(soglu_ctx* ctx, const char* key, int64_t value);

/* ---------------- A2. multi-GPU: one process per GPU, 2D block-cyclic block ownership -------
 * Block (brow, bcol) lives on GPU ((brow / nb) mod grid_rows) * grid_cols + ((bcol / nb) mod grid_cols), nb = option dist_nb; an
 * operation runs where its result lives and pulls remote operands over NVLink (peer memory
 * mapped through CUDA IPC); dependency counters of peers are counted down with
 * system-scope reductions.  The reference has no multi-device path; this extends
 * BlockPlanner::calculate (BlockPlanner.cpp:376-651).  Call order on EVERY rank:
 *   soglu_create_dist -> soglu_set_blocks / soglu_set_graph (with block_row / block_col) /
 *   soglu_set_factors (the same full problem on every rank) -> soglu_dist_export ->
 *   [all-gather the blobs] -> soglu_dist_import -> per factorisation: soglu_dist_reset ->
 *   [barrier] -> soglu_factor -> [barrier] -> soglu_solve on EVERY rank (x on rank 0) -> [barrier]. */
int soglu_create_dist(soglu_ctx** out, int device, int rank, int world, int grid_rows, int grid_cols);
int64_t soglu_dist_blob_bytes(void);
int soglu_dist_export(soglu_ctx* ctx, void* blob);
int soglu_dist_import(soglu_ctx* ctx, const void* all_blobs_in_rank_order);
int soglu_dist_reset(soglu_ctx* ctx);
/* When one GPU's share does not fit its HBM the factorisation runs as several launches ("segments") with
 * pool slots recycled in between: for seg in 0..soglu_dist_segments-1: soglu_dist_set_segment(seg) ->
 * soglu_factor -> [barrier].  One segment: soglu_factor alone. */
int soglu_dist_segments(soglu_ctx* ctx);
int soglu_dist_set_segment(soglu_ctx* ctx, int segment);
/* out5: tasks run here (incl. fetch tasks), pool slots (owned + mirrors), dependency edges to other GPUs,
 * operand block reads from other GPUs, remote blocks mirrored locally */
int soglu_dist_info(soglu_ctx* ctx, int64_t* out5);
/* Collective calls in this mode: soglu_factor (per segment), soglu_solve / soglu_solve_refined (at most one refinement
 * step; the residual is formed on rank 0, which alone needs soglu_set_matrix).  The ranks must load the same problem
 * with the same options: soglu_dist_import rejects peers whose sharded layout (tasks, pool slots, segment boundaries per
 * GPU) differs -- the pool capacity is derived from the device's TOTAL memory for that reason, not from what is free. */

/* ---------------- B. host front-end (bit-exact integer planning) ------------------------ */

typedef struct soglu_problem soglu_problem;

/* read <path> (+ <base>_b.mtx), mirror symmetric entries, GPS-reorder, plan (solver.cpp:121-163, 50-100) */
int soglu_problem_from_mtx(const char* path, soglu_problem** out);
int soglu_problem_from_coo(int32_t dim, int64_t nnz, int symmetric, const int32_t* index_i, const int32_t* index_j,
                           const double* vals, const double* b, soglu_problem** out);
void soglu_problem_free(soglu_problem* p);
/* what: "dim" "nnz" "n_ext" "block_rows" "storage" "n_ops" "n_input" "n_L" "n_U" "coarse_ops"
 *       "coarse_storage" "symmetric" "gps_levels" "gps_width" "gps_start" ... ; -1 if unknown */
int64_t soglu_problem_size(const soglu_problem* p, const char* what);
/* what: "perm_new2old" "perm_old2new" [dim]; "ops" "coarse_ops" [n x 8: op src src2 result
 *       result2 stage group seq]; "stage" "laststage" "block_row" "block_col" [storage];
 *       "inputs" "L" "U" [n x 3: id brow bcol] */
int soglu_problem_get_i32(const soglu_problem* p, const char* what, int32_t* out);
/* what: "b" [dim]; "b_perm" [n_ext]; "input_vals" [n_input x 4096]; "flops" [1] */
int soglu_problem_get_f64(const soglu_problem* p, const char* what, double* out);
const char* soglu_problem_log(const soglu_problem* p);

/* upload a planned problem (set_blocks + set_graph + set_factors) */
int soglu_load_problem(soglu_ctx* ctx, const soglu_problem* p);
/* permuted padded solve + un-permutation: x[dim] in the ORIGINAL ordering (GPSOrder.cpp:41-53).
 * b may be NULL (use the problem's own rhs).  refine > 0 adds that many iterative-refinement
 * steps on the device (residual in FP64 on the original matrix). */
int soglu_solve_problem(soglu_ctx* ctx, const soglu_problem* p, const double* b, double* x, int refine, soglu_stats* out);

/* ---------------- C. drop-in for SOGLU::solveLU (solver.h:20-24) --------------------------- */

/* Same arguments as SOGLU::solveLU; returns a malloc'ed x[dim] the caller frees with
 * soglu_free (the reference returns arena memory released by memutil::clear()). NULL on error. */
double* soglu_solveLU(int dim, int valcount, int symmetric, const int* index_i, const int* index_j,
                      const double* vals, const double* b);
void soglu_free(void* p);

/* Host threads used by the front-end (ordering is serial; planner, task compiler and the array conversions
 * are OpenMP loops, results independent of the count).  n <= 0 restores the OpenMP default.  The reference
 * fixes its team at min(16, cores) (BlockPlanner.cpp:376-398); launchers such as torchrun export
 * OMP_NUM_THREADS=1, which this call overrides.  Returns the thread count in effect. */
int soglu_set_host_threads(int n);

/* synthetic generators used by bench/tests; kind: "lap2d" "lap3d" "nine2d" */
int soglu_write_stencil_mtx(const char* kind, int nx, int ny, int nz, int symmetric, const char* path);

#ifdef __cplusplus
}
#endif
#endif /* SOGLU_H */
