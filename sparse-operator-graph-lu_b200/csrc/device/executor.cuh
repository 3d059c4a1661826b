// Launch interface of the device kernels (implemented in executor.cu / trsv.cu).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "tasks.h"

namespace soglu {

struct ExecParams {
    // Multi-GPU: per-owner base pointers (peer-mapped through CUDA IPC); block / task references carry
    // the owner in their top 3 bits (tasks.h make_ref).  Single GPU: world = 1, index 0 only.
    double* pools[MAX_GPUS];
    int32_t* deps[MAX_GPUS];
    int32_t world, rank;
    double* pool;          // this GPU's block pool, slot s at pool + s*BLK_ELEMS
    const Task* tasks;
    const Pair* pairs;
    const int32_t* succ;   // task references: owner | log2(slices) | local id (tasks.h)
    int32_t* dep;          // live dependency counters (reset before every run)
    int32_t* head;         // next position to claim
    int32_t n_tasks;       // tasks of this launch
    int32_t task0;         // signal = 1: first task of the segment this launch executes (position s = task task0 + s)
    int32_t signal;        // 1: the product path -- tasks are claimed in task order, wait on their dependency counters and
                           //    count their successors' counters down; 0: debug executor (one launch per dependency
                           //    level, no counters): position s = task ready[s]
    const int32_t* ready;  // signal = 0: the task list of this launch
    int32_t* abort;        // watchdog word of this GPU {flag, task position, CTA, rank}; aborts[g] = the peers' (multi-GPU);
                           // abort[8] counts diagonal blocks whose U U^-1 fails the reference's inv_check_diag
    int32_t* aborts[MAX_GPUS];
    unsigned long long watchdog_ns;   // a scheduler lane that has waited this long since the launch gives up (0 = never)
    int32_t debug_drop;    // test hook for the watchdog: the completion of this task is NOT propagated (-1 = none)
    unsigned long long* trace;   // optional: 6 x u64 per task (counter seen at zero, issued, loaded = first math, computed, signalled, smid)
};

// persistent dependency-counted executor; grid = resident CTAs (1 per SM)
cudaError_t launch_executor(const ExecParams& p, int grid, cudaStream_t stream);
int executor_max_grid(int device);        // co-resident CTAs for the executor kernel
size_t executor_smem_bytes();

cudaError_t launch_diag_bench(double* pool, int iters, long long* cycles, cudaStream_t stream);

// scatter dense 64x64 blocks (row-major, ld 64) into pool slots (ld 68) and back
cudaError_t launch_pack_blocks(double* pool, const double* dense, const int32_t* slots, int64_t n, cudaStream_t stream);
cudaError_t launch_scatter_entries(double* pool, const int32_t* slots, int64_t n_blocks, const int32_t* entry_input, const int32_t* entry_pos,
                                   const double* vals, int64_t n_entries, cudaStream_t stream);
cudaError_t launch_unpack_block(const double* pool, int32_t slot, double* dense, cudaStream_t stream);

// ---- block triangular solve -----------------------------------------------------------
// Sharded run: every GPU processes the block rows whose diagonal block it owns (my_rows), reads the factor blocks of
// those rows (its own, or a peer's over NVLink) and publishes each finished 64-value segment into EVERY GPU's copy of
// y / x, so consumers always poll local memory.
struct TrsvParams {
    const double* pools[MAX_GPUS];   // factor blocks may live on peer GPUs (references as in ExecParams)
    const double* pool;
    // CSR by block row over off-diagonal factor blocks, columns ascending
    const int64_t* l_ptr; const int32_t* l_col; const int32_t* l_slot; const int32_t* l_diag; const int32_t* l_dinv;
    const int64_t* u_ptr; const int32_t* u_col; const int32_t* u_slot; const int32_t* u_diag; const int32_t* u_dinv;
    int32_t n_rows;        // block rows
    const int32_t* my_rows;   // the block rows this GPU processes, ascending (all of them on one GPU)
    int32_t n_my_rows;
    int32_t world, rank;
    const double* b;       // n_rows*64: right-hand side (local copy)
    int32_t b_polled;      // b arrives through the sentinel protocol (the refinement's residual, written by rank 0)
    double* y;             // forward result, this GPU's copy (pre-filled with the NaN sentinel)
    double* x;             // backward result (idem)
    double* y_all[MAX_GPUS];   // every GPU's copy: a finished segment is stored into all of them
    double* x_all[MAX_GPUS];
    int32_t symmetric;     // U = L^T: backward sweep reads L blocks transposed (CSC of L passed in u_*)
    int32_t* abort;        // watchdog word (as in ExecParams): a consumer that has polled for watchdog_ns raises it
    unsigned long long watchdog_ns;
};
cudaError_t launch_trsv(const TrsvParams& p, int grid, cudaStream_t stream);
int trsv_max_grid(int device);
cudaError_t launch_fill_sentinel(double* a, int64_t n, cudaStream_t stream);
// device-side barrier across the GPUs of a sharded solve (flags_all[g] = GPU g's array of world epoch slots)
cudaError_t launch_peer_epoch(int32_t* const* flags_all, int world, int rank, int32_t epoch, int32_t* abort, unsigned long long watchdog_ns, cudaStream_t stream);
// iterative refinement: r = b - A x (CSR of the permuted padded system), written into every GPU's copy of r; x += d
cudaError_t launch_residual(const int64_t* rp, const int32_t* ci, const double* v, const double* b, const double* x, double* const* r_all, int world,
                            int64_t n, cudaStream_t stream);
cudaError_t launch_axpy(double* x, const double* d, int64_t n, cudaStream_t stream);

}  // namespace soglu
