// Device data model of the factorisation: block layout in HBM and the task graph the
// persistent executor consumes.  Built on the host by compile_tasks() from the reference's
// flat operation list (struct operation, operation.h:37-52).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../bigalloc.h"

namespace soglu {

// ---- block layout in HBM ---------------------------------------------------------------
// A block is 64x64 FP64 stored row-major with a leading dimension of 68 doubles
// (34 816 B).  The 4 pad doubles shift consecutive rows by 8 shared-memory banks, which
// makes BOTH m8n8k4 DMMA fragment patterns (8 rows x 4 cols for the left operand, 4 rows
// x 8 cols for the right one) hit every bank exactly twice = the 2-wavefront minimum of a
// 256-byte warp load, after ONE contiguous cp.async.bulk of the block into shared memory.
constexpr int BLK = 64;
constexpr int BLK_LD = 68;
constexpr int BLK_ELEMS = BLK * BLK_LD;                  // 4352 doubles
constexpr int BLK_BYTES = BLK_ELEMS * (int)sizeof(double);  // 34816

// references to blocks / tasks on other GPUs: owner in the top 3 bits, local index in the low 29
constexpr int REF_SHIFT = 29;
constexpr int32_t REF_MASK = (1 << REF_SHIFT) - 1;
constexpr int MAX_GPUS = 8;
inline int32_t make_ref(int owner, int32_t local) { return (int32_t)(((uint32_t)owner << REF_SHIFT) | (uint32_t)local); }
// Task references in successor lists name a task GROUP (a task, or the 2 / 4 consecutive row slices of a
// split GEMM task, which share the dependency counter of their first slice): owner | log2(group size) | local
// index of the first slice.
constexpr int TASK_SPLIT_SHIFT = 26;                       // 2 bits: log2(slices)
constexpr int32_t TASK_LOCAL_MASK = (1 << TASK_SPLIT_SHIFT) - 1;
inline int32_t make_task_ref(int owner, int log2_slices, int32_t local) {
    return make_ref(owner, local) | (log2_slices << TASK_SPLIT_SHIFT);
}

enum TaskType : int32_t {
    T_GEMM = 0,      // out = init +/- sum_p A_p * B_p   (mul / mulneg / mult chains, optional fused sub)
    T_SUB = 1,       // out = S2 - S1                   (missing source = zero block)
    T_LU = 2,        // (out, out2) = LU(src), no pivoting, |u_ii| < 1e-9 clamped
    T_LLT = 3,       // out = chol(src), pivot < 1e-20 clamped
    T_LOWERINV = 4,  // out = src^-1, src lower triangular
    T_UPPERINV = 5,  // out = src^-1, src upper triangular
    T_EXIT = 99
};
enum TaskFlags : int32_t {
    TF_NEGATE = 1,   // GEMM: subtract the product sum
    TF_TRANSB = 2,   // GEMM: use B^T (mult)
    TF_INIT = 4,     // GEMM: start from block `init` instead of zero (fused sub)
    TF_LINV = 8,     // LU/LLT: also produce out3 = L^-1 (fused lowerInv)
    TF_UINV = 16,    // LU: also produce out4 = U^-1 (fused upperInv)
    TF_TRI_OUT = 32, // LU/LLT: every result block sits in a pool slot that no other block ever occupies, so the half of the triangle
                     // that is identically zero still holds the zeros of the pool's initial memset and need not be written
    // GEMM row split: the task computes rows [16*row0, 16*row0 + 16*nrows) of the target block
    TF_ROW0_SHIFT = 8,    // 2 bits: first row / 16
    TF_NROWS_SHIFT = 12   // 3 bits: rows / 16 (4 = whole block, 2 = half, 1 = quarter)
};

struct Pair { int32_t a, b; };   // pool slots; non-GEMM: a = src (or S2), b = S1 / unused

struct Task {          // 64 bytes = one 64-byte line: the scheduler needs ONE load per task
    int32_t type;
    int32_t flags;
    int32_t n_pairs;     // GEMM: number of (A,B) pairs; others: 1
    int32_t pair_begin;  // index into the pair array
    int32_t out;         // pool slot of the result
    int32_t out2;        // LU: slot of U
    int32_t init;        // GEMM + TF_INIT: slot of the initial value; LU + TF_LINV: slot of L^-1 (out3)
    int32_t succ_begin, succ_end;  // successor task ids (CSR)
    int32_t n_deps;      // initial dependency counter (of the group; only the leader's counter is used at run time)
    int32_t level;       // ASAP level (0 = ready at start)
    int32_t out4;        // LU + TF_UINV: slot of U^-1
    Pair first[2];       // copy of the first two operand pairs (saves a dependent fetch)
};

// row slices of one split GEMM task are consecutive; the first one (row0 == 0) leads the group
inline int task_group_size(const Task& t) {
    const int rows16 = (t.flags >> TF_NROWS_SHIFT) & 7;
    return (t.type == T_GEMM && rows16 > 0) ? 4 / rows16 : 1;
}
inline bool task_is_leader(const Task& t) { return t.type != T_GEMM || ((t.flags >> TF_ROW0_SHIFT) & 3) == 0; }
inline int task_log2_slices(const Task& t) { const int g = task_group_size(t); return g == 4 ? 2 : (g == 2 ? 1 : 0); }

struct TaskGraph {
    BigVec<Task> tasks;
    BigVec<Pair> pairs;
    BigVec<int32_t> succ;          // successor GROUP leaders (task indices); the slices of one task share their list
    std::vector<int32_t> initial;       // tasks with n_deps == 0, in task order
    std::vector<int32_t> slot_of;       // block id -> pool slot (0 = zero block / none)
    std::vector<int32_t> task_of;       // block id -> producing task (-1 = input / none)
    int64_t n_slots = 1;                // slot 0 is the all-zero block
    // Segments: contiguous task ranges executed by one executor launch each.  A single segment
    // unless the block pool is smaller than the number of blocks; then slots are recycled at
    // segment boundaries (every reader of the previous occupant finished with the launch).
    std::vector<int32_t> seg_begin;     // n_segments + 1 task indices
    std::vector<int32_t> seg_init;      // n_segments + 1 offsets into `initial`
    BigVec<int32_t> succ_enc;           // succ as task references (single-GPU upload)
    std::vector<uint8_t> recycled;      // block id -> its slot is reused later (contents do not survive)
    // multi-GPU (n_owners > 1): slots are numbered per owner and every block / zero-block reference in
    // tasks and pairs carries its owner (make_ref); slot_of[] stays the owner-local slot
    int n_owners = 1;
    std::vector<int8_t> owner_of;       // block id (incl. mirror ids appended after the caller's ids) -> owner
    std::vector<int8_t> task_owner;     // task -> GPU that runs it
    std::vector<int64_t> slots_per_owner;
    std::vector<int64_t> mirrors_per_owner;
    int32_t n_levels = 0;
    double flops = 0;                   // dense-block convention, SURVEY.md 8(d)
    int64_t n_gemm_pairs = 0;
    int64_t fused_subs = 0;
    int64_t fused_invs = 0;
    int64_t aliased_invs = 0;
    int64_t split_tasks = 0;   // GEMM tasks that were row-split
};

struct CompileOptions {
    bool fuse_sub = true;   // fold `sub` into the producing mul chain when it is the only reader
    bool fuse_inv = true;   // fold the first lowerInv / upperInv of an lu's factors into the lu task
    int split_narrow = 1;   // split GEMM tasks of narrow dependency levels by output rows (latency-bound phases)
    int n_sms = 148;        // width against which a level counts as narrow
    double split_slack_us = 0;   // > 0: GEMM tasks with less estimated slack than this are split 4-way in wide levels too
    double order_alpha = 0.5;    // static order key = alpha * latest start + (1 - alpha) * earliest start
    bool static_order = true;    // tasks sorted by latest start time: the executor claims them in task order (see Compiler::static_order)
    int64_t max_slots = 0;  // block-pool capacity in slots PER GPU (0 = unlimited: no recycling)
    // multi-GPU: owner GPU of every block id (nullptr: everything on GPU 0).  Blocks a GPU reads at
    // least `mirror_min` times from a peer get a local mirror filled by a fetch task.
    const int8_t* owner_of_id = nullptr;
    int n_owners = 1;
    int mirror_min = 1;
};

// ---- multi-GPU: 2D block-cyclic owner-computes sharding -------------------------------------------
// Block (brow, bcol) lives on GPU ((brow/nb) mod pr) * pc + ((bcol/nb) mod pc); a task runs where its
// result lives and pulls remote operands over NVLink (directly, or once into a local mirror).
struct DistLayout {     // one GPU's share of an owner-compiled TaskGraph
    int rank = 0, world = 1;
    std::vector<int8_t> task_owner;      // global task -> owner
    std::vector<int32_t> task_local;     // global task -> index in the owner's task array
    std::vector<int64_t> tasks_per_rank, slots_per_rank;
    BigVec<Task> tasks;                  // successor ids rewritten to make_ref(owner, local task)
    BigVec<Pair> pairs;
    BigVec<int32_t> succ;
    std::vector<int32_t> initial;        // local task ids, grouped by segment
    std::vector<int32_t> seg_begin, seg_init;            // as in TaskGraph, for this GPU's tasks
    std::vector<std::vector<int32_t>> seg_begin_all;     // [owner][segment] first local task (part of the layout hash the ranks compare)
    int64_t remote_edges = 0, remote_operands = 0, mirrored = 0;
};
std::string localize_tasks(const TaskGraph& G, int rank, DistLayout& out);

// Returns "" on success, otherwise the violated invariant.
std::string compile_tasks(int64_t n_block_ids, int64_t n_input, const int32_t* input_ids, int64_t n_ops,
                          const int32_t* src, const int32_t* src2, const uint8_t* op, const int32_t* result,
                          const int32_t* result2, const std::vector<int32_t>& keep_ids, const CompileOptions& opt,
                          TaskGraph& out);

}  // namespace soglu
