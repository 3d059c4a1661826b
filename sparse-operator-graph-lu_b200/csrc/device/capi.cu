// Section A of include/soglu.h: context, device-resident block pool, task-graph upload,
// factor and solve entry points.  All numeric work is done by the kernels in executor.cu
// and trsv.cu; there is no CPU path -- every entry point fails if CUDA is unavailable.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

#include <cuda_runtime.h>

#include "../../../include/soglu.h"
#include "executor.cuh"
#include "tasks.h"

namespace soglu {
void set_error(const std::string& s);
}

using namespace soglu;

namespace {

struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    // grow-only: a buffer that is already large enough is kept (cudaMalloc / cudaFree cost milliseconds each once tens of
    // GB and peer mappings exist -- the per-step staging buffers of a refactorisation must not pay that)
    cudaError_t reserve(size_t n) {
        if (p && bytes >= n) return cudaSuccess;
        return alloc(n);
    }
    cudaError_t alloc(size_t n) {
        release();
        if (n == 0) return cudaSuccess;
        cudaError_t e = cudaMalloc(&p, n);
        if (e == cudaSuccess) bytes = n; else p = nullptr;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; bytes = 0; }
    template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

}  // namespace

struct soglu_ctx {
    int device = 0;
    int sms = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int exec_grid = 0, trsv_grid = 0;
    // options
    int64_t opt_exec_mode = 0;     // 0 persistent DAG executor, 1 one launch per level (debug)
    int64_t opt_fuse_sub = 1;
    int64_t opt_fuse_inv = 1;
    int64_t opt_split_slack = 100; // GEMM tasks within this slack (us) of the longest chain are row-split in wide levels too (measured -5..6 % on the
                                   // latency-bound configs, profiles/r02_call1_options.md); 0 = narrow levels only
    int64_t opt_dist_nb = 0;       // multi-GPU ownership granularity (blocks); 0 = chosen in finalize(): 4 x 4 squares when the run is
                                   // work-bound (finer squares balance the moving band: 100^3 on 2 GPUs, nb 16 / 8 / 4 / 2 = 1478 / 1377 /
                                   // 1325 / 1322 ms), 16 x 16 when the dependency chain bounds it (every square boundary on the chain
                                   // costs remote hops: 64^3 on 2 GPUs, nb 1 / 4 / 16 / 64 = 204 / 164 / 156 / 165 ms)
    int64_t opt_mirror_min = 1;    // mirror a remote block locally when it is read at least this often
    int64_t opt_split = 1;
    int64_t opt_static_order = 1;  // 1: tasks sorted most-urgent-first (latest start time); 0: in the order of the operation list
    int64_t opt_grid = 0;          // override CTA count (0 = all resident)
    int64_t opt_trace = 0;         // record per-task timestamps (debug; adds overhead)
    int64_t opt_order_alpha = 50;      // static order key = alpha % latest start + (100 - alpha) % earliest start (measured: 40..70 best, 100 = pure
                                       // latest start loses 3-10 % because just-in-time tasks that slip delay the critical chain, 0 = earliest start loses 5-15 %)
    int64_t opt_debug_drop = -1;       // test hook: lose the completion signal of this task (the watchdog must catch the hang)
    int64_t opt_watchdog_ms = 60000;   // a kernel whose waiters see no progress for this long aborts with SOGLU_ERR_CUDA (0 = off)
    DevBuf trace;

    // multi-GPU (one process per GPU): process grid, localized task graph, peer mappings
    bool dist = false;
    int rank = 0, world = 1, pr = 1, pc = 1;
    DistLayout D;
    std::vector<int32_t> brow, bcol;      // per block id (ownership)
    bool peers_ready = false;
    bool ipc_peers = true;              // peer pointers come from cudaIpcOpenMemHandle (one process per GPU); false: same process
    std::vector<soglu_ctx*> members;    // in-process group: this context only forwards to its members (rank g on device g)
    soglu_ctx* leader = nullptr;        // member of a group: rank 0 of it (owns the compiled graph)
    int dist_segment = 0;               // segment the next soglu_factor call runs (multi-GPU)
    void* peer_pool[MAX_GPUS] = {}, *peer_dep[MAX_GPUS] = {}, *peer_counters[MAX_GPUS] = {};

    // host-side description (borrowed arrays are copied)
    int64_t n_ids = 0, n_input = 0;
    std::vector<int32_t> input_ids;
    DevBuf in_dense;               // staging for dense input blocks until slots are known
    DevBuf in_entry_input, in_entry_pos, in_entry_val;   // staging for a sparse entry list (soglu_set_blocks_sparse); kept for the next upload
    DevBuf in_slots;               // pool slot of every input block on this GPU (-1: another GPU's), built once
    int64_t n_entries = -1;        // >= 0: the pending inputs are an entry list
    bool inputs_pending = false;
    int64_t n_ops = 0;
    BigVec<int32_t> src, src2, result, result2;   // released once the graph is compiled
    BigVec<uint8_t> op;
    std::vector<int32_t> L_ids, L_brow, L_bcol, U_ids, U_brow, U_bcol;
    int32_t n_block_rows = 0;
    int dist_nb_used = 0;          // side of the ownership squares this context was compiled with
    int symmetric = 0;
    bool have_blocks = false, have_graph = false, have_factors = false;

    // compiled state
    bool compiled = false;
    bool factored = false;
    TaskGraph G_own;
    TaskGraph* Gp = &G_own;          // members of an in-process group (soglu_create with n_gpus > 1) share rank 0's graph
    std::vector<int32_t> level_order;   // tasks sorted by level (debug executor)
    std::vector<int64_t> level_ptr;
    DevBuf pool, tasks, pairs, succ, dep0, dep, ready, counters, counters0;
    // watchdog word {flag, task position / block row, CTA, rank}: the 64 bytes behind the claim counters (one allocation,
    // so the peers reach it through the IPC mapping of the counters)
    int32_t* abort_word() const { return counters.as<int32_t>() + counters0.bytes / 4; }
    int64_t opt_max_slots = 0;     // debug: cap the block pool (forces segments + slot recycling)
    // solve structures
    DevBuf l_ptr, l_col, l_slot, l_diag, l_dinv, u_ptr, u_col, u_slot, u_diag, u_dinv, d_b;
    // solve vectors, peer-visible in a sharded run: [y0 | x0 | y1 | x1 | r0 | r1], n_ext doubles each (two sets so that a
    // refinement step never refills a vector a slower peer may still read)
    DevBuf sv, my_rows;       // (sv: + 64 bytes of epoch slots for the device-side barrier of the sharded solve)
    int32_t n_my_rows = 0;
    int64_t diag_warnings = 0; // diagonal blocks of the last factorisation that failed inv_check_diag (NaN / Inf pivots)
    int32_t solve_epoch = 0;  // collective solves so far (the ranks call in lockstep)
    void* peer_sv[MAX_GPUS] = {};
    int64_t nL_off = 0, nU_off = 0;
    DevBuf m_rp, m_ci, m_v, d_xacc;   // CSR of the permuted padded matrix (iterative refinement)
    int64_t m_n = 0, m_nnz = 0;
    int64_t launches = 0;
    double h2d = 0, d2h = 0;
};

namespace {

#define CU(call)                                                                                  \
    do {                                                                                          \
        cudaError_t e__ = (call);                                                                 \
        if (e__ != cudaSuccess) {                                                                 \
            soglu::set_error(std::string(#call) + ": " + cudaGetErrorString(e__));                \
            return SOGLU_ERR_CUDA;                                                                \
        }                                                                                         \
    } while (0)

int fail(int code, const std::string& msg) { soglu::set_error(msg); return code; }

// after the stream has been synchronised: did a kernel's watchdog give up?  (clears the word for the next call)
int check_watchdog(soglu_ctx* c, const char* what) {
    int32_t w[12] = {};
    if (cudaMemcpy(w, c->abort_word(), sizeof w, cudaMemcpyDeviceToHost) != cudaSuccess) return fail(SOGLU_ERR_CUDA, "cannot read the watchdog word");
    if (what[0] == 'f') {           // factorisation: the inv_check_diag counter (reset for the next one)
        c->diag_warnings = w[8];
        if (w[8]) cudaMemset(c->abort_word() + 8, 0, 4);
    }
    if (w[0] == 0) return SOGLU_OK;
    cudaMemset(c->abort_word(), 0, 64);
    char m[320];
    if (w[0] == 2) snprintf(m, sizeof m, "%s aborted: a peer GPU's watchdog gave up (see its error)", what);
    else if (w[1] < 0) snprintf(m, sizeof m, "%s aborted by the watchdog after %lld ms: GPU %d never reached the collective call (every rank of a sharded run must call it)", what,
                                (long long)c->opt_watchdog_ms, -1 - w[1]);
    else snprintf(m, sizeof m, "%s aborted by the watchdog after %lld ms without progress: CTA %d was waiting for %s %d%s", what, (long long)c->opt_watchdog_ms,
                  w[2], c->factored && what[0] == 's' ? "block row" : "task (position in its segment)", w[1],
                  " (a lost dependency signal: a peer that failed, or an operation list with a missing edge)");
    return fail(SOGLU_ERR_CUDA, m);
}

template <typename T, typename A>
int upload(DevBuf& b, const std::vector<T, A>& v, soglu_ctx* c) {
    CU(b.alloc(std::max<size_t>(v.size(), 1) * sizeof(T)));
    if (!v.empty()) CU(cudaMemcpyAsync(b.p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, c->stream));
    c->h2d += (double)(v.size() * sizeof(T));
    return SOGLU_OK;
}

// block id -> block reference used by the kernels (owner in the top bits; plain slot on one GPU)
int32_t id_ref(const soglu_ctx* c, int32_t id) {
    const int32_t sl = c->Gp->slot_of[id];
    return sl > 0 ? make_ref(c->Gp->owner_of[id], sl) : 0;
}

// CSR over off-diagonal factor blocks of one triangle; lower: cols < row, upper: cols > row.
// transpose = true builds the structure of the transposed factor (CSC of L for L^T).
int build_tri(soglu_ctx* c, const std::vector<int32_t>& ids, const std::vector<int32_t>& br, const std::vector<int32_t>& bc,
              bool upper, bool transpose, DevBuf& dptr, DevBuf& dcol, DevBuf& dslot, DevBuf& ddiag, DevBuf& ddinv, int64_t& n_off,
              std::vector<int32_t>* diag_out = nullptr) {
    const int n = c->n_block_rows;
    std::vector<int64_t> ptr(n + 1, 0);
    std::vector<int32_t> diag(n, -1);
    const size_t m = ids.size();
    for (size_t k = 0; k < m; k++) {
        int r = transpose ? bc[k] : br[k], cc = transpose ? br[k] : bc[k];
        if (r < 0 || r >= n || cc < 0 || cc >= n) return fail(SOGLU_ERR_ARG, "factor block coordinate out of range");
        if (ids[k] <= 0 || ids[k] >= c->n_ids) return fail(SOGLU_ERR_ARG, "factor block id out of range");
        if (r == cc) { diag[r] = id_ref(c, ids[k]); continue; }
        if (upper ? (cc < r) : (cc > r)) return fail(SOGLU_ERR_ARG, "factor block on the wrong side of the diagonal");
        ptr[r + 1]++;
    }
    for (int r = 0; r < n; r++) {
        if (diag[r] == -1 || diag[r] == 0) return fail(SOGLU_ERR_GRAPH, "factor has no diagonal block in block row " + std::to_string(r));
        ptr[r + 1] += ptr[r];
    }
    n_off = ptr[n];
    std::vector<int32_t> col(n_off), slot(n_off);
    std::vector<int64_t> pos(ptr.begin(), ptr.end() - 1);
    for (size_t k = 0; k < m; k++) {
        int r = transpose ? bc[k] : br[k], cc = transpose ? br[k] : bc[k];
        if (r == cc) continue;
        col[pos[r]] = cc;
        slot[pos[r]] = id_ref(c, ids[k]);
        pos[r]++;
    }
    // ascending columns inside each row
    std::vector<std::pair<int32_t, int32_t>> tmp;
    for (int r = 0; r < n; r++) {
        tmp.clear();
        for (int64_t q = ptr[r]; q < ptr[r + 1]; q++) tmp.push_back({col[q], slot[q]});
        std::sort(tmp.begin(), tmp.end());
        for (int64_t q = ptr[r]; q < ptr[r + 1]; q++) { col[q] = tmp[q - ptr[r]].first; slot[q] = tmp[q - ptr[r]].second; }
    }
    for (int64_t q = 0; q < n_off; q++)
        if (slot[q] == 0) return fail(SOGLU_ERR_GRAPH, "factor block is never produced by the operation list");
    int rc;
    if ((rc = upload(dptr, ptr, c))) return rc;
    if ((rc = upload(dcol, col, c))) return rc;
    if ((rc = upload(dslot, slot, c))) return rc;
    if ((rc = upload(ddiag, diag, c))) return rc;
    // explicit inverses of the diagonal blocks, where the factorisation produced them (fused lu tasks)
    std::unordered_map<int32_t, int32_t> inv_of_ref;   // reference of a diagonal factor block -> reference of its inverse
    for (const Task& T : c->Gp->tasks)
        if (T.type == T_LU || T.type == T_LLT) {      // llt: L^-1 also serves the transposed sweep, (L^T)^-1 = (L^-1)^T
            if (T.flags & TF_LINV) inv_of_ref[T.out] = T.init;
            if (T.flags & TF_UINV) inv_of_ref[T.out2] = T.out4;
        }
    std::vector<int32_t> dinv(n, 0);
    for (int r = 0; r < n; r++) {
        auto it = inv_of_ref.find(diag[r]);
        if (it != inv_of_ref.end()) dinv[r] = it->second;
    }
    if ((rc = upload(ddinv, dinv, c))) return rc;
    if (diag_out) *diag_out = diag;
    return SOGLU_OK;
}

int pack_pending_inputs(soglu_ctx* c) {
    if (!c->inputs_pending) return SOGLU_OK;
    if (!c->in_slots.p) {
        std::vector<int32_t> slots(c->n_input);
        for (int64_t k = 0; k < c->n_input; k++) {
            const int32_t id = c->input_ids[k];
            slots[k] = (!c->dist || c->Gp->owner_of[id] == c->rank) ? c->Gp->slot_of[id] : -1;   // other GPUs' inputs are skipped
        }
        int rc = upload(c->in_slots, slots, c);
        if (rc) return rc;
    }
    if (c->n_entries >= 0) {
        CU(launch_scatter_entries(c->pool.as<double>(), c->in_slots.as<int32_t>(), c->n_input, c->in_entry_input.as<int32_t>(), c->in_entry_pos.as<int32_t>(),
                                  c->in_entry_val.as<double>(), c->n_entries, c->stream));
        c->launches += c->n_entries > 0 ? 2 : 1;
    } else {
        CU(launch_pack_blocks(c->pool.as<double>(), c->in_dense.as<double>(), c->in_slots.as<int32_t>(), c->n_input, c->stream));
        c->launches++;
    }
    CU(cudaStreamSynchronize(c->stream));
    c->in_dense.release();         // (dense staging is 30x larger than the entry list: not kept)
    c->n_entries = -1;
    c->inputs_pending = false;
    return SOGLU_OK;
}

int finalize(soglu_ctx* c) {
    if (c->compiled) return pack_pending_inputs(c);
    const bool follower = c->leader && c->leader != c;     // in-process group: rank 0 compiled the graph for everyone
    if (follower) {
        if (!c->leader->compiled) return fail(SOGLU_ERR_ARG, "internal: group leader not compiled");
        if (!c->have_blocks) return fail(SOGLU_ERR_ARG, "soglu_set_blocks must precede soglu_factor");
        c->Gp = c->leader->Gp;
    } else if (!c->have_blocks || !c->have_graph || !c->have_factors) {
        return fail(SOGLU_ERR_ARG, "soglu_set_blocks, soglu_set_graph and soglu_set_factors must precede soglu_factor");
    }
    if (!follower) {
    std::vector<int32_t> keep;
    keep.reserve(c->L_ids.size() + c->U_ids.size());
    keep.insert(keep.end(), c->L_ids.begin(), c->L_ids.end());
    keep.insert(keep.end(), c->U_ids.begin(), c->U_ids.end());
    CompileOptions co;
    co.fuse_sub = c->opt_fuse_sub != 0;
    co.fuse_inv = c->opt_fuse_inv != 0;
    co.split_narrow = (int)c->opt_split;
    co.n_sms = c->sms;
    co.split_slack_us = (double)std::max<int64_t>(0, c->opt_split_slack);
    co.static_order = c->opt_static_order != 0;
    co.order_alpha = (double)c->opt_order_alpha / 100.0;
    {
        // Pool capacity: what is free now minus the graph arrays (estimated from the op count) and a margin.
        // Sharded runs: every rank compiles the WHOLE graph and must arrive at the same slot numbers and segment
        // boundaries for every GPU, so the capacity must not depend on rank-local state (a few MB of difference in
        // free memory would shift the recycling boundaries): it is derived from the device's TOTAL memory, the same
        // number on the identical GPUs of one box, with a fixed 7 % reserve for contexts and NCCL buffers; the
        // resulting layout is hashed into the peer blob and soglu_dist_import rejects a mismatch.
        size_t free_b = 0, total_b = 0;
        CU(cudaMemGetInfo(&free_b, &total_b));
        const double graph_est = (double)c->n_ops * 40.0 + (double)c->n_block_rows * 64 * 8 * 4 + 1.5e9;
        const double budget = c->dist ? 0.93 * (double)total_b : (double)free_b;
        double cap = (budget - graph_est / (c->dist ? c->world : 1)) / (double)BLK_BYTES;
        if (cap < 16) cap = 16;
        co.max_slots = (int64_t)cap;
        if (c->opt_max_slots > 0 && c->opt_max_slots < co.max_slots) co.max_slots = c->opt_max_slots;
    }
    std::vector<int8_t> owners;
    if (c->dist) {
        // 2D block-cyclic ownership over nb x nb squares of blocks (coordinates from soglu_set_graph)
        if (c->brow.empty()) return fail(SOGLU_ERR_ARG, "multi-GPU context: soglu_set_graph needs block_row / block_col");
        int nb = (int)c->opt_dist_nb;
        if (nb <= 0) {
            // work per GPU at the measured GEMM rate against the dependency chain (one diagonal block after the other,
            // ~36 us per block row on one GPU): the same numbers on every rank
            int64_t products = 0;
            for (int64_t k = 0; k < c->n_ops; k++) products += (c->op[k] == 8 || c->op[k] == 9 || c->op[k] == 11);     // mul, mulneg, mult (soglu.h op codes)
            const double t_work = (double)products * 2.0 * BLK * BLK * BLK / 33e12 / c->world;
            const double t_chain = (double)c->n_block_rows * 36e-6;
            nb = t_work > 1.3 * t_chain ? 4 : 16;
        }
        c->dist_nb_used = nb;
        owners.assign(c->n_ids, 0);
        for (int64_t id = 1; id < c->n_ids; id++)
            if (c->brow[id] >= 0 && c->bcol[id] >= 0) owners[id] = (int8_t)(((c->brow[id] / nb) % c->pr) * c->pc + ((c->bcol[id] / nb) % c->pc));
        co.owner_of_id = owners.data();
        co.n_owners = c->world;
        co.mirror_min = (int)std::max<int64_t>(1, c->opt_mirror_min);
    }
    std::string err = compile_tasks(c->n_ids, c->n_input, c->input_ids.data(), c->n_ops, c->src.data(), c->src2.data(), c->op.data(),
                                    c->result.data(), c->result2.data(), keep, co, *c->Gp);
    if (!err.empty()) return fail(SOGLU_ERR_GRAPH, err);
    }   // !follower
    TaskGraph& G = *c->Gp;
    if (c->dist) {
        std::string err = localize_tasks(G, c->rank, c->D);
        if (!err.empty()) return fail(SOGLU_ERR_GRAPH, err);
    }
    size_t free_b = 0, total_b = 0;
    CU(cudaMemGetInfo(&free_b, &total_b));
    const BigVec<Task>& tasks_up = c->dist ? c->D.tasks : G.tasks;
    const BigVec<Pair>& pairs_up = c->dist ? c->D.pairs : G.pairs;
    const BigVec<int32_t>& succ_up = c->dist ? c->D.succ : G.succ_enc;
    const size_t pool_bytes = (size_t)G.slots_per_owner[c->dist ? c->rank : 0] * BLK_BYTES;
    const size_t aux = tasks_up.size() * (sizeof(Task) + 12) + pairs_up.size() * sizeof(Pair) + succ_up.size() * 4 + (256u << 20);
    if (pool_bytes + aux > free_b) {
        char m[256];
        snprintf(m, sizeof m, "block pool needs %.1f GB (+%.1f GB graph) but only %.1f GB of HBM are free", pool_bytes * 1e-9, aux * 1e-9, free_b * 1e-9);
        return fail(SOGLU_ERR_OOM, m);
    }
    CU(c->pool.alloc(pool_bytes));
    CU(cudaMemsetAsync(c->pool.p, 0, pool_bytes, c->stream));
    int rc;
    if ((rc = upload(c->tasks, tasks_up, c))) return rc;
    if ((rc = upload(c->pairs, pairs_up, c))) return rc;
    if ((rc = upload(c->succ, succ_up, c))) return rc;
    {
        std::vector<int32_t> d0(tasks_up.size());
        for (size_t t = 0; t < tasks_up.size(); t++) d0[t] = tasks_up[t].n_deps;
        if ((rc = upload(c->dep0, d0, c))) return rc;
    }
    {
        // counter image: per segment one 512-byte record {next task to claim @ int 0}; the abort word sits behind it
        const std::vector<int32_t>& sb = c->dist ? c->D.seg_begin : G.seg_begin;
        const int nseg = (int)sb.size() - 1;
        std::vector<int32_t> c0((size_t)std::max(nseg, 1) * 128, 0);
        if ((rc = upload(c->counters0, c0, c))) return rc;
        CU(c->counters.alloc(c0.size() * 4 + 64));
        CU(cudaMemsetAsync(c->counters.as<char>() + c0.size() * 4, 0, 64, c->stream));
    }
    CU(c->dep.alloc(std::max<size_t>(tasks_up.size(), 1) * 4));
    CU(c->ready.alloc(std::max<size_t>(tasks_up.size(), 1) * 4));
    // level order for the debug executor
    if (!c->dist) {
        const int64_t nt = (int64_t)G.tasks.size();
        c->level_ptr.assign(G.n_levels + 1, 0);
        for (int64_t t = 0; t < nt; t++) c->level_ptr[G.tasks[t].level + 1]++;
        for (int l = 0; l < G.n_levels; l++) c->level_ptr[l + 1] += c->level_ptr[l];
        c->level_order.resize(nt);
        std::vector<int64_t> pos(c->level_ptr.begin(), c->level_ptr.end() - 1);
        for (int64_t t = 0; t < nt; t++) c->level_order[pos[G.tasks[t].level]++] = (int32_t)t;
    }
    // triangular-solve structures, on every rank: a sharded solve gives each GPU the block rows whose diagonal block it owns
    {
        const soglu_ctx* fac = follower ? c->leader : c;       // the factor lists live with the group's rank 0
        c->n_block_rows = fac->n_block_rows;
        c->symmetric = fac->symmetric;
        std::vector<int32_t> ldiag;
        if ((rc = build_tri(c, fac->L_ids, fac->L_brow, fac->L_bcol, false, false, c->l_ptr, c->l_col, c->l_slot, c->l_diag, c->l_dinv, c->nL_off, &ldiag))) return rc;
        if (c->symmetric) {
            if ((rc = build_tri(c, fac->L_ids, fac->L_brow, fac->L_bcol, true, true, c->u_ptr, c->u_col, c->u_slot, c->u_diag, c->u_dinv, c->nU_off))) return rc;
        } else {
            if ((rc = build_tri(c, fac->U_ids, fac->U_brow, fac->U_bcol, true, false, c->u_ptr, c->u_col, c->u_slot, c->u_diag, c->u_dinv, c->nU_off))) return rc;
        }
        std::vector<int32_t> mine;
        for (int r = 0; r < c->n_block_rows; r++)
            if (!c->dist || (int)((uint32_t)ldiag[r] >> REF_SHIFT) == c->rank) mine.push_back(r);
        c->n_my_rows = (int32_t)mine.size();
        if ((rc = upload(c->my_rows, mine, c))) return rc;
        const size_t next = (size_t)c->n_block_rows * BLK * sizeof(double);
        CU(c->d_b.alloc(next));
        CU(c->sv.alloc(6 * next + 64));
        CU(cudaMemsetAsync(c->sv.as<char>() + 6 * next, 0, 64, c->stream));
    }
    c->compiled = true;
    // only now are the op arrays no longer needed on the host: a failure above (pool does not fit, a factor without
    // diagonal block, ...) leaves the context uncompiled WITH its graph, so a retry with other options is possible
    BigVec<int32_t>().swap(c->src); BigVec<int32_t>().swap(c->src2);
    BigVec<int32_t>().swap(c->result); BigVec<int32_t>().swap(c->result2);
    BigVec<uint8_t>().swap(c->op);
    return pack_pending_inputs(c);
}

}  // namespace

static int group_create(soglu_ctx** out, int n_gpus, const int* device_ids);

extern "C" {

int soglu_create(soglu_ctx** out, int n_gpus, const int* device_ids) {
    if (!out) return fail(SOGLU_ERR_ARG, "null output pointer");
    *out = nullptr;
    if (n_gpus > 1) return group_create(out, n_gpus, device_ids);
    if (n_gpus != 1) return fail(SOGLU_ERR_ARG, "n_gpus must be >= 1");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(SOGLU_ERR_NO_DEVICE, std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") +
                                             " (soglu-b200 has no CPU fallback)");
    int dev = device_ids ? device_ids[0] : 0;
    if (dev < 0 || dev >= count) return fail(SOGLU_ERR_ARG, "device id out of range");
    CU(cudaSetDevice(dev));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, dev));
    if (prop.major < 10) return fail(SOGLU_ERR_NO_DEVICE, std::string("device ") + prop.name + " is not sm_100 class; kernels are built for sm_100a only");
    soglu_ctx* c = new soglu_ctx();
    c->device = dev;
    c->sms = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess || cudaEventCreate(&c->ev0) != cudaSuccess ||
        cudaEventCreate(&c->ev1) != cudaSuccess) {
        delete c;
        return fail(SOGLU_ERR_CUDA, "stream/event creation failed");
    }
    c->exec_grid = executor_max_grid(dev);
    c->trsv_grid = trsv_max_grid(dev);
    if (c->exec_grid <= 0 || c->trsv_grid <= 0) {
        std::string m = std::string("kernel image not loadable on this device: ") + cudaGetErrorString(cudaGetLastError());
        soglu_destroy(c);
        return fail(SOGLU_ERR_CUDA, m);
    }
    *out = c;
    return SOGLU_OK;
}

// ---- multi-GPU: one process per GPU, peers mapped through CUDA IPC -------------------------------
struct DistBlob { cudaIpcMemHandle_t pool, dep, counters, sv; int32_t rank, valid; uint64_t layout_hash; };

// FNV-1a over what every rank must agree on: tasks and pool slots per GPU, segment boundaries of every GPU
static uint64_t dist_layout_hash(const soglu_ctx* c) {
    uint64_t h = 1469598103934665603ull;
    auto mix = [&](const void* p, size_t n) { const unsigned char* b = (const unsigned char*)p; for (size_t i = 0; i < n; i++) { h ^= b[i]; h *= 1099511628211ull; } };
    mix(c->D.tasks_per_rank.data(), c->D.tasks_per_rank.size() * sizeof(int64_t));
    mix(c->D.slots_per_rank.data(), c->D.slots_per_rank.size() * sizeof(int64_t));
    for (const auto& v : c->D.seg_begin_all) mix(v.data(), v.size() * sizeof(int32_t));
    return h;
}

int soglu_create_dist(soglu_ctx** out, int device, int rank, int world, int grid_rows, int grid_cols) {
    if (world < 1 || world > MAX_GPUS || grid_rows * grid_cols != world || rank < 0 || rank >= world)
        return fail(SOGLU_ERR_ARG, "bad process grid (world <= 8, grid_rows * grid_cols == world)");
    int dev[1] = {device};
    int rc = soglu_create(out, 1, dev);
    if (rc) return rc;
    soglu_ctx* c = *out;
    c->dist = world > 1;
    c->rank = rank; c->world = world; c->pr = grid_rows; c->pc = grid_cols;
    return SOGLU_OK;
}

int64_t soglu_dist_blob_bytes(void) { return (int64_t)sizeof(DistBlob); }

// compile + allocate this GPU's share, then write the IPC handles of its pool / counters / solve vectors
int soglu_dist_export(soglu_ctx* c, void* blob) {
    try {
    if (c && !c->members.empty()) return fail(SOGLU_ERR_ARG, "in-process multi-GPU context (soglu_create with n_gpus > 1): peers are wired by the library, the soglu_dist_* protocol is for one process per GPU");
    if (!c || !blob) return fail(SOGLU_ERR_ARG, "bad argument");
    CU(cudaSetDevice(c->device));
    int rc = finalize(c);
    if (rc) return rc;
    DistBlob b;
    std::memset(&b, 0, sizeof b);
    b.rank = c->rank; b.valid = 1;
    if (c->dist) {
        b.layout_hash = dist_layout_hash(c);
        CU(cudaIpcGetMemHandle(&b.pool, c->pool.p));
        CU(cudaIpcGetMemHandle(&b.dep, c->dep.p));
        CU(cudaIpcGetMemHandle(&b.counters, c->counters.p));
        CU(cudaIpcGetMemHandle(&b.sv, c->sv.p));
    }
    std::memcpy(blob, &b, sizeof b);
    return SOGLU_OK;
    } catch (const std::bad_alloc&) {
        return fail(SOGLU_ERR_OOM, "out of host memory");
    } catch (const std::exception& e) {
        return fail(SOGLU_ERR_ARG, std::string("internal error: ") + e.what());
    }
}

// all_blobs: world blobs in rank order (gathered by the caller, e.g. torch.distributed.all_gather)
int soglu_dist_import(soglu_ctx* c, const void* all_blobs) {
    try {
    if (c && !c->members.empty()) return fail(SOGLU_ERR_ARG, "in-process multi-GPU context (soglu_create with n_gpus > 1): peers are wired by the library, the soglu_dist_* protocol is for one process per GPU");
    if (!c || !all_blobs) return fail(SOGLU_ERR_ARG, "bad argument");
    if (!c->compiled) return fail(SOGLU_ERR_ARG, "soglu_dist_export must precede soglu_dist_import");
    if (c->peers_ready) return fail(SOGLU_ERR_ARG, "peer handles are already imported (one import per context; destroy it and build a new one to re-shard)");
    CU(cudaSetDevice(c->device));
    const DistBlob* bl = reinterpret_cast<const DistBlob*>(all_blobs);
    for (int g = 0; g < c->world; g++) {
        if (g == c->rank || !c->dist) {
            c->peer_pool[g] = c->pool.p; c->peer_dep[g] = c->dep.p; c->peer_counters[g] = c->counters.p; c->peer_sv[g] = c->sv.p;
            continue;
        }
        if (!bl[g].valid || bl[g].rank != g) return fail(SOGLU_ERR_ARG, "peer handle blob " + std::to_string(g) + " is missing or out of order");
        if (bl[g].layout_hash != dist_layout_hash(c))
            return fail(SOGLU_ERR_ARG, "rank " + std::to_string(g) + " compiled a different sharded layout (tasks / pool slots / segments per GPU): the ranks must "
                                       "load the same problem with the same options; set max_slots explicitly if the GPUs differ");
        CU(cudaIpcOpenMemHandle(&c->peer_pool[g], bl[g].pool, cudaIpcMemLazyEnablePeerAccess));
        CU(cudaIpcOpenMemHandle(&c->peer_dep[g], bl[g].dep, cudaIpcMemLazyEnablePeerAccess));
        CU(cudaIpcOpenMemHandle(&c->peer_counters[g], bl[g].counters, cudaIpcMemLazyEnablePeerAccess));
        CU(cudaIpcOpenMemHandle(&c->peer_sv[g], bl[g].sv, cudaIpcMemLazyEnablePeerAccess));
    }
    c->peers_ready = true;
    return SOGLU_OK;
    } catch (const std::bad_alloc&) {
        return fail(SOGLU_ERR_OOM, "out of host memory");
    } catch (const std::exception& e) {
        return fail(SOGLU_ERR_ARG, std::string("internal error: ") + e.what());
    }
}

// reset this GPU's dependency counters and claim counters; the caller must put a barrier across all
// ranks between soglu_dist_reset and soglu_factor, and another one after soglu_factor
int soglu_dist_reset(soglu_ctx* c) {
    if (c && !c->members.empty()) return fail(SOGLU_ERR_ARG, "in-process multi-GPU context (soglu_create with n_gpus > 1): peers are wired by the library, the soglu_dist_* protocol is for one process per GPU");
    if (!c || !c->compiled) return fail(SOGLU_ERR_ARG, "nothing compiled");
    CU(cudaSetDevice(c->device));
    int rc = pack_pending_inputs(c);
    if (rc) return rc;
    const size_t nt = c->dist ? c->D.tasks.size() : c->Gp->tasks.size();
    if (nt) {
        CU(cudaMemcpyAsync(c->dep.p, c->dep0.p, nt * 4, cudaMemcpyDeviceToDevice, c->stream));
    }
    CU(cudaMemcpyAsync(c->counters.p, c->counters0.p, c->counters0.bytes, cudaMemcpyDeviceToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    c->dist_segment = 0;
    return SOGLU_OK;
}

// number of executor launches (segments) one factorisation needs; > 1 when the pools recycle slots
int soglu_dist_segments(soglu_ctx* c) {
    if (c && !c->members.empty()) c = c->members[0];
    if (!c || !c->compiled) return -1;
    return (int)((c->dist ? c->D.seg_begin.size() : c->Gp->seg_begin.size()) - 1);
}
// select the segment the next soglu_factor call runs (all ranks the same one, barrier in between)
int soglu_dist_set_segment(soglu_ctx* c, int seg) {
    if (c && !c->members.empty()) return fail(SOGLU_ERR_ARG, "in-process multi-GPU context (soglu_create with n_gpus > 1): peers are wired by the library, the soglu_dist_* protocol is for one process per GPU");
    if (!c || !c->compiled) return fail(SOGLU_ERR_ARG, "nothing compiled");
    c->dist_segment = seg;
    return SOGLU_OK;
}

// owned tasks / blocks / remote edges / remote operand reads of this rank
int soglu_dist_info(soglu_ctx* c, int64_t* out4) {
    if (c && !c->members.empty()) return fail(SOGLU_ERR_ARG, "soglu_dist_info: ask per rank (one process per GPU) -- an in-process group has no single rank");
    if (!c || !out4 || !c->compiled) return fail(SOGLU_ERR_ARG, "bad argument");
    out4[0] = (int64_t)(c->dist ? c->D.tasks.size() : c->Gp->tasks.size());
    out4[1] = c->Gp->slots_per_owner[c->dist ? c->rank : 0];
    out4[2] = c->dist ? c->D.remote_edges : 0;
    out4[3] = c->dist ? c->D.remote_operands : 0;
    out4[4] = c->dist ? c->D.mirrored : 0;
    return SOGLU_OK;
}

void soglu_destroy(soglu_ctx* c) {
    if (!c) return;
    if (!c->members.empty()) {
        // all kernels of the group have been synchronised by the calls that launched them; members free their own memory
        for (soglu_ctx* m : c->members) soglu_destroy(m);
        delete c;
        return;
    }
    cudaSetDevice(c->device);
    if (c->dist && c->ipc_peers)
        for (int g = 0; g < c->world; g++) {
            if (g == c->rank) continue;
            for (void* p : {c->peer_pool[g], c->peer_dep[g], c->peer_counters[g], c->peer_sv[g]})
                if (p) cudaIpcCloseMemHandle(p);
        }
    for (DevBuf* b : {&c->in_dense, &c->in_entry_input, &c->in_entry_pos, &c->in_entry_val, &c->in_slots, &c->pool, &c->tasks, &c->pairs, &c->succ, &c->dep0, &c->dep, &c->ready, &c->counters, &c->counters0,
                      &c->l_ptr, &c->l_col, &c->l_slot, &c->l_diag, &c->l_dinv, &c->u_ptr, &c->u_col, &c->u_slot, &c->u_diag, &c->u_dinv, &c->d_b, &c->sv, &c->my_rows, &c->trace, &c->m_rp, &c->m_ci, &c->m_v, &c->d_xacc})
        b->release();
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

int soglu_set_option(soglu_ctx* c, const char* key, int64_t value) {
    if (!c || !key) return fail(SOGLU_ERR_ARG, "bad argument");
    if (!c->members.empty()) {
        for (soglu_ctx* m : c->members) { int rc = soglu_set_option(m, key, value); if (rc) return rc; }
        return SOGLU_OK;
    }
    std::string k(key);
    if (k == "exec_mode") c->opt_exec_mode = value;
    else if (k == "fuse_sub") { if (c->compiled) return fail(SOGLU_ERR_ARG, "fuse_sub must be set before the first factor"); c->opt_fuse_sub = value; }
    else if (k == "fuse_inv") { if (c->compiled) return fail(SOGLU_ERR_ARG, "fuse_inv must be set before the first factor"); c->opt_fuse_inv = value; }
    else if (k == "split") { if (c->compiled) return fail(SOGLU_ERR_ARG, "split must be set before the first factor"); c->opt_split = value; }
    else if (k == "max_slots") { if (c->compiled) return fail(SOGLU_ERR_ARG, "max_slots must be set before the first factor"); c->opt_max_slots = value; }
    else if (k == "mirror_min") { if (c->compiled) return fail(SOGLU_ERR_ARG, "mirror_min must be set before the first factor"); c->opt_mirror_min = value; }
    else if (k == "dist_nb") { if (c->compiled) return fail(SOGLU_ERR_ARG, "dist_nb must be set before the first factor"); c->opt_dist_nb = value; }
    else if (k == "static_order") { if (c->compiled) return fail(SOGLU_ERR_ARG, "static_order must be set before the first factor"); c->opt_static_order = value; }
    else if (k == "split_slack") { if (c->compiled) return fail(SOGLU_ERR_ARG, "split_slack must be set before the first factor"); c->opt_split_slack = value; }
    else if (k == "grid") c->opt_grid = value;
    else if (k == "trace") c->opt_trace = value;
    else if (k == "watchdog_ms") c->opt_watchdog_ms = value;
    else if (k == "debug_drop_task") c->opt_debug_drop = value;
    else if (k == "order_alpha") { if (c->compiled) return fail(SOGLU_ERR_ARG, "order_alpha must be set before the first factor"); c->opt_order_alpha = std::min<int64_t>(100, std::max<int64_t>(0, value)); }
    else return fail(SOGLU_ERR_ARG, "unknown option " + k);
    return SOGLU_OK;
}

static int register_input_pattern(soglu_ctx* c, int64_t n_block_ids, int64_t n_input, const int32_t* input_ids) {
    if (c->compiled) {
        // refactorisation with new values on the same pattern
        if (n_block_ids != c->n_ids || n_input != c->n_input || std::memcmp(input_ids, c->input_ids.data(), n_input * sizeof(int32_t)) != 0)
            return fail(SOGLU_ERR_ARG, "soglu_set_blocks after compilation must keep the block pattern");
    } else {
        c->n_ids = n_block_ids;
        c->n_input = n_input;
        c->input_ids.assign(input_ids, input_ids + n_input);
    }
    return SOGLU_OK;
}

int soglu_set_blocks(soglu_ctx* c, int64_t n_block_ids, int64_t n_input, const int32_t* input_ids, const double* dense) {
    try {
    if (!c || n_block_ids < 1 || n_input < 0 || (n_input > 0 && (!input_ids || !dense))) return fail(SOGLU_ERR_ARG, "bad argument");
    if (!c->members.empty()) {      // every GPU stages the inputs and keeps its share
        for (soglu_ctx* m : c->members) { int rc = soglu_set_blocks(m, n_block_ids, n_input, input_ids, dense); if (rc) return rc; }
        return SOGLU_OK;
    }
    CU(cudaSetDevice(c->device));
    int rc = register_input_pattern(c, n_block_ids, n_input, input_ids);
    if (rc) return rc;
    const size_t bytes = (size_t)n_input * BLK * BLK * sizeof(double);
    CU(c->in_dense.alloc(bytes));
    if (bytes) CU(cudaMemcpyAsync(c->in_dense.p, dense, bytes, cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    c->h2d += (double)bytes;
    c->n_entries = -1;
    c->inputs_pending = true;
    c->have_blocks = true;
    c->factored = false;
    return SOGLU_OK;
    } catch (const std::bad_alloc&) {
        return fail(SOGLU_ERR_OOM, "out of host memory");
    } catch (const std::exception& e) {
        return fail(SOGLU_ERR_ARG, std::string("internal error: ") + e.what());
    }
}

int soglu_set_blocks_sparse(soglu_ctx* c, int64_t n_block_ids, int64_t n_input, const int32_t* input_ids, int64_t n_entries,
                            const int32_t* entry_input, const int32_t* entry_pos, const double* vals) {
    try {
    if (!c || n_block_ids < 1 || n_input < 0 || n_entries < 0 || (n_input > 0 && !input_ids) || (n_entries > 0 && (!entry_input || !entry_pos || !vals)))
        return fail(SOGLU_ERR_ARG, "bad argument");
    int bad = 0;
    if (!c->members.empty()) {
        for (soglu_ctx* m : c->members) { int rc = soglu_set_blocks_sparse(m, n_block_ids, n_input, input_ids, n_entries, entry_input, entry_pos, vals); if (rc) return rc; }
        return SOGLU_OK;
    }
#pragma omp parallel for schedule(static) reduction(| : bad)
    for (int64_t k = 0; k < n_entries; k++)
        if (entry_input[k] < 0 || entry_input[k] >= n_input || entry_pos[k] < 0 || entry_pos[k] >= BLK * BLK) bad = 1;
    if (bad) return fail(SOGLU_ERR_ARG, "soglu_set_blocks_sparse: entry outside its block or input list");
    CU(cudaSetDevice(c->device));
    int rc = register_input_pattern(c, n_block_ids, n_input, input_ids);
    if (rc) return rc;
    CU(c->in_entry_input.reserve(std::max<size_t>(n_entries, 1) * 4));
    CU(c->in_entry_pos.reserve(std::max<size_t>(n_entries, 1) * 4));
    CU(c->in_entry_val.reserve(std::max<size_t>(n_entries, 1) * 8));
    if (n_entries) {
        CU(cudaMemcpyAsync(c->in_entry_input.p, entry_input, (size_t)n_entries * 4, cudaMemcpyHostToDevice, c->stream));
        CU(cudaMemcpyAsync(c->in_entry_pos.p, entry_pos, (size_t)n_entries * 4, cudaMemcpyHostToDevice, c->stream));
        CU(cudaMemcpyAsync(c->in_entry_val.p, vals, (size_t)n_entries * 8, cudaMemcpyHostToDevice, c->stream));
    }
    CU(cudaStreamSynchronize(c->stream));
    c->h2d += (double)n_entries * 16.0;
    c->in_dense.release();
    c->n_entries = n_entries;
    c->inputs_pending = true;
    c->have_blocks = true;
    c->factored = false;
    return SOGLU_OK;
    } catch (const std::bad_alloc&) {
        return fail(SOGLU_ERR_OOM, "out of host memory");
    } catch (const std::exception& e) {
        return fail(SOGLU_ERR_ARG, std::string("internal error: ") + e.what());
    }
}

int soglu_set_graph(soglu_ctx* c, int64_t n_ops, const int32_t* src, const int32_t* src2, const uint8_t* op, const int32_t* result,
                    const int32_t* result2, const int32_t* stage, const int32_t* block_row, const int32_t* block_col) {
    try {
    (void)stage;
    if (!c || n_ops < 0 || (n_ops > 0 && (!src || !src2 || !op || !result || !result2))) return fail(SOGLU_ERR_ARG, "bad argument");
    if (!c->members.empty()) {      // ONE copy of the operation list, ONE compilation: rank 0 holds the graph for the group
        if (!block_row || !block_col) return fail(SOGLU_ERR_ARG, "multi-GPU context: soglu_set_graph needs block_row / block_col (they decide the block ownership)");
        return soglu_set_graph(c->members[0], n_ops, src, src2, op, result, result2, stage, block_row, block_col);
    }
    if (c->compiled) return fail(SOGLU_ERR_ARG, "graph already compiled; create a new context for a new pattern");
    c->n_ops = n_ops;
    c->src.resize(n_ops); c->src2.resize(n_ops); c->result.resize(n_ops); c->result2.resize(n_ops); c->op.resize(n_ops);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n_ops; i++) { c->src[i] = src[i]; c->src2[i] = src2[i]; c->result[i] = result[i]; c->result2[i] = result2[i]; c->op[i] = op[i]; }
    if (block_row && block_col) {
        if (!c->have_blocks) return fail(SOGLU_ERR_ARG, "soglu_set_blocks must precede soglu_set_graph when block coordinates are passed");
        c->brow.assign(block_row, block_row + c->n_ids);
        c->bcol.assign(block_col, block_col + c->n_ids);
    }
    c->have_graph = true;
    return SOGLU_OK;
    } catch (const std::bad_alloc&) {
        return fail(SOGLU_ERR_OOM, "out of host memory");
    } catch (const std::exception& e) {
        return fail(SOGLU_ERR_ARG, std::string("internal error: ") + e.what());
    }
}

int soglu_set_factors(soglu_ctx* c, int64_t nL, const int32_t* L_ids, const int32_t* L_brow, const int32_t* L_bcol, int64_t nU,
                      const int32_t* U_ids, const int32_t* U_brow, const int32_t* U_bcol, int32_t n_block_rows, int symmetric) {
    try {
    if (!c || nL <= 0 || !L_ids || !L_brow || !L_bcol || n_block_rows <= 0) return fail(SOGLU_ERR_ARG, "bad argument");
    if (!c->members.empty()) {
        for (soglu_ctx* m : c->members) m->n_block_rows = n_block_rows;
        return soglu_set_factors(c->members[0], nL, L_ids, L_brow, L_bcol, nU, U_ids, U_brow, U_bcol, n_block_rows, symmetric);
    }
    if (!symmetric && (nU <= 0 || !U_ids || !U_brow || !U_bcol)) return fail(SOGLU_ERR_ARG, "U factor missing");
    if (c->compiled) return fail(SOGLU_ERR_ARG, "graph already compiled; create a new context for a new pattern");
    c->L_ids.assign(L_ids, L_ids + nL); c->L_brow.assign(L_brow, L_brow + nL); c->L_bcol.assign(L_bcol, L_bcol + nL);
    if (!symmetric) { c->U_ids.assign(U_ids, U_ids + nU); c->U_brow.assign(U_brow, U_brow + nU); c->U_bcol.assign(U_bcol, U_bcol + nU); }
    c->n_block_rows = n_block_rows;
    c->symmetric = symmetric ? 1 : 0;
    c->have_factors = true;
    return SOGLU_OK;
    } catch (const std::bad_alloc&) {
        return fail(SOGLU_ERR_OOM, "out of host memory");
    } catch (const std::exception& e) {
        return fail(SOGLU_ERR_ARG, std::string("internal error: ") + e.what());
    }
}

// enqueue the executor launch(es) of one factorisation (sharded: of the selected segment) on the context's stream
static int factor_launch(soglu_ctx* c) {
    CU(cudaSetDevice(c->device));
    int rc = finalize(c);
    if (rc) return rc;
    TaskGraph& G = *c->Gp;
    const int32_t nt = (int32_t)(c->dist ? c->D.tasks.size() : G.tasks.size());
    int grid = c->exec_grid;
    if (c->opt_grid > 0 && c->opt_grid < grid) grid = (int)c->opt_grid;
    ExecParams P = {};
    P.pool = c->pool.as<double>();
    P.world = c->world; P.rank = c->rank;
    if (c->dist) {
        if (!c->peers_ready) return fail(SOGLU_ERR_ARG, "multi-GPU context: exchange peer handles (soglu_dist_export / soglu_dist_import) and call soglu_dist_reset before soglu_factor");
        for (int g = 0; g < c->world; g++) { P.pools[g] = (double*)c->peer_pool[g]; P.deps[g] = (int32_t*)c->peer_dep[g]; }
    } else {
        P.pools[0] = c->pool.as<double>();
    }
    P.tasks = c->tasks.as<Task>();
    P.pairs = c->pairs.as<Pair>();
    P.succ = c->succ.as<int32_t>();
    P.dep = c->dep.as<int32_t>();
    P.trace = nullptr;
    P.abort = c->abort_word();
    P.watchdog_ns = (unsigned long long)std::max<int64_t>(0, c->opt_watchdog_ms) * 1000000ull;
    P.debug_drop = (int32_t)c->opt_debug_drop;
    if (c->dist) for (int g = 0; g < c->world; g++) P.aborts[g] = (int32_t*)c->peer_counters[g] + c->counters0.bytes / 4;
    if (c->opt_trace && nt > 0) {
        if (!c->trace.p) CU(c->trace.alloc((size_t)nt * 6 * sizeof(unsigned long long)));
        if (!c->dist || c->dist_segment == 0) CU(cudaMemsetAsync(c->trace.p, 0, (size_t)nt * 6 * sizeof(unsigned long long), c->stream));   // (all segments of one factorisation)
        P.trace = c->trace.as<unsigned long long>();
    }
    // claim counter and task range of one segment
    auto set_segment = [&](int sg) {
        const std::vector<int32_t>& sb = c->dist ? c->D.seg_begin : G.seg_begin;
        int32_t* cnt = c->counters.as<int32_t>() + (size_t)sg * 128;
        P.ready = nullptr;
        P.head = cnt;
        P.n_tasks = sb[sg + 1] - sb[sg];
        P.task0 = sb[sg];
    };
    CU(cudaEventRecord(c->ev0, c->stream));
    if (nt > 0) {
        if (c->dist) {
            // the counters were reset by soglu_dist_reset (all ranks, then a barrier) -- a late
            // reset here could wipe a signal a faster peer has already delivered
            const int nseg = (int)c->D.seg_begin.size() - 1;
            const int sg = c->dist_segment;
            if (sg < 0 || sg >= nseg) return fail(SOGLU_ERR_ARG, "segment out of range");
            set_segment(sg);
            P.signal = 1;
            if (P.n_tasks > 0) {
                CU(launch_executor(P, grid, c->stream));
                c->launches++;
            }
        } else if (c->opt_exec_mode == 0) {
            const int nseg = (int)G.seg_begin.size() - 1;
            CU(cudaMemcpyAsync(c->dep.p, c->dep0.p, (size_t)nt * 4, cudaMemcpyDeviceToDevice, c->stream));
            CU(cudaMemcpyAsync(c->counters.p, c->counters0.p, c->counters0.bytes, cudaMemcpyDeviceToDevice, c->stream));
            P.signal = 1;
            // one persistent launch per segment (a single one unless the pool forces slot recycling)
            for (int sg = 0; sg < nseg; sg++) {
                set_segment(sg);
                if (P.n_tasks == 0) continue;
                CU(launch_executor(P, grid, c->stream));
                c->launches++;
            }
        } else {
            // debug: one launch per dependency level, no in-kernel signalling
            for (int l = 0; l < G.n_levels; l++) {
                const int64_t b = c->level_ptr[l], e = c->level_ptr[l + 1];
                CU(cudaMemcpyAsync(c->ready.p, c->level_order.data() + b, (size_t)(e - b) * 4, cudaMemcpyHostToDevice, c->stream));
                CU(cudaMemsetAsync(c->counters.p, 0, 512, c->stream));
                int32_t* cnt = c->counters.as<int32_t>();
                P.ready = c->ready.as<int32_t>();
                P.head = cnt;
                P.n_tasks = (int32_t)(e - b);
                P.signal = 0;
                CU(launch_executor(P, (int)std::min<int64_t>(grid, e - b), c->stream));
                c->launches++;
            }
        }
    }
    CU(cudaEventRecord(c->ev1, c->stream));
    return SOGLU_OK;
}

// wait for factor_launch, check the watchdog; ms = device time between the events
static int factor_finish(soglu_ctx* c, float* ms) {
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    int rc = check_watchdog(c, "factorisation");
    if (rc) return rc;
    CU(cudaEventElapsedTime(ms, c->ev0, c->ev1));
    c->factored = true;
    return SOGLU_OK;
}

static int group_factor(soglu_ctx* grp, soglu_stats* out);

int soglu_factor(soglu_ctx* c, soglu_stats* out) {
    try {
    if (!c) return fail(SOGLU_ERR_ARG, "null context");
    if (!c->members.empty()) return group_factor(c, out);
    const int64_t launches0 = c->launches;
    int rc = factor_launch(c);
    if (rc) return rc;
    float ms = 0;
    if ((rc = factor_finish(c, &ms))) return rc;
    if (out) {
        std::memset(out, 0, sizeof *out);
        out->seconds = ms * 1e-3;
        out->flops = c->Gp->flops;
        out->bytes = 0;
        out->kernel_launches = c->launches - launches0;
        out->tasks = (int64_t)(c->dist ? c->D.tasks.size() : c->Gp->tasks.size());
        out->pool_blocks = c->Gp->slots_per_owner[c->dist ? c->rank : 0];
        out->h2d_bytes = c->h2d;
        out->d2h_bytes = c->d2h;
    }
    return SOGLU_OK;
    } catch (const std::bad_alloc&) {
        return fail(SOGLU_ERR_OOM, "out of host memory");
    } catch (const std::exception& e) {
        return fail(SOGLU_ERR_ARG, std::string("internal error: ") + e.what());
    }
}

// vector k of GPU g's solve buffer: 0 / 1 = y, x of set 0; 2 / 3 = y, x of set 1; 4 / 5 = r of refinement step 0 / 1
static double* sv_vec(const soglu_ctx* c, int g, int k) {
    char* base = (char*)(c->dist ? c->peer_sv[g] : c->sv.p);
    return (double*)(base + (size_t)k * c->n_block_rows * BLK * sizeof(double));
}

// one forward/back substitution: rhs (n_ext, local; polled if it is a refinement residual) -> x of `set` on every GPU
static int run_trsv(soglu_ctx* c, const double* d_rhs, bool rhs_polled, int set) {
    TrsvParams P = {};
    P.pool = c->pool.as<double>();
    P.world = c->dist ? c->world : 1;
    P.rank = c->dist ? c->rank : 0;
    for (int g = 0; g < P.world; g++) {
        P.pools[g] = c->dist ? (const double*)c->peer_pool[g] : c->pool.as<double>();
        P.y_all[g] = sv_vec(c, g, 2 * set);
        P.x_all[g] = sv_vec(c, g, 2 * set + 1);
    }
    P.y = P.y_all[P.rank]; P.x = P.x_all[P.rank];
    P.l_ptr = c->l_ptr.as<int64_t>(); P.l_col = c->l_col.as<int32_t>(); P.l_slot = c->l_slot.as<int32_t>(); P.l_diag = c->l_diag.as<int32_t>();
    P.l_dinv = c->l_dinv.as<int32_t>();
    P.u_ptr = c->u_ptr.as<int64_t>(); P.u_col = c->u_col.as<int32_t>(); P.u_slot = c->u_slot.as<int32_t>(); P.u_diag = c->u_diag.as<int32_t>();
    P.u_dinv = c->u_dinv.as<int32_t>();
    P.n_rows = c->n_block_rows;
    P.my_rows = c->my_rows.as<int32_t>(); P.n_my_rows = c->n_my_rows;
    P.b = d_rhs; P.b_polled = rhs_polled ? 1 : 0;
    P.symmetric = c->symmetric;
    P.abort = c->abort_word();
    P.watchdog_ns = (unsigned long long)std::max<int64_t>(0, c->opt_watchdog_ms) * 1000000ull;
    if (c->n_my_rows > 0) {
        CU(launch_trsv(P, std::min(c->trsv_grid, (int)c->n_my_rows), c->stream));
        c->launches++;
    }
    return SOGLU_OK;
}

// Enqueue solve + `refine` refinement steps on the context's stream.  Sharded: EVERY rank calls this with the same b and
// the same `refine` (a collective, like soglu_factor); each GPU solves its block rows, rank 0 forms the residuals
// (soglu_set_matrix is only needed there) and holds the answer.
static int solve_launch(soglu_ctx* c, const double* b_ext, int refine) {
    if (!c->factored) return fail(SOGLU_ERR_ARG, "soglu_factor must precede soglu_solve");
    const bool lead = !c->dist || c->rank == 0;
    if (refine > 0 && lead && c->m_n == 0) return fail(SOGLU_ERR_ARG, "iterative refinement needs the matrix: call soglu_set_matrix first");
    if (refine > 1 && c->dist) return fail(SOGLU_ERR_ARG, "a sharded solve supports at most one refinement step");
    CU(cudaSetDevice(c->device));
    const int64_t n = (int64_t)c->n_block_rows * BLK;
    const size_t next = (size_t)n * sizeof(double);
    CU(cudaMemcpyAsync(c->d_b.p, b_ext, next, cudaMemcpyHostToDevice, c->stream));
    CU(cudaEventRecord(c->ev0, c->stream));
    // every vector of this call starts as "not yet computed" BEFORE the first kernel: a peer's residual / segments may
    // arrive any time after this GPU's first solve kernel has started
    CU(launch_fill_sentinel(c->sv.as<double>(), (refine > 0 ? 6 : 2) * n, c->stream));
    c->launches++;
    if (c->dist) {
        // ... and no GPU publishes into a peer's vectors before that peer has filled them: device-side barrier
        int32_t* flags_all[MAX_GPUS];
        for (int g = 0; g < c->world; g++) flags_all[g] = (int32_t*)((char*)c->peer_sv[g] + 6 * next);
        CU(launch_peer_epoch(flags_all, c->world, c->rank, ++c->solve_epoch, c->abort_word(),
                             (unsigned long long)std::max<int64_t>(0, c->opt_watchdog_ms) * 1000000ull, c->stream));
        c->launches++;
    }
    int rc = run_trsv(c, c->d_b.as<double>(), false, 0);
    if (rc) return rc;
    int set = 0;
    for (int it = 0; it < refine; it++) {
        // x_acc = x; r = b - A x_acc (rank 0, written into every GPU's r); solve A d = r; x_acc += d   (all FP64, on the device)
        double* r_mine = sv_vec(c, c->dist ? c->rank : 0, 4 + (it & 1));
        if (lead) {
            if (!c->d_xacc.p) CU(c->d_xacc.alloc(next));
            if (it == 0) CU(cudaMemcpyAsync(c->d_xacc.p, sv_vec(c, c->dist ? c->rank : 0, 1), next, cudaMemcpyDeviceToDevice, c->stream));
            double* r_all[MAX_GPUS];
            const int world = c->dist ? c->world : 1;
            for (int g = 0; g < world; g++) r_all[g] = sv_vec(c, g, 4 + (it & 1));
            CU(launch_residual(c->m_rp.as<int64_t>(), c->m_ci.as<int32_t>(), c->m_v.as<double>(), c->d_b.as<double>(), c->d_xacc.as<double>(), r_all, world, n, c->stream));
            c->launches++;
        }
        set ^= 1;
        if (it >= 1) {      // (single GPU only) the set and the r vector of two steps ago are reused: same stream, so refilling is safe
            CU(launch_fill_sentinel(sv_vec(c, 0, 2 * set), 2 * n, c->stream));
        }
        if ((rc = run_trsv(c, r_mine, true, set))) return rc;
        if (lead) {
            CU(launch_axpy(c->d_xacc.as<double>(), sv_vec(c, c->dist ? c->rank : 0, 2 * set + 1), n, c->stream));
            c->launches++;
            if (it + 2 < refine) CU(launch_fill_sentinel(sv_vec(c, 0, 4 + (it & 1)), n, c->stream));   // r of this step, reused at it + 2
        }
    }
    CU(cudaEventRecord(c->ev1, c->stream));
    return SOGLU_OK;
}

static int solve_finish(soglu_ctx* c, double* x_ext, int refine, int64_t launches0, soglu_stats* out) {
    CU(cudaSetDevice(c->device));
    const size_t next = (size_t)c->n_block_rows * BLK * sizeof(double);
    const bool lead = !c->dist || c->rank == 0;
    if (lead && x_ext) CU(cudaMemcpyAsync(x_ext, refine > 0 ? c->d_xacc.p : (void*)sv_vec(c, c->dist ? c->rank : 0, 1), next, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    int rc = check_watchdog(c, "solve");
    if (rc) return rc;
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    c->h2d += (double)next;
    c->d2h += (double)next;
    if (out) {
        std::memset(out, 0, sizeof *out);
        out->seconds = ms * 1e-3;
        const double nblk = (double)(c->nL_off + c->nU_off + 2.0 * c->n_block_rows) * (1 + refine);
        out->flops = 2.0 * 4096.0 * nblk;
        out->bytes = (double)BLK * BLK * 8.0 * nblk + 8.0 * 3.0 * c->n_block_rows * BLK * (1 + refine);
        out->kernel_launches = c->launches - launches0;
        out->tasks = 2 * (int64_t)c->n_block_rows * (1 + refine);
        out->pool_blocks = c->Gp->slots_per_owner[c->dist ? c->rank : 0];
        out->h2d_bytes = (double)next;
        out->d2h_bytes = (double)next;
    }
    return SOGLU_OK;
}

static int solve_impl(soglu_ctx* c, const double* b_ext, double* x_ext, int refine, soglu_stats* out) {
    if (!c || !b_ext) return fail(SOGLU_ERR_ARG, "bad argument");
    if (!c->members.empty()) {
        // in-process group: launch on every GPU, then collect; rank 0 holds x
        if (!x_ext) return fail(SOGLU_ERR_ARG, "bad argument");
        std::vector<int64_t> l0;
        for (soglu_ctx* m : c->members) { l0.push_back(m->launches); int rc = solve_launch(m, b_ext, refine); if (rc) return rc; }
        int first_err = 0;
        std::string first_msg;
        for (size_t g = c->members.size(); g-- > 0;) {       // rank 0 last: its stats are the ones reported
            int rc = solve_finish(c->members[g], g == 0 ? x_ext : nullptr, refine, l0[g], g == 0 ? out : nullptr);
            if (rc && !first_err) { first_err = rc; first_msg = soglu_last_error(); }
        }
        return first_err ? fail(first_err, first_msg) : SOGLU_OK;
    }
    if (!x_ext && (!c->dist || c->rank == 0)) return fail(SOGLU_ERR_ARG, "bad argument");
    if (c->dist && !c->peers_ready) return fail(SOGLU_ERR_ARG, "multi-GPU context: exchange peer handles first (soglu_dist_export / soglu_dist_import)");
    const int64_t launches0 = c->launches;
    int rc = solve_launch(c, b_ext, refine);
    if (rc) return rc;
    return solve_finish(c, x_ext, refine, launches0, out);
}

static int solve_guarded(soglu_ctx* c, const double* b_ext, double* x_ext, int refine, soglu_stats* out) {
    try {
        return solve_impl(c, b_ext, x_ext, refine, out);
    } catch (const std::bad_alloc&) {
        return fail(SOGLU_ERR_OOM, "out of host memory");
    } catch (const std::exception& e) {
        return fail(SOGLU_ERR_ARG, std::string("internal error: ") + e.what());
    }
}

int soglu_solve(soglu_ctx* c, const double* b_ext, double* x_ext, soglu_stats* out) { return solve_guarded(c, b_ext, x_ext, 0, out); }

// solve + `steps` rounds of iterative refinement on the device (needs soglu_set_matrix)
int soglu_solve_refined(soglu_ctx* c, const double* b_ext, double* x_ext, int steps, soglu_stats* out) {
    return solve_guarded(c, b_ext, x_ext, steps < 0 ? 0 : steps, out);
}

// CSR of the permuted system padded with the identity to n_block_rows*64 rows (what the factors factorise)
int soglu_set_matrix(soglu_ctx* c, int64_t n_ext, int64_t nnz, const int64_t* row_ptr, const int32_t* col, const double* val) {
    try {
    if (!c || n_ext <= 0 || nnz < 0 || !row_ptr || (nnz > 0 && (!col || !val))) return fail(SOGLU_ERR_ARG, "bad argument");
    if (!c->members.empty()) return soglu_set_matrix(c->members[0], n_ext, nnz, row_ptr, col, val);
    CU(cudaSetDevice(c->device));
    CU(c->m_rp.alloc((size_t)(n_ext + 1) * 8)); CU(c->m_ci.alloc(std::max<size_t>(nnz, 1) * 4)); CU(c->m_v.alloc(std::max<size_t>(nnz, 1) * 8));
    CU(cudaMemcpyAsync(c->m_rp.p, row_ptr, (size_t)(n_ext + 1) * 8, cudaMemcpyHostToDevice, c->stream));
    if (nnz) {
        CU(cudaMemcpyAsync(c->m_ci.p, col, (size_t)nnz * 4, cudaMemcpyHostToDevice, c->stream));
        CU(cudaMemcpyAsync(c->m_v.p, val, (size_t)nnz * 8, cudaMemcpyHostToDevice, c->stream));
    }
    CU(cudaStreamSynchronize(c->stream));
    c->m_n = n_ext; c->m_nnz = nnz;
    c->h2d += (double)((n_ext + 1) * 8 + nnz * 12);
    return SOGLU_OK;
    } catch (const std::bad_alloc&) {
        return fail(SOGLU_ERR_OOM, "out of host memory");
    } catch (const std::exception& e) {
        return fail(SOGLU_ERR_ARG, std::string("internal error: ") + e.what());
    }
}

// debug: cycles of lu / write-out / inverses / total for one diagonal block (slot 1 = first input)
int soglu_debug_diag_bench(soglu_ctx* c, int iters, long long* cycles4) {
    if (!c || !c->compiled || c->Gp->n_slots < 12) return fail(SOGLU_ERR_ARG, "need a compiled problem");
    DevBuf d;
    CU(d.alloc(128));
    CU(cudaMemsetAsync(d.p, 0, 128, c->stream));
    CU(launch_diag_bench(c->pool.as<double>(), iters, d.as<long long>(), c->stream));
    CU(cudaMemcpyAsync(cycles4, d.p, 96, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    d.release();
    return SOGLU_OK;
}

// debug: copy the per-task trace (6 x u64 per task) and the task table (type, n_pairs, level,
// n_deps per task) of the last traced soglu_factor; returns the number of tasks
int64_t soglu_debug_trace(soglu_ctx* c, unsigned long long* trace_out, int32_t* task_info_out, int32_t* succ_ptr_out, int32_t* succ_out) {
    if (!c || !c->compiled) return -1;
    const int64_t nt = (int64_t)c->Gp->tasks.size();
    if (trace_out) {
        if (!c->trace.p) return -1;
        if (cudaMemcpy(trace_out, c->trace.p, (size_t)nt * 6 * sizeof(unsigned long long), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    }
    if (task_info_out)
        for (int64_t t = 0; t < nt; t++) {
            const Task& T = c->Gp->tasks[t];
            task_info_out[4 * t] = T.type; task_info_out[4 * t + 1] = T.n_pairs; task_info_out[4 * t + 2] = T.level; task_info_out[4 * t + 3] = T.n_deps;
        }
    // successor lists name group leaders and are shared by the slices of a task: export them expanded to
    // one plain CSR (every slice -> every slice of every successor) for the analysis scripts
    if (succ_ptr_out || succ_out) {
        int64_t pos = 0;
        for (int64_t t = 0; t < nt; t++) {
            const Task& T = c->Gp->tasks[t];
            if (succ_ptr_out) succ_ptr_out[t] = (int32_t)pos;
            for (int32_t e = T.succ_begin; e < T.succ_end; e++) {
                const int32_t s2 = c->Gp->succ[e];
                for (int q = 0, g = task_group_size(c->Gp->tasks[s2]); q < g; q++, pos++)
                    if (succ_out) succ_out[pos] = s2 + q;
            }
        }
        if (succ_ptr_out) succ_ptr_out[nt] = (int32_t)pos;
    }
    return nt;
}

// debug: utilisation profile of the last traced factorisation on THIS GPU (works per rank of a sharded run; the idle
// time between the segments of a recycled pool is part of the makespan).  out[0] = tasks with a trace, out[1] = makespan (ns, first claim to last signal), out[2..4] = summed ns
// of operand wait (claimed -> loaded), math (loaded -> computed), release (computed -> signalled), out[5] = CTAs,
// out[8 + b] = busy CTA-ns (loaded -> computed) in time bin b of nbins equal bins over the makespan.
int soglu_debug_trace_summary(soglu_ctx* c, int nbins, double* out) {
    if (!c || !out || nbins < 1 || !c->compiled || !c->trace.p) return fail(SOGLU_ERR_ARG, "no trace (set option trace = 1 before soglu_factor)");
    if (!c->members.empty()) return fail(SOGLU_ERR_ARG, "ask the members (one process per GPU)");
    const int64_t nt = (int64_t)(c->dist ? c->D.tasks.size() : c->Gp->tasks.size());
    std::vector<unsigned long long> tr((size_t)nt * 6);
    CU(cudaSetDevice(c->device));
    CU(cudaMemcpy(tr.data(), c->trace.p, tr.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    unsigned long long t0 = ~0ull, t1 = 0;
    int64_t seen = 0;
    for (int64_t t = 0; t < nt; t++) {
        const unsigned long long* r = &tr[6 * t];
        if (r[1] == 0 || r[3] == 0) continue;
        seen++;
        t0 = std::min(t0, r[1]);
        t1 = std::max(t1, std::max(r[3], r[4]));
    }
    for (int k = 0; k < 8 + nbins; k++) out[k] = 0;
    out[0] = (double)seen;
    if (!seen) return SOGLU_OK;
    const double span = (double)(t1 - t0), bw = span / nbins;
    out[1] = span;
    for (int64_t t = 0; t < nt; t++) {
        const unsigned long long* r = &tr[6 * t];
        if (r[1] == 0 || r[3] == 0) continue;
        out[2] += (double)(r[2] - r[1]);
        out[3] += (double)(r[3] - r[2]);
        if (r[4] > r[3]) out[4] += (double)(r[4] - r[3]);
        double a = (double)(r[2] - t0), b = (double)(r[3] - t0);
        for (int k = std::max(0, (int)(a / bw)); k < nbins && k * bw < b; k++) out[8 + k] += std::min(b, (k + 1) * bw) - std::max(a, k * bw);
    }
    int grid = c->exec_grid;
    if (c->opt_grid > 0 && c->opt_grid < grid) grid = (int)c->opt_grid;
    out[5] = grid;
    return SOGLU_OK;
}

// diagonal blocks of the last soglu_factor whose U U^-1 failed the reference's inv_check_diag (1 +- 1e-3 on the diagonal;
// MatrixStdDouble.cpp:2871, printed as " upper out of tolerance" at BlockPlanner.cpp:575-577): 0 for a healthy factorisation
int64_t soglu_diag_warnings(soglu_ctx* c) {
    if (!c) return -1;
    if (!c->members.empty()) { int64_t n = 0; for (soglu_ctx* m : c->members) n += m->diag_warnings; return n; }
    return c->diag_warnings;
}

int soglu_get_block(soglu_ctx* c, int32_t id, double* out_64x64) {
    try {
    if (!c || !out_64x64) return fail(SOGLU_ERR_ARG, "bad argument");
    if (!c->members.empty()) {
        soglu_ctx* m0 = c->members[0];
        if (!m0->compiled) return fail(SOGLU_ERR_ARG, "nothing compiled yet");
        if (id <= 0 || id >= m0->n_ids) return fail(SOGLU_ERR_ARG, "block id out of range");
        return soglu_get_block(c->members[m0->Gp->owner_of[id]], id, out_64x64);
    }
    if (!c->compiled) return fail(SOGLU_ERR_ARG, "nothing compiled yet");
    if (id <= 0 || id >= c->n_ids) return fail(SOGLU_ERR_ARG, "block id out of range");
    const int32_t slot = c->Gp->slot_of[id];
    if (c->Gp->recycled[id]) return fail(SOGLU_ERR_ARG, "block " + std::to_string(id) + " was recycled (its pool slot was reused after its last reader)");
    if (slot <= 0) return fail(SOGLU_ERR_ARG, "block " + std::to_string(id) + " has no storage (never produced, or folded into a fused task)");
    CU(cudaSetDevice(c->device));
    DevBuf tmp;
    CU(tmp.alloc(BLK * BLK * sizeof(double)));
    const int32_t local = slot;
    if (c->dist && c->Gp->owner_of[id] != c->rank) { tmp.release(); return fail(SOGLU_ERR_ARG, "block " + std::to_string(id) + " lives on another GPU"); }
    CU(launch_unpack_block(c->pool.as<double>(), local, tmp.as<double>(), c->stream));
    c->launches++;
    CU(cudaMemcpyAsync(out_64x64, tmp.p, BLK * BLK * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    tmp.release();
    return SOGLU_OK;
    } catch (const std::bad_alloc&) {
        return fail(SOGLU_ERR_OOM, "out of host memory");
    } catch (const std::exception& e) {
        return fail(SOGLU_ERR_ARG, std::string("internal error: ") + e.what());
    }
}

}  // extern "C"

// ---- in-process multi-GPU: soglu_create(n_gpus > 1) ----------------------------------------------------------------
// The group context owns one ordinary sharded context per GPU (rank g on device_ids[g], the same 2D block-cyclic
// ownership as the one-process-per-GPU mode) and forwards the ABI to them.  Differences to that mode: ONE copy of the
// operation list and ONE compilation (rank 0's TaskGraph, shared by pointer -- the other ranks only localise their
// part), peers reached through cudaDeviceEnablePeerAccess pointers instead of IPC mappings, and the per-segment
// barriers are this thread waiting for every stream.  This is what ./solve and SOGLU::solveLU use (SOGLU_GPUS=N).
static void default_grid(int world, int* pr, int* pc) {
    int r = 1;
    while (r * r * 2 <= world) r *= 2;
    if (world % r) r = 1;
    *pr = r; *pc = world / r;
}

static int group_create(soglu_ctx** out, int n_gpus, const int* device_ids) {
    if (n_gpus > MAX_GPUS) return fail(SOGLU_ERR_ARG, "at most 8 GPUs");
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count < n_gpus)
        return fail(SOGLU_ERR_NO_DEVICE, "soglu_create: " + std::to_string(n_gpus) + " GPUs requested, " + std::to_string(count) + " visible (soglu-b200 has no CPU fallback)");
    int pr, pc;
    default_grid(n_gpus, &pr, &pc);
    soglu_ctx* grp = new soglu_ctx();
    for (int g = 0; g < n_gpus; g++) {
        soglu_ctx* m = nullptr;
        int rc = soglu_create_dist(&m, device_ids ? device_ids[g] : g, g, n_gpus, pr, pc);
        if (rc) { soglu_destroy(grp); return rc; }
        m->ipc_peers = false;
        grp->members.push_back(m);
    }
    for (soglu_ctx* m : grp->members) m->leader = grp->members[0];
    // peer access between every pair (an error other than "already enabled" means the box has no P2P path)
    for (soglu_ctx* a : grp->members)
        for (soglu_ctx* b : grp->members) {
            if (a == b) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, a->device, b->device);
            if (!can) { soglu_destroy(grp); return fail(SOGLU_ERR_NO_DEVICE, "GPUs " + std::to_string(a->device) + " and " + std::to_string(b->device) + " have no peer access"); }
            cudaSetDevice(a->device);
            cudaError_t e = cudaDeviceEnablePeerAccess(b->device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { soglu_destroy(grp); return fail(SOGLU_ERR_CUDA, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e)); }
            cudaGetLastError();
        }
    grp->world = n_gpus;
    *out = grp;
    return SOGLU_OK;
}

static int group_factor(soglu_ctx* grp, soglu_stats* out) {
    std::vector<soglu_ctx*>& M = grp->members;
    int64_t launches0 = 0;
    for (soglu_ctx* m : M) launches0 += m->launches;
    int rc;
    // compile on rank 0 (the first call), allocate + upload everywhere, wire the peers
    for (soglu_ctx* m : M) {
        cudaSetDevice(m->device);
        if ((rc = finalize(m))) return rc;
    }
    if (!M[0]->peers_ready)
        for (soglu_ctx* m : M) {
            for (size_t g = 0; g < M.size(); g++) {
                m->peer_pool[g] = M[g]->pool.p; m->peer_dep[g] = M[g]->dep.p; m->peer_counters[g] = M[g]->counters.p; m->peer_sv[g] = M[g]->sv.p;
            }
            m->peers_ready = true;
        }
    for (soglu_ctx* m : M) if ((rc = soglu_dist_reset(m))) return rc;
    const int nseg = soglu_dist_segments(M[0]);
    double seconds = 0;
    for (int sg = 0; sg < nseg; sg++) {
        for (soglu_ctx* m : M) { m->dist_segment = sg; if ((rc = factor_launch(m))) return rc; }
        float worst = 0;
        int first_err = 0;
        std::string first_msg;
        for (soglu_ctx* m : M) {          // wait for EVERY GPU even if one reports an error (its peers drain through the abort word)
            float ms = 0;
            rc = factor_finish(m, &ms);
            if (rc && !first_err) { first_err = rc; first_msg = soglu_last_error(); }
            worst = std::max(worst, ms);
        }
        if (first_err) return fail(first_err, first_msg);
        seconds += worst * 1e-3;
    }
    if (out) {
        std::memset(out, 0, sizeof *out);
        out->seconds = seconds;
        out->flops = M[0]->Gp->flops;
        for (soglu_ctx* m : M) {
            out->kernel_launches += m->launches;
            out->tasks += (int64_t)m->D.tasks.size();
            out->pool_blocks = std::max<int64_t>(out->pool_blocks, m->Gp->slots_per_owner[m->rank]);
            out->h2d_bytes += m->h2d; out->d2h_bytes += m->d2h;
        }
        out->kernel_launches -= launches0;
    }
    return SOGLU_OK;
}

