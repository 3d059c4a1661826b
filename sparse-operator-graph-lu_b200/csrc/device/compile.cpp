// Host-side compilation of the reference's flat operation list into the task graph of the
// persistent executor.
//
// The reference executes the list stage by stage and relies on "all writers of one result
// sit in one stage and one stream" for race freedom (BlockPlanner.cpp:270-276, 399-416).
// Here every RESULT block becomes one task that applies all of its pending updates
// (one CTA owns one target block, SURVEY.md App. E invariants), and dependencies are the
// true producer -> consumer edges on block ids, so stages are not needed at run time.
//
// Fusions (results unchanged, op list stays the bit-exact input):
//   * `sub` whose subtrahend is a mul/mulneg/mult chain read by nobody else is folded into
//     that chain: R = S2 -/+ sum A*B   (removes one level of the critical path
//     lu -> inv -> mul -> mul -> sub -> lu and one block write + two block reads).
#include "tasks.h"
#include "cost_model.h"

#include <algorithm>
#include <cmath>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <utility>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace soglu {
namespace {
enum : uint8_t { OP_LU = 1, OP_LOWERINV = 2, OP_UPPERINV = 3, OP_SUB = 4, OP_MUL = 8, OP_MULNEG = 9, OP_LLT = 10, OP_MULT = 11 };

struct IdInfo {
    int64_t first_writer = -1;  // op index
    int32_t n_writers = 0;
    int32_t n_readers = 0;      // ops reading the id
    uint8_t kind = 0xff;        // op code of the writers
    bool is_input = false;
    bool keep = false;
    bool fused_away = false;    // product folded into its single `sub` reader
};
}  // namespace

namespace {
// The compilation is a sequence of passes over the operation list / the task array; what one pass leaves for the next
// lives here.  compile_tasks() runs them in order; each returns "" or the violated invariant.
struct Compiler {
    // ---- input (borrowed) -----------------------------------------------------------------------------------------
    const int64_t n_ids_caller, n_input;
    const int32_t* const input_ids;
    const int64_t n_ops;
    const int32_t *const src, *const src2;
    const uint8_t* const op;
    const int32_t *const result, *const result2;
    const std::vector<int32_t>& keep_ids;
    const CompileOptions& opt;
    TaskGraph& G;
    // ---- state shared by the passes -------------------------------------------------------------------------------
    char msg[256];
    const int64_t n_ids;                    // block ids of the caller (mirrors are appended: nid)
    std::vector<IdInfo> info;
    std::vector<int64_t> fused_sub_of;      // product id -> index of the sub op it is folded into
    std::vector<char> op_fused;             // per op: 1 = folded into an lu task, 2 = alias of an earlier inverse
    std::vector<int32_t> alias_to;          // result id -> canonical result id
    std::vector<int64_t> inv_of;            // source id -> first inverse op
    int64_t nt = 0;                         // tasks so far
    int64_t nid = 0;                        // block ids including mirrors
    int nown = 1;
    std::vector<int32_t> seg_of;            // task -> segment
    std::vector<uint8_t> reused_slot;       // block id -> its pool slot is shared in time with another block (recycling)
    std::vector<int64_t> poff;              // predecessor lists: offsets into pflat (capacity per task), pcnt entries each
    BigVec<int32_t> pflat;
    std::vector<int32_t> pcnt;
    // SOGLU_TIMING=1 prints the phase times to stderr (diagnostics only)
    const bool timing = std::getenv("SOGLU_TIMING") != nullptr;
    std::chrono::steady_clock::time_point t_last = std::chrono::steady_clock::now();

    static constexpr int64_t NONE = INT64_MAX;
    static bool is_acc(uint8_t o) { return o == OP_MUL || o == OP_MULNEG || o == OP_MULT; }
    static void atomic_min(int64_t* p, int64_t v) {
        int64_t cur = __atomic_load_n(p, __ATOMIC_RELAXED);
        while (v < cur && !__atomic_compare_exchange_n(p, &cur, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    }
    void lap(const char* what) {
        if (!timing) return;
        auto t1 = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[compile] %-25s %8.3f s\n", what, std::chrono::duration<double>(t1 - t_last).count());
        t_last = t1;
    }

    Compiler(int64_t n_ids_caller_, int64_t n_input_, const int32_t* input_ids_, int64_t n_ops_, const int32_t* src_, const int32_t* src2_,
             const uint8_t* op_, const int32_t* result_, const int32_t* result2_, const std::vector<int32_t>& keep_ids_, const CompileOptions& opt_,
             TaskGraph& G_)
        : n_ids_caller(n_ids_caller_), n_input(n_input_), input_ids(input_ids_), n_ops(n_ops_), src(src_), src2(src2_), op(op_), result(result_),
          result2(result2_), keep_ids(keep_ids_), opt(opt_), G(G_), n_ids(n_ids_caller_) {}

    std::string run() {
        G = TaskGraph();
        if (n_ids_caller < 1) return "n_block_ids must be >= 1";
        if (n_ids > 0x7fffffff) return "too many block ids";
        if (n_ops > 0x7fffffff) return "too many operations";
        info.assign(n_ids, IdInfo());
        for (int64_t k = 0; k < n_input; k++) {
            int32_t id = input_ids[k];
            if (id <= 0 || id >= n_ids_caller) return "input block id out of range";
            info[id].is_input = true;
        }
        for (int32_t id : keep_ids)
            if (id > 0 && id < n_ids_caller) info[id].keep = true;
        std::string e;
        if (!(e = validate()).empty()) return e;
        if (!(e = decide_fusions()).empty()) return e;
        if (!(e = make_tasks()).empty()) return e;
        if (!(e = fill_pairs()).empty()) return e;
        if (!(e = shard()).empty()) return e;
        if (!(e = assign_slots()).empty()) return e;
        if (!(e = dependencies()).empty()) return e;
        if (!(e = levels()).empty()) return e;
        if (!(e = split_rows()).empty()) return e;
        if (!(e = static_order()).empty()) return e;
        if (!(e = finish()).empty()) return e;
        return "";
    }

    // ==== validation + per-block writer / reader statistics ===================================================
    std::string validate() {
    // ---- validation + per-block writer / reader statistics --------------------------------------
    // All passes over the op list run on every host thread; where the serial order matters (first writer
    // of a block, first inverse of a factor, the lowest offending op) it is recovered with atomic minima.
    auto bad_id = [&](int32_t id) { return id < 0 || id >= n_ids_caller; };
    auto op_error = [&](int64_t i) -> const char* {          // checks that need no other op
        const uint8_t o = op[i];
        if (!(o == OP_LU || o == OP_LOWERINV || o == OP_UPPERINV || o == OP_SUB || o == OP_MUL || o == OP_MULNEG || o == OP_LLT || o == OP_MULT))
            return "unsupported op code";
        if (bad_id(src[i]) || bad_id(src2[i]) || bad_id(result[i]) || bad_id(result2[i]) || result[i] <= 0) return "block id out of range";
        if (o == OP_LU && result2[i] <= 0) return "lu without second result";
        if (src[i] == result[i] || src2[i] == result[i] || (result2[i] > 0 && (src[i] == result2[i] || result2[i] == result[i])))
            return "reads its own result";
        if (info[result[i]].is_input || (o == OP_LU && info[result2[i]].is_input)) return "writes an input block";
        return nullptr;
    };
#pragma omp parallel for schedule(static)
    for (int64_t id = 0; id < n_ids; id++) info[id].first_writer = NONE;
    {
        int64_t first_bad = n_ops;
#pragma omp parallel for schedule(static) reduction(min : first_bad)
        for (int64_t i = 0; i < n_ops; i++) {
            if (op_error(i)) { first_bad = std::min(first_bad, i); continue; }
            const int32_t rs[2] = {result[i], op[i] == OP_LU ? result2[i] : 0};
            for (int32_t r : rs) {
                if (r <= 0) continue;
                __atomic_fetch_add(&info[r].n_writers, 1, __ATOMIC_RELAXED);
                atomic_min(&info[r].first_writer, i);
            }
            if (src[i] > 0) __atomic_fetch_add(&info[src[i]].n_readers, 1, __ATOMIC_RELAXED);
            if (src2[i] > 0 && src2[i] != src[i]) __atomic_fetch_add(&info[src2[i]].n_readers, 1, __ATOMIC_RELAXED);
        }
        if (first_bad < n_ops) {
            const char* what = op_error(first_bad);
            if (std::string(what) == "unsupported op code") snprintf(msg, sizeof msg, "op %lld: unsupported op code %d", (long long)first_bad, (int)op[first_bad]);
            else if (std::string(what) == "writes an input block")
                snprintf(msg, sizeof msg, "op %lld writes input block %d", (long long)first_bad,
                         info[result[first_bad]].is_input ? result[first_bad] : result2[first_bad]);
            else if (std::string(what) == "lu without second result") snprintf(msg, sizeof msg, "lu without second result");
            else if (std::string(what) == "reads its own result") snprintf(msg, sizeof msg, "op %lld reads its own result", (long long)first_bad);
            else snprintf(msg, sizeof msg, "op %lld: %s", (long long)first_bad, what);
            return msg;
        }
    }
#pragma omp parallel for schedule(static)
    for (int64_t id = 0; id < n_ids; id++) {
        IdInfo& w = info[id];
        if (w.first_writer == NONE) w.first_writer = -1;
        else w.kind = op[w.first_writer];
    }
    {
        // several writers of one block must all be accumulating ops of one kind
        int64_t first_bad = n_ops;
#pragma omp parallel for schedule(static) reduction(min : first_bad)
        for (int64_t i = 0; i < n_ops; i++) {
            const int32_t rs[2] = {result[i], op[i] == OP_LU ? result2[i] : 0};
            for (int32_t r : rs)
                if (r > 0 && info[r].n_writers > 1 && (info[r].kind != op[i] || !is_acc(op[i]))) first_bad = std::min(first_bad, i);
        }
        if (first_bad < n_ops) {
            int32_t r = result[first_bad];
            if (!(info[r].n_writers > 1 && (info[r].kind != op[first_bad] || !is_acc(op[first_bad])))) r = result2[first_bad];
            snprintf(msg, sizeof msg, "block %d has writers of mixed or non-accumulating kinds", r);
            return msg;
        }
    }
    lap("validate + id info");
        return "";
    }

    // ==== which subs / inverses are folded into other tasks ===================================================
    std::string decide_fusions() {
    // ---- fusion decisions ------------------------------------------------------------
    fused_sub_of.assign(opt.fuse_sub ? n_ids : 0, -1);
    if (opt.fuse_sub) {
        int64_t nf = 0;
#pragma omp parallel for schedule(static) reduction(+ : nf)
        for (int64_t i = 0; i < n_ops; i++) {
            if (op[i] != OP_SUB) continue;
            int32_t p = src[i];
            if (p <= 0) continue;
            IdInfo& w = info[p];     // read by this sub only (n_readers == 1): no other thread touches it
            if (w.is_input || w.keep || w.n_writers == 0 || w.n_readers != 1) continue;
            if (!(w.kind == OP_MUL || w.kind == OP_MULNEG || w.kind == OP_MULT)) continue;
            if (src2[i] == p) continue;
            w.fused_away = true;
            fused_sub_of[p] = i;
            nf++;
        }
        G.fused_subs = nf;
    }

    // Inverses: (a) every further lowerInv / upperInv of a block that already has one is the
    // same computation on the same input -> its result id aliases the first one's block
    // (the planner re-derives bottom-right inverses at every recursion level, BlockPlanner.cpp:
    // 1035-1036, 1081-1082); (b) the first inverse of an lu factor is folded into the lu task.
    op_fused.assign(n_ops, 0);
    alias_to.assign(n_ids, 0);
    inv_of.assign(n_ids, NONE);
    {
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < n_ops; i++)
            if ((op[i] == OP_LOWERINV || op[i] == OP_UPPERINV) && src[i] > 0) atomic_min(&inv_of[src[i]], i);
#pragma omp parallel for schedule(static)
        for (int64_t id = 0; id < n_ids; id++)
            if (inv_of[id] == NONE) inv_of[id] = -1;
        int64_t n_alias = 0, n_fused = 0;
#pragma omp parallel for schedule(static) reduction(+ : n_alias, n_fused)
        for (int64_t i = 0; i < n_ops; i++) {
            if (op[i] != OP_LOWERINV && op[i] != OP_UPPERINV) continue;
            const int32_t f = src[i];
            if (f <= 0) continue;
            const int64_t first = inv_of[f];
            if (first != i) {
                if (op[first] == op[i] && opt.fuse_inv && !info[result[i]].keep) {
                    alias_to[result[i]] = result[first];
                    op_fused[i] = 2;
                    n_alias++;
                }
                continue;
            }
            if (!opt.fuse_inv) continue;
            const IdInfo& w = info[f];
            if (w.n_writers != 1 || (w.kind != OP_LU && w.kind != OP_LLT)) continue;
            const int64_t lu_i = w.first_writer;
            // lowerInv must read the L result, upperInv the U result of that lu; an llt only has L
            if (w.kind == OP_LLT ? (op[i] != OP_LOWERINV) : (op[i] == OP_LOWERINV ? (result[lu_i] != f) : (result2[lu_i] != f))) continue;
            op_fused[i] = 1;
            n_fused++;
        }
        G.aliased_invs = n_alias;
        G.fused_invs = n_fused;
    }
    lap("fusion decisions");
        return "";
    }

    // ==== one task per produced block =========================================================================
    std::string make_tasks() {
    // ---- one task per produced block (lu: one task, two blocks) -------------------------
    // task order = order of the first contributing op, i.e. the reference's stage order
    G.task_of.assign(n_ids, -1);
    G.slot_of.assign(n_ids, 0);
    auto opens_task = [&](int64_t i) {
        const IdInfo& w = info[result[i]];
        return w.first_writer == i && !w.fused_away && !op_fused[i];   // fused products are opened by their sub, folded inverses by their lu
    };
    {
        int nth = 1;
#ifdef _OPENMP
        nth = omp_get_max_threads();
#endif
        const int64_t chunk = (n_ops + nth - 1) / nth;
        std::vector<int64_t> first_task(nth + 1, 0);
#pragma omp parallel for schedule(static, 1) num_threads(nth)
        for (int c = 0; c < nth; c++) {
            int64_t k = 0;
            for (int64_t i = c * chunk, e = std::min(n_ops, i + chunk); i < e; i++)
                if (opens_task(i)) k++;
            first_task[c + 1] = k;
        }
        for (int c = 0; c < nth; c++) first_task[c + 1] += first_task[c];
        if (first_task[nth] > 0x7fffffff) return "too many tasks";
        G.tasks.resize(first_task[nth]);
#pragma omp parallel for schedule(static, 1) num_threads(nth)
        for (int c = 0; c < nth; c++) {
            int64_t tid = first_task[c];
            for (int64_t i = c * chunk, e = std::min(n_ops, i + chunk); i < e; i++) {
                if (!opens_task(i)) continue;
                const int32_t r = result[i];
                const IdInfo& w = info[r];
                Task t = {};
                t.out = r;                              // block ids for now; slots are patched below
                t.n_pairs = 1;
                switch (op[i]) {
                    case OP_LU:
                        t.type = T_LU;
                        t.out2 = result2[i];
                        if (inv_of[r] >= 0 && op_fused[inv_of[r]] == 1) { t.flags |= TF_LINV; t.init = result[inv_of[r]]; }
                        if (inv_of[result2[i]] >= 0 && op_fused[inv_of[result2[i]]] == 1) { t.flags |= TF_UINV; t.out4 = result[inv_of[result2[i]]]; }
                        break;
                    case OP_LLT:
                        t.type = T_LLT;
                        if (inv_of[r] >= 0 && op_fused[inv_of[r]] == 1) { t.flags |= TF_LINV; t.init = result[inv_of[r]]; }
                        break;
                    case OP_LOWERINV: t.type = T_LOWERINV; break;
                    case OP_UPPERINV: t.type = T_UPPERINV; break;
                    case OP_MUL: t.type = T_GEMM; t.n_pairs = w.n_writers; break;
                    case OP_MULNEG: t.type = T_GEMM; t.flags = TF_NEGATE; t.n_pairs = w.n_writers; break;
                    case OP_MULT: t.type = T_GEMM; t.flags = TF_TRANSB; t.n_pairs = w.n_writers; break;
                    case OP_SUB: {
                        int32_t p = src[i];
                        if (p > 0 && info[p].fused_away && fused_sub_of[p] == i) {
                            const IdInfo& pw = info[p];
                            t.type = T_GEMM;
                            t.n_pairs = pw.n_writers;
                            t.flags = (pw.kind == OP_MULNEG ? 0 : TF_NEGATE) | (pw.kind == OP_MULT ? TF_TRANSB : 0);
                            if (src2[i] > 0) { t.flags |= TF_INIT; t.init = src2[i]; }
                        } else {
                            t.type = T_SUB;
                        }
                        break;
                    }
                }
                G.task_of[r] = (int32_t)tid;
                if (t.type == T_LU) G.task_of[t.out2] = (int32_t)tid;
                if (t.type == T_LU || t.type == T_LLT) {
                    if (t.flags & TF_LINV) G.task_of[t.init] = (int32_t)tid;
                    if (t.flags & TF_UINV) G.task_of[t.out4] = (int32_t)tid;
                }
                G.tasks[tid++] = t;
            }
        }
    }
    nt = (int64_t)G.tasks.size();
    // redirect fused products to the task of their sub's result
    if (opt.fuse_sub) {
#pragma omp parallel for schedule(static)
        for (int64_t id = 1; id < n_ids; id++)
            if (info[id].fused_away) G.task_of[id] = G.task_of[result[fused_sub_of[id]]];
    }
    // (pool slots are assigned after the pairs are known: recycling needs every block's last reader)
#pragma omp parallel for schedule(static)
    for (int64_t id = 1; id < n_ids; id++)
        if (alias_to[id]) G.task_of[id] = G.task_of[alias_to[id]];
    lap("tasks");
        return "";
    }

    // ==== operand pairs of the accumulation chains, in op-list order ==========================================
    std::string fill_pairs() {
    // ---- pairs ------------------------------------------------------------------------------
    {
        int64_t total = 0;
        for (Task& t : G.tasks) { t.pair_begin = (int32_t)total; total += t.n_pairs; }
        if (total > 0x7fffffff) return "too many operand pairs";
        G.pairs.assign(total, Pair{0, 0});
        // the operands of an accumulation chain are summed in op-list order: threads claim positions atomically
        // and every chain is then put back into op order
        BigVec<int32_t> fill(nt, 0), pair_op(total);
        double flops = 0;
        int64_t gemm_pairs = 0;
#pragma omp parallel for schedule(static) reduction(+ : flops, gemm_pairs)
        for (int64_t i = 0; i < n_ops; i++) {
            const int32_t tid = G.task_of[result[i]];   // own task, or the task the op was folded into
            const uint8_t o = op[i];
            Task& t = G.tasks[tid];
            if (is_acc(o)) {
                const int32_t k = __atomic_fetch_add(&fill[tid], 1, __ATOMIC_RELAXED);
                if (k < t.n_pairs) { G.pairs[t.pair_begin + k] = Pair{src[i], src2[i]}; pair_op[t.pair_begin + k] = (int32_t)i; }
                flops += 524288.0;
                gemm_pairs++;
            } else if (o == OP_SUB) {
                flops += 4096.0;
                if (t.type == T_SUB) G.pairs[t.pair_begin] = Pair{src2[i], src[i]};   // a = S2, b = S1
            } else {
                if (!op_fused[i]) G.pairs[t.pair_begin] = Pair{src[i], 0};
                flops += (o == OP_LU) ? 174763.0 : 87381.0;
            }
        }
        G.flops = flops;           // sums of integers below 2^53: exact in any order
        G.n_gemm_pairs = gemm_pairs;
        int mismatch = 0;
#pragma omp parallel for schedule(dynamic, 4096) reduction(| : mismatch)
        for (int64_t t = 0; t < nt; t++) {
            const Task& T = G.tasks[t];
            if (T.type != T_GEMM) continue;
            const int32_t n = T.n_pairs, pb = T.pair_begin;
            if (fill[t] != n) { mismatch = 1; continue; }
            bool sorted = true;
            for (int32_t k = 1; k < n; k++) sorted = sorted && pair_op[pb + k - 1] < pair_op[pb + k];
            if (sorted) continue;
            // sort by op index (chains are short: up to a few hundred operands)
            std::vector<std::pair<int32_t, Pair>> tmp(n);
            for (int32_t k = 0; k < n; k++) tmp[k] = {pair_op[pb + k], G.pairs[pb + k]};
            std::sort(tmp.begin(), tmp.end(), [](const auto& x, const auto& y) { return x.first < y.first; });
            for (int32_t k = 0; k < n; k++) G.pairs[pb + k] = tmp[k].second;
        }
        if (mismatch) return "internal: pair count mismatch";
    }
    lap("pairs");
        return "";
    }

    // ==== multi-GPU: owners, mirrors of remote blocks and their fetch tasks ===================================
    std::string shard() {
    // ---- multi-GPU: owners, mirrors of remote blocks and their fetch tasks ---------------------------
    // A task runs on the GPU that owns its result block.  A produced block that a GPU reads at least
    // mirror_min times from a peer is MIRRORED there: a fetch task (a T_SUB "copy": out = remote - 0)
    // owned by the reader copies it once over NVLink as soon as it is complete, and the reader's
    // tasks use the copy (the panel broadcast of the north star, pulled by the consumer).  Mirrors
    // are ordinary blocks (ids appended after the caller's) so slot recycling and segments cover them.
    nid = n_ids;
    nown = std::max(1, opt.n_owners);
    G.n_owners = nown;
    G.owner_of.assign(n_ids, 0);
    if (nown > 1) {
        if (nown > MAX_GPUS) return "too many GPUs";
        if (opt.owner_of_id)
            for (int64_t id = 1; id < n_ids_caller; id++) {
                if (opt.owner_of_id[id] < 0 || opt.owner_of_id[id] >= nown) return "block owner out of range";
                G.owner_of[id] = opt.owner_of_id[id];
            }
        for (int64_t id = 1; id < n_ids; id++)
            if (alias_to[id]) G.owner_of[id] = G.owner_of[alias_to[id]];
        auto cn = [&](int32_t id) { return alias_to[id] ? alias_to[id] : id; };
        std::vector<int8_t> town(nt);
        for (int64_t t = 0; t < nt; t++) town[t] = G.owner_of[cn(G.tasks[t].out)];
        for (int64_t t = 0; t < nt; t++) {
            const Task& T = G.tasks[t];
            if ((T.type == T_LU && G.owner_of[T.out2] != town[t]) ||
                ((T.type == T_LU || T.type == T_LLT) && (((T.flags & TF_LINV) && G.owner_of[cn(T.init)] != town[t]) ||
                                                         ((T.flags & TF_UINV) && G.owner_of[cn(T.out4)] != town[t]))))
                return "lu outputs with different owners";
        }
        auto for_src = [&](Task& T, auto&& fn) {
            for (int32_t k = 0; k < T.n_pairs; k++) { fn(G.pairs[T.pair_begin + k].a); fn(G.pairs[T.pair_begin + k].b); }
            if (T.flags & TF_INIT) fn(T.init);
        };
        std::vector<uint8_t> reads((size_t)nown * n_ids, 0);
        for (int64_t t = 0; t < nt; t++) {
            const int r = town[t];
            for_src(G.tasks[t], [&](int32_t& sid) {
                if (sid <= 0) return;
                const int32_t c = cn(sid);
                if (G.owner_of[c] != r && G.task_of[c] >= 0) { uint8_t& n = reads[(size_t)r * n_ids + c]; if (n < 255) n++; }
            });
        }
        // fetch tasks, grouped behind the task that completes the block they copy
        std::vector<int32_t> mirror_id((size_t)nown * n_ids, 0);
        std::vector<std::vector<std::pair<int32_t, int8_t>>> fetch_after(nt);   // producer task -> (block id, reader)
        G.mirrors_per_owner.assign(nown, 0);
        for (int r = 0; r < nown; r++)
            for (int64_t id = 1; id < n_ids; id++)
                if (reads[(size_t)r * n_ids + id] >= opt.mirror_min) {
                    mirror_id[(size_t)r * n_ids + id] = (int32_t)nid++;
                    fetch_after[G.task_of[id]].push_back({(int32_t)id, (int8_t)r});
                    G.mirrors_per_owner[r]++;
                }
        std::vector<uint8_t>().swap(reads);
        if (nid > n_ids) {
            if (nid > 0x7fffffff) return "too many block ids";
            info.resize(nid);
            alias_to.resize(nid, 0);
            G.task_of.resize(nid, -1);
            G.slot_of.resize(nid, 0);
            G.owner_of.resize(nid, 0);
            // operands of every task: remote mirrored sources -> the reader's mirror id
            for (int64_t t = 0; t < nt; t++) {
                const int r = town[t];
                for_src(G.tasks[t], [&](int32_t& sid) {
                    if (sid <= 0) return;
                    const int32_t c = cn(sid);
                    if (c < n_ids && G.owner_of[c] != r) { const int32_t m = mirror_id[(size_t)r * n_ids + c]; if (m) sid = m; }
                });
            }
            // merged task order: every task followed by the fetch tasks of the blocks it completes
            BigVec<Task> merged;
            merged.reserve(nt + (nid - n_ids));
            std::vector<int32_t> new_index(nt);
            std::vector<int8_t> town2;
            town2.reserve(nt + (nid - n_ids));
            for (int64_t t = 0; t < nt; t++) {
                new_index[t] = (int32_t)merged.size();
                merged.push_back(G.tasks[t]);
                town2.push_back(town[t]);
                for (const auto& f : fetch_after[t]) {
                    const int32_t m = mirror_id[(size_t)f.second * n_ids + f.first];
                    Task F = {};
                    F.type = T_SUB;                      // out = a - b with b = zero block: a copy
                    F.n_pairs = 1;
                    F.pair_begin = (int32_t)G.pairs.size();
                    G.pairs.push_back(Pair{f.first, 0});
                    F.out = m;
                    IdInfo& w = info[m];
                    w.n_writers = 1; w.kind = OP_SUB; w.first_writer = -1;
                    G.owner_of[m] = f.second;
                    G.task_of[m] = (int32_t)merged.size();
                    merged.push_back(F);
                    town2.push_back(f.second);
                }
            }
            for (int64_t id = 1; id < n_ids; id++)
                if (G.task_of[id] >= 0) G.task_of[id] = new_index[G.task_of[id]];
            G.tasks.swap(merged);
            town.swap(town2);
            nt = (int64_t)G.tasks.size();
        }
        G.task_owner = town;
    } else {
        G.task_owner.assign(nt, 0);
    }

    lap("owners + mirrors");
        return "";
    }

    // ==== pool slots, segments and slot recycling =============================================================
    std::string assign_slots() {
    // ---- pool slots and segments -------------------------------------------------------------------
    // Unlimited pool: every input / produced block gets its own slot, one segment.  Limited pool
    // (opt.max_slots): walk the tasks in order (a topological order: the op list is stage-sorted),
    // hand out free slots, and when none is left close the segment there; at a segment boundary
    // every block whose producer and readers all lie before it is dead and its slot returns to the
    // free list.  Inputs, kept blocks (L, U) and the diagonal inverses the solve uses are pinned.
    auto canon = [&](int32_t id) { return alias_to[id] ? alias_to[id] : id; };
    seg_of.assign(nt, 0);
    G.recycled.assign(nid, 0);
    reused_slot.assign(nid, 0);
    G.seg_begin.assign(1, 0);
    G.slots_per_owner.assign(nown, 1);          // local slot 0 of every GPU is its all-zero block
    {
        auto needs_slot = [&](int64_t id) { const IdInfo& w = info[id]; return w.is_input || (w.n_writers > 0 && !w.fused_away && !alias_to[id]); };
        std::vector<int64_t> need(nown, 1), n_in(nown, 0);
        for (int64_t id = 1; id < nid; id++) { if (needs_slot(id)) need[G.owner_of[id]]++; if (info[id].is_input) n_in[G.owner_of[id]]++; }
        bool fits = true;
        for (int o = 0; o < nown; o++) fits = fits && (opt.max_slots <= 0 || need[o] <= opt.max_slots);
        if (fits) {
            for (int64_t id = 1; id < nid; id++)
                if (needs_slot(id)) G.slot_of[id] = (int32_t)G.slots_per_owner[G.owner_of[id]]++;
        } else {
            for (int o = 0; o < nown; o++)
                if (opt.max_slots <= n_in[o] + 8) return "block pool too small even for the input blocks";
            // last task that touches each block (producer or reader)
            std::vector<int32_t> last_touch(nid, -1);
            std::vector<char> pinned(nid, 0);
            for (int64_t id = 1; id < nid; id++) pinned[id] = info[id].is_input || info[id].keep;
            for (int64_t t = 0; t < nt; t++) {
                const Task& T = G.tasks[t];
                auto touch = [&](int32_t id) { if (id > 0) { id = canon(id); if (last_touch[id] < t) last_touch[id] = (int32_t)t; } };
                touch(T.out);
                if (T.type == T_LU) touch(T.out2);
                if (T.type == T_LU || T.type == T_LLT) {
                    if (T.flags & TF_LINV) { touch(T.init); pinned[canon(T.init)] = 1; }
                    if (T.flags & TF_UINV) { touch(T.out4); pinned[canon(T.out4)] = 1; }
                }
                if (T.flags & TF_INIT) touch(T.init);
                for (int32_t k = 0; k < T.n_pairs; k++) { touch(G.pairs[T.pair_begin + k].a); touch(G.pairs[T.pair_begin + k].b); }
            }
            // blocks ordered by last touch, for the release sweep at boundaries
            std::vector<int32_t> by_touch;
            for (int64_t id = 1; id < nid; id++)
                if (needs_slot(id) && !pinned[id]) by_touch.push_back((int32_t)id);
            std::sort(by_touch.begin(), by_touch.end(), [&](int32_t x, int32_t y) { return last_touch[x] < last_touch[y]; });
            size_t rel = 0;
            std::vector<std::vector<int32_t>> freelist(nown);
            for (int64_t id = 1; id < nid; id++)
                if (info[id].is_input) G.slot_of[id] = (int32_t)G.slots_per_owner[G.owner_of[id]]++;
            bool took_fresh = false;
            auto take = [&](int o) -> int32_t {
                took_fresh = false;
                if (!freelist[o].empty()) { int32_t sl = freelist[o].back(); freelist[o].pop_back(); return sl; }
                if (G.slots_per_owner[o] < opt.max_slots) { took_fresh = true; return (int32_t)G.slots_per_owner[o]++; }
                return -1;
            };
            int32_t cur_seg = 0;
            for (int64_t t = 0; t < nt; t++) {
                const Task& T = G.tasks[t];
                int32_t outs[4] = {T.out, 0, 0, 0};
                int no = 1;
                if (T.type == T_LU) outs[no++] = T.out2;
                if (T.type == T_LU || T.type == T_LLT) {
                    if (T.flags & TF_LINV) outs[no++] = T.init;
                    if (T.flags & TF_UINV) outs[no++] = T.out4;
                }
                for (int k = 0; k < no; k++) {
                    const int32_t id = outs[k];
                    if (alias_to[id] || G.slot_of[id] > 0) continue;
                    const int o = G.owner_of[id];
                    int32_t sl = take(o);
                    if (sl < 0) {
                        // close the segment before task t and release everything dead by then (all GPUs)
                        if (G.seg_begin.back() == (int32_t)t) return "block pool too small: a single task's live set does not fit";
                        G.seg_begin.push_back((int32_t)t);
                        cur_seg++;
                        while (rel < by_touch.size() && last_touch[by_touch[rel]] < t) {
                            const int32_t dead = by_touch[rel++];
                            if (G.slot_of[dead] > 0) { freelist[G.owner_of[dead]].push_back(G.slot_of[dead]); G.recycled[dead] = 1; }
                        }
                        sl = take(o);
                        if (sl < 0) return "block pool too small for the live set of the factorisation";
                    }
                    G.slot_of[id] = sl;
                    if (!took_fresh || !pinned[id]) reused_slot[id] = 1;    // the slot has held, or will hold, another block
                }
                seg_of[t] = cur_seg;
            }
        }
        for (int64_t id = 1; id < nid; id++)
            if (alias_to[id]) { G.slot_of[id] = G.slot_of[alias_to[id]]; G.recycled[id] = G.recycled[alias_to[id]]; }
        G.n_slots = G.slots_per_owner[0];
        G.seg_begin.push_back((int32_t)nt);
    }

    lap("slots + segments");
        return "";
    }

    // ==== producer -> consumer edges, successor lists =========================================================
    std::string dependencies() {
    // ---- dependencies: distinct producer tasks of every source, inside the same segment ----------
    // (producers in earlier segments have finished before the launch starts)
    // predecessor lists in one flat array (capacity = operand count per task), sorted and made unique per task
    poff.assign(nt + 1, 0);
    for (int64_t t = 0; t < nt; t++) poff[t + 1] = poff[t] + 2 * (int64_t)G.tasks[t].n_pairs + ((G.tasks[t].flags & TF_INIT) ? 1 : 0);
    pflat = BigVec<int32_t>(poff[nt]);
    pcnt.assign(nt, 0);
    {
        int order_error = 0;
#pragma omp parallel for schedule(dynamic, 2048) reduction(| : order_error)
        for (int64_t t = 0; t < nt; t++) {
            const Task& T = G.tasks[t];
            int32_t* v = pflat.data() + poff[t];
            int32_t n = 0;
            auto add = [&](int32_t id) {
                if (id <= 0) return;
                const int32_t p = G.task_of[id];
                if (p > t) order_error = 1;   // the op list is not topologically ordered
                if (p >= 0 && p != t && seg_of[p] == seg_of[t]) v[n++] = p;
            };
            for (int32_t k = 0; k < T.n_pairs; k++) {
                add(G.pairs[T.pair_begin + k].a);
                add(G.pairs[T.pair_begin + k].b);
            }
            if (T.flags & TF_INIT) add(T.init);
            if (n > 1) {
                std::sort(v, v + n);
                n = (int32_t)(std::unique(v, v + n) - v);
            }
            pcnt[t] = n;
        }
        lap("  deps: pred lists");
        if (order_error) return "operation list is not in dependency order (a block is read before its producer's first op)";
        int64_t nsucc = 0;
        for (int64_t t = 0; t < nt; t++) { G.tasks[t].n_deps = pcnt[t]; nsucc += pcnt[t]; }
        if (nsucc > 0x7fffffff) return "too many dependency edges";
        // transpose (predecessor lists -> successor lists): positions are claimed atomically, then every list is
        // sorted, which restores the ascending task order a serial pass would give
        BigVec<int32_t> cnt(nt + 1, 0);
#pragma omp parallel for schedule(static)
        for (int64_t t = 0; t < nt; t++) {
            const int32_t* v = pflat.data() + poff[t];
            for (int32_t k = 0; k < pcnt[t]; k++) __atomic_fetch_add(&cnt[v[k] + 1], 1, __ATOMIC_RELAXED);
        }
        for (int64_t t = 0; t < nt; t++) cnt[t + 1] += cnt[t];
        lap("  deps: count");
        G.succ.assign(nsucc, 0);
        {
            BigVec<int32_t> pos(cnt.begin(), cnt.end() - 1);
#pragma omp parallel for schedule(static)
            for (int64_t t = 0; t < nt; t++) {
                const int32_t* v = pflat.data() + poff[t];
                for (int32_t k = 0; k < pcnt[t]; k++) G.succ[__atomic_fetch_add(&pos[v[k]], 1, __ATOMIC_RELAXED)] = (int32_t)t;
            }
        }
        lap("  deps: fill");
#pragma omp parallel for schedule(dynamic, 4096)
        for (int64_t t = 0; t < nt; t++)
            if (cnt[t + 1] - cnt[t] > 1) std::sort(G.succ.begin() + cnt[t], G.succ.begin() + cnt[t + 1]);
        for (int64_t t = 0; t < nt; t++) { G.tasks[t].succ_begin = cnt[t]; G.tasks[t].succ_end = cnt[t + 1]; }
    }
    lap("dependencies");
        return "";
    }

    // ==== ASAP levels =========================================================================================
    std::string levels() {
    // ---- levels: longest path from a source (every predecessor precedes its task, checked above) -----------
    {
        int32_t maxlev = 0;
        for (int64_t t = 0; t < nt; t++) {
            const int32_t* v = pflat.data() + poff[t];
            int32_t lv = 0;
            for (int32_t k = 0; k < pcnt[t]; k++) lv = std::max(lv, G.tasks[v[k]].level + 1);
            G.tasks[t].level = lv;
            maxlev = std::max(maxlev, lv);
        }
        // levels restart in every segment: make them globally increasing (segment order)
        const int nseg = (int)G.seg_begin.size() - 1;
        if (nseg > 1) {
            std::vector<int32_t> segmax(nseg, 0), off(nseg + 1, 0);
            for (int64_t t = 0; t < nt; t++) segmax[seg_of[t]] = std::max(segmax[seg_of[t]], G.tasks[t].level);
            for (int sg = 0; sg < nseg; sg++) off[sg + 1] = off[sg] + segmax[sg] + 1;
            for (int64_t t = 0; t < nt; t++) G.tasks[t].level += off[seg_of[t]];
            maxlev = off[nseg] - 1;
        }
        G.n_levels = nt ? maxlev + 1 : 0;
    }
    lap("levels");
        return "";
    }

    // ==== row slices of GEMM tasks in narrow levels / near the critical path ==================================
    std::string split_rows() {
    // ---- row split of GEMM tasks in narrow levels -------------------------------------------------
    // In a level with fewer GEMM tasks than SMs the factorisation is latency-bound: one 64x64x64
    // product occupies one SM for ~2.4 us per pair while the others idle.  Such tasks are split
    // into 2 or 4 row slices (each slice loads its rows of A and all of B); consumers depend on
    // every slice.  Wide levels stay whole (no extra operand traffic where throughput matters).
    for (Task& t : G.tasks)
        if (t.type == T_GEMM) t.flags |= (4 << TF_NROWS_SHIFT);
    if (opt.split_narrow && nt > 0) {
        std::vector<int32_t> level_gemms(G.n_levels, 0);
        for (const Task& t : G.tasks)
            if (t.type == T_GEMM) level_gemms[t.level]++;
        std::vector<int32_t> split(nt, 1), base(nt + 1, 0);
        for (int64_t t = 0; t < nt; t++) {
            const Task& T = G.tasks[t];
            if (T.type == T_GEMM) {
                const int w = level_gemms[T.level];
                if (4 * w <= opt.n_sms) split[t] = 4;
                else if (2 * w <= opt.n_sms) split[t] = 2;
            }
        }
        if (opt.split_slack_us > 0) {
            // A task on (or near) the longest dependent chain delays everything behind it even when its level is wide:
            // split those too.  Slack = longest chain - longest chain through the task, under the cost model of
            // cost_model.h with the width-based slices chosen above.
            const ModelParams M;
            std::vector<float> dur(nt), tl(nt, 0.f), bl(nt, 0.f);
            for (int64_t t = 0; t < nt; t++) dur[t] = (float)model_hop_us(G.tasks[t], M, 4 / split[t]);
            for (int64_t t = 0; t < nt; t++)
                for (int32_t e = G.tasks[t].succ_begin; e < G.tasks[t].succ_end; e++) tl[G.succ[e]] = std::max(tl[G.succ[e]], tl[t] + dur[t]);
            float cp = 0.f;
            for (int64_t t = nt - 1; t >= 0; t--) {
                float m = 0.f;
                for (int32_t e = G.tasks[t].succ_begin; e < G.tasks[t].succ_end; e++) m = std::max(m, bl[G.succ[e]]);
                bl[t] = dur[t] + m;
                cp = std::max(cp, tl[t] + bl[t]);
            }
            for (int64_t t = 0; t < nt; t++)
                if (G.tasks[t].type == T_GEMM && split[t] < 4 && cp - (tl[t] + bl[t]) < (float)opt.split_slack_us) split[t] = 4;
        }
        for (int64_t t = 0; t < nt; t++) {
            base[t + 1] = base[t] + split[t];
            if (split[t] > 1) G.split_tasks++;
        }
        if (G.split_tasks > 0) {
            // The slices of one task form a group with ONE dependency counter (the leader's): a finishing slice
            // decrements each successor group once, so a group waits for every slice of every predecessor, and the
            // slices share their task's successor list (no edge multiplication).
            const int64_t nt2 = base[nt];
            BigVec<Task> tasks2(nt2);
#pragma omp parallel for schedule(static)
            for (int64_t t = 0; t < nt; t++) {
                const Task& T = G.tasks[t];
                int32_t nd = 0;
                const int32_t* v = pflat.data() + poff[t];
                for (int32_t k = 0; k < pcnt[t]; k++) nd += split[v[k]];
                for (int s = 0; s < split[t]; s++) {
                    Task N = T;
                    if (split[t] > 1) {
                        const int rows16 = 4 / split[t];
                        N.flags = (T.flags & 0xff) | ((s * rows16) << TF_ROW0_SHIFT) | (rows16 << TF_NROWS_SHIFT);
                    }
                    N.n_deps = nd;
                    tasks2[base[t] + s] = N;
                }
            }
#pragma omp parallel for schedule(static)
            for (int64_t e = 0; e < (int64_t)G.succ.size(); e++) G.succ[e] = base[G.succ[e]];
            {
                std::vector<int8_t> own2(nt2);
#pragma omp parallel for schedule(static)
                for (int64_t t = 0; t < nt; t++)
                    for (int q = 0; q < split[t]; q++) own2[base[t] + q] = G.task_owner[t];
                G.task_owner.swap(own2);
            }
            G.tasks.swap(tasks2);
#pragma omp parallel for schedule(static)
            for (int64_t id = 0; id < (int64_t)G.task_of.size(); id++)
                if (G.task_of[id] >= 0) G.task_of[id] = base[G.task_of[id]];
            for (int32_t& b : G.seg_begin) b = base[b];
        }
    }
    lap("row split");
        return "";
    }

    // ==== static execution order: most urgent first ===========================================================
    std::string static_order() {
    // The executor claims tasks IN TASK ORDER (position s of a segment = its s-th task) and waits on the claimed task's
    // dependency counter, so the order is the schedule's priority list.  Tasks are sorted by a blend of their latest
    // start time under the cost model (longest chain of the segment minus the longest chain from the task to a sink:
    // work on the critical chain first, bulk work ordered by when the chain will need it) and their earliest start
    // time (longest chain from a source).  Pure latest-start order is "just in time": a task that slips -- the model
    // assumes a free SM for every task -- then delays the critical chain; the blend pulls work with slack forward
    // (measured on a B200, alpha = 1 / 0.7 / 0.4 / 0: 64^3 179 / 165 / 164 / 190 ms, banded 135 / 124 / 121 / 145 ms).
    // Both keys grow strictly along every edge (by the predecessor's duration), so any blend is a topological order:
    // the lowest unfinished task of a run is always claimed and ready -- no deadlock, on one GPU or on several (each
    // GPU keeps its tasks in this global order).
    if (!opt.static_order || G.tasks.empty()) return "";
    const int64_t n = (int64_t)G.tasks.size();
    const ModelParams M;
    std::vector<double> bl(n, 0.0), key(n, 0.0);
    std::vector<int32_t> seg(n, 0);
    const int nseg = (int)G.seg_begin.size() - 1;
    for (int sg = 0; sg < nseg; sg++)
        for (int32_t t = G.seg_begin[sg]; t < G.seg_begin[sg + 1]; t++) seg[t] = sg;
    // longest chain from the start of a task to a sink of its segment; successor lists name group leaders, the
    // slices of one task share their leader's list
    std::vector<double> cp(std::max(nseg, 1), 0.0);
    for (int64_t t = n - 1; t >= 0; t--) {
        const Task& T = G.tasks[t];
        double m = 0.0;
        for (int32_t e = T.succ_begin; e < T.succ_end; e++)
            if (seg[G.succ[e]] == seg[t]) m = std::max(m, bl[G.succ[e]]);
        bl[t] = model_hop_us(T, M) + m;
    }
    // a group's urgency is its most urgent slice's; groups stay contiguous, leader first
    std::vector<int32_t> leaders;
    leaders.reserve(n);
    for (int64_t t = 0; t < n; t++) {
        if (!task_is_leader(G.tasks[t])) continue;
        double b = 0.0;
        for (int q = 0, g = task_group_size(G.tasks[t]); q < g; q++) b = std::max(b, bl[t + q]);
        bl[t] = b;
        cp[seg[t]] = std::max(cp[seg[t]], b);
        leaders.push_back((int32_t)t);
    }
    for (int32_t t : leaders) key[t] = cp[seg[t]] - bl[t];
    if (opt.order_alpha < 1.0) {
        // earliest start (longest chain from a source of the segment), per group
        std::vector<double> tl(n, 0.0);
        for (int64_t t = 0; t < n; t++) {
            const Task& T = G.tasks[t];
            const int32_t ld = (int32_t)t - (task_is_leader(T) ? 0 : ((T.flags >> TF_ROW0_SHIFT) & 3) / ((T.flags >> TF_NROWS_SHIFT) & 7));
            const double fin = tl[ld] + model_hop_us(T, M);
            for (int32_t e = T.succ_begin; e < T.succ_end; e++)
                if (seg[G.succ[e]] == seg[t]) tl[G.succ[e]] = std::max(tl[G.succ[e]], fin);
        }
        for (int32_t t : leaders) key[t] = opt.order_alpha * key[t] + (1.0 - opt.order_alpha) * tl[t];
    }
    std::stable_sort(leaders.begin(), leaders.end(), [&](int32_t a, int32_t b) {
        if (seg[a] != seg[b]) return seg[a] < seg[b];
        return key[a] < key[b];
    });
    std::vector<int32_t> new_of(n);
    {
        int32_t pos = 0;
        for (int32_t t : leaders)
            for (int q = 0, g = task_group_size(G.tasks[t]); q < g; q++) new_of[t + q] = pos++;
    }
    // a predecessor's latest start is at least its own duration earlier: the order must have stayed topological
    for (int64_t t = 0; t < n; t++)
        for (int32_t e = G.tasks[t].succ_begin; e < G.tasks[t].succ_end; e++)
            if (new_of[G.succ[e]] <= new_of[t]) return "static order: a successor precedes its predecessor";
    BigVec<Task> tasks2(n);
    std::vector<int8_t> own2(n);
#pragma omp parallel for schedule(static)
    for (int64_t t = 0; t < n; t++) { tasks2[new_of[t]] = G.tasks[t]; own2[new_of[t]] = G.task_owner[t]; }
    G.tasks.swap(tasks2);
    G.task_owner.swap(own2);
#pragma omp parallel for schedule(static)
    for (int64_t e = 0; e < (int64_t)G.succ.size(); e++) G.succ[e] = new_of[G.succ[e]];
    // successor lists most urgent first (the slices of a group share one list: the leader sorts it)
#pragma omp parallel for schedule(dynamic, 4096)
    for (int64_t t = 0; t < n; t++)
        if (task_is_leader(G.tasks[t]) && G.tasks[t].succ_end - G.tasks[t].succ_begin > 1)
            std::sort(G.succ.begin() + G.tasks[t].succ_begin, G.succ.begin() + G.tasks[t].succ_end);
#pragma omp parallel for schedule(static)
    for (int64_t id = 0; id < (int64_t)G.task_of.size(); id++)
        if (G.task_of[id] >= 0) G.task_of[id] = new_of[G.task_of[id]];
    lap("static order");
        return "";
    }

    // ==== successor references, initially ready tasks, block ids -> block references ==========================
    std::string finish() {
    // successor references (single-GPU upload)
    G.succ_enc.resize(G.succ.size());
#pragma omp parallel for schedule(static)
    for (int64_t t = 0; t < (int64_t)G.tasks.size(); t++) {
        const Task& T = G.tasks[t];
        if (!task_is_leader(T)) continue;              // the slices of one task share their leader's list
        for (int32_t e = T.succ_begin; e < T.succ_end; e++) G.succ_enc[e] = make_task_ref(0, task_log2_slices(G.tasks[G.succ[e]]), G.succ[e]);
    }
    if ((int64_t)G.tasks.size() > TASK_LOCAL_MASK) return "too many tasks";

    // per-segment lists of initially ready tasks (single GPU: every task is owned by GPU 0; the multi-GPU split is
    // redone per rank in localize_tasks)
    {
        const int nseg = (int)G.seg_begin.size() - 1;
        G.initial.clear();
        G.seg_init.assign(1, 0);
        for (int sg = 0; sg < nseg; sg++) {
            for (int32_t t = G.seg_begin[sg]; t < G.seg_begin[sg + 1]; t++)
                if (G.tasks[t].n_deps == 0) G.initial.push_back(t);
            G.seg_init.push_back((int32_t)G.initial.size());
        }
    }
    lap("successor refs + initial");
    // ---- patch block ids -> block references (owner in the top bits; plain slots on one GPU) --------
    {
        auto ref = [&](int32_t id, int reader) -> int32_t {
            if (id <= 0 || G.slot_of[id] <= 0) return make_ref(reader, 0);     // the reader's own zero block
            return make_ref(G.owner_of[id], G.slot_of[id]);
        };
        // pairs are shared by the row slices of one task: the leading slice patches them
#pragma omp parallel for schedule(static)
        for (int64_t t = 0; t < (int64_t)G.tasks.size(); t++) {
            Task& T = G.tasks[t];
            const int o = G.task_owner[t];
            if (T.type == T_LU || T.type == T_LLT) {
                bool clean = !reused_slot[T.out];
                if (T.type == T_LU) clean = clean && !reused_slot[T.out2];
                if (T.flags & TF_LINV) clean = clean && !reused_slot[T.init];
                if (T.flags & TF_UINV) clean = clean && !reused_slot[T.out4];
                if (clean) T.flags |= TF_TRI_OUT;
            }
            T.out = ref(T.out, o);
            if (T.type == T_LU) T.out2 = ref(T.out2, o);
            if (T.flags & (TF_INIT | TF_LINV)) T.init = ref(T.init, o);
            if (T.flags & TF_UINV) T.out4 = ref(T.out4, o);
            if (!task_is_leader(T)) continue;
            for (int32_t k = 0; k < T.n_pairs; k++) {
                const size_t q = (size_t)T.pair_begin + k;
                G.pairs[q].a = ref(G.pairs[q].a, o);
                G.pairs[q].b = ref(G.pairs[q].b, o);
            }
        }
#pragma omp parallel for schedule(static)
        for (int64_t t = 0; t < (int64_t)G.tasks.size(); t++) {
            Task& T = G.tasks[t];
            for (int k = 0; k < 2; k++) T.first[k] = (k < T.n_pairs) ? G.pairs[T.pair_begin + k] : Pair{0, 0};
        }
    }
    lap("patch refs");
        return "";
    }

};
}  // namespace

std::string compile_tasks(int64_t n_ids_caller, int64_t n_input, const int32_t* input_ids, int64_t n_ops,
                          const int32_t* src, const int32_t* src2, const uint8_t* op, const int32_t* result,
                          const int32_t* result2, const std::vector<int32_t>& keep_ids, const CompileOptions& opt,
                          TaskGraph& G) {
    return Compiler(n_ids_caller, n_input, input_ids, n_ops, src, src2, op, result, result2, keep_ids, opt, G).run();
}



std::string localize_tasks(const TaskGraph& G, int rank, DistLayout& D) {
    // The graph was compiled with owners: tasks, block references and mirrors are already per GPU.
    // A rank keeps its tasks (renumbered in order) and rewrites successor ids to (owner, local).
    D = DistLayout();
    D.rank = rank; D.world = G.n_owners;
    if (rank < 0 || rank >= G.n_owners) return "bad rank";
    const int64_t nt = (int64_t)G.tasks.size();
    D.task_owner = G.task_owner;
    D.task_local.resize(nt);
    D.tasks_per_rank.assign(G.n_owners, 0);
    for (int64_t t = 0; t < nt; t++) D.task_local[t] = (int32_t)D.tasks_per_rank[G.task_owner[t]]++;
    D.slots_per_rank = G.slots_per_owner;
    D.mirrored = G.mirrors_per_owner.empty() ? 0 : G.mirrors_per_owner[rank];
    const int nseg = (int)G.seg_begin.size() - 1;
    D.seg_begin.assign(1, 0);
    D.seg_init.assign(1, 0);
    D.seg_begin_all.assign(G.n_owners, std::vector<int32_t>(nseg + 1, 0));
    for (int sg = 0; sg < nseg; sg++) {
        for (int o = 0; o < G.n_owners; o++) D.seg_begin_all[o][sg + 1] = D.seg_begin_all[o][sg];
        for (int32_t t = G.seg_begin[sg]; t < G.seg_begin[sg + 1]; t++) D.seg_begin_all[G.task_owner[t]][sg + 1]++;
    }
    int32_t shared_from = -1, shared_begin = 0, shared_end = 0;
    for (int sg = 0; sg < nseg; sg++) {
        for (int32_t t = G.seg_begin[sg]; t < G.seg_begin[sg + 1]; t++) {
            if (G.task_owner[t] != rank) continue;
            Task T = G.tasks[t];
            const int32_t pb = (int32_t)D.pairs.size();
            for (int32_t k = 0; k < T.n_pairs; k++) {
                const Pair& p = G.pairs[T.pair_begin + k];
                D.remote_operands += ((int)((uint32_t)p.a >> REF_SHIFT) != rank) + ((int)((uint32_t)p.b >> REF_SHIFT) != rank);
                D.pairs.push_back(p);
            }
            T.pair_begin = pb;
            if (T.succ_begin != shared_from || T.succ_end - T.succ_begin != shared_end - shared_begin) {   // slices share one list
                shared_from = T.succ_begin;
                shared_begin = (int32_t)D.succ.size();
                for (int32_t e = T.succ_begin; e < T.succ_end; e++) {
                    const int32_t s2 = G.succ[e];
                    D.succ.push_back(make_task_ref(G.task_owner[s2], task_log2_slices(G.tasks[s2]), D.task_local[s2]));
                }
                shared_end = (int32_t)D.succ.size();
            }
            for (int32_t e = T.succ_begin; e < T.succ_end; e++) D.remote_edges += G.task_owner[G.succ[e]] != rank;
            T.succ_begin = shared_begin;
            T.succ_end = shared_end;
            if (T.n_deps == 0) D.initial.push_back((int32_t)D.tasks.size());
            D.tasks.push_back(T);
        }
        D.seg_init.push_back((int32_t)D.initial.size());
        D.seg_begin.push_back((int32_t)D.tasks.size());
    }
    return "";
}

}  // namespace soglu
