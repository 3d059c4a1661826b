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

std::string compile_tasks(int64_t n_ids_caller, int64_t n_input, const int32_t* input_ids, int64_t n_ops,
                          const int32_t* src, const int32_t* src2, const uint8_t* op, const int32_t* result,
                          const int32_t* result2, const std::vector<int32_t>& keep_ids, const CompileOptions& opt,
                          TaskGraph& G) {
    G = TaskGraph();
    char msg[256];
    // SOGLU_TIMING=1 prints the phase times to stderr (diagnostics only)
    const bool timing = std::getenv("SOGLU_TIMING") != nullptr;
    auto t_last = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (!timing) return;
        auto t1 = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[compile] %-25s %8.3f s\n", what, std::chrono::duration<double>(t1 - t_last).count());
        t_last = t1;
    };
    if (n_ids_caller < 1) return "n_block_ids must be >= 1";
    // Chain cuts (CompileOptions::chain_cuts): the early part of a cut accumulation chain writes a temporary block;
    // the temporaries get the ids n_ids_caller .. n_ids-1 and are ordinary blocks for everything below.
    const int64_t n_cuts = opt.chain_cuts ? (int64_t)opt.chain_cuts->size() : 0;
    const int64_t n_ids = n_ids_caller + n_cuts;
    if (n_ids > 0x7fffffff) return "too many block ids";
    if (n_ops > 0x7fffffff) return "too many operations";
    std::vector<IdInfo> info(n_ids);
    for (int64_t k = 0; k < n_input; k++) {
        int32_t id = input_ids[k];
        if (id <= 0 || id >= n_ids_caller) return "input block id out of range";
        info[id].is_input = true;
    }
    for (int32_t id : keep_ids)
        if (id > 0 && id < n_ids_caller) info[id].keep = true;

    // ---- validation + per-block writer / reader statistics --------------------------------------
    // All passes over the op list run on every host thread; where the serial order matters (first writer
    // of a block, first inverse of a factor, the lowest offending op) it is recovered with atomic minima.
    auto bad_id = [&](int32_t id) { return id < 0 || id >= n_ids_caller; };
    auto is_acc = [](uint8_t o) { return o == OP_MUL || o == OP_MULNEG || o == OP_MULT; };
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
    auto atomic_min = [](int64_t* p, int64_t v) {
        int64_t cur = __atomic_load_n(p, __ATOMIC_RELAXED);
        while (v < cur && !__atomic_compare_exchange_n(p, &cur, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    };
    const int64_t NONE = INT64_MAX;
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

    // ---- fusion decisions ------------------------------------------------------------
    std::vector<int64_t> fused_sub_of(opt.fuse_sub ? n_ids : 0, -1);  // product id -> index of the sub op
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
    std::vector<char> op_fused(n_ops, 0);                 // 1 = folded into an lu task, 2 = alias
    std::vector<int32_t> alias_to(n_ids, 0);              // result id -> canonical result id
    std::vector<int64_t> inv_of(n_ids, NONE);             // source id -> first inverse op
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

    // ---- one task per produced block (lu: one task, two blocks) -------------------------
    // task order = order of the first contributing op, i.e. the reference's stage order
    G.task_of.assign(n_ids, -1);
    G.slot_of.assign(n_ids, 0);
    auto opens_task = [&](int64_t i) {
        const IdInfo& w = info[result[i]];
        return w.first_writer == i && !w.fused_away && !op_fused[i];   // fused products are opened by their sub, folded inverses by their lu
    };
    // chain cuts by the block id the (uncut) task produces; a cut is applied if it still fits the chain
    std::vector<int32_t> cut_of(n_cuts ? n_ids : 0, -1);
    for (int64_t c = 0; c < n_cuts; c++) {
        const ChainCut& cc = (*opt.chain_cuts)[c];
        if (cc.out_id > 0 && cc.out_id < n_ids_caller && cut_of[cc.out_id] < 0) cut_of[cc.out_id] = (int32_t)c;
    }
    // number of operand pairs of the task op i opens, if it is a GEMM task (0 otherwise)
    auto chain_len = [&](int64_t i) -> int32_t {
        const uint8_t o = op[i];
        if (is_acc(o)) return info[result[i]].n_writers;
        if (o == OP_SUB && src[i] > 0 && info[src[i]].fused_away && fused_sub_of[src[i]] == i) return info[src[i]].n_writers;
        return 0;
    };
    auto cut_for = [&](int64_t i) -> int32_t {       // index of the applicable cut of the task op i opens, or -1
        if (!n_cuts) return -1;
        const int32_t c = cut_of[result[i]];
        if (c < 0) return -1;
        const ChainCut& cc = (*opt.chain_cuts)[c];
        const int32_t n = chain_len(i);
        if (cc.n_early <= 0 || cc.n_early >= n || (int32_t)cc.early_pos.size() != cc.n_early || cc.early_pos.back() >= n) return -1;
        return c;
    };
    std::vector<int32_t> cut_of_task;                // final task of a cut chain -> cut index (its early task precedes it)
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
                if (opens_task(i)) k += (cut_for(i) >= 0) ? 2 : 1;
            first_task[c + 1] = k;
        }
        for (int c = 0; c < nth; c++) first_task[c + 1] += first_task[c];
        if (first_task[nth] > 0x7fffffff) return "too many tasks";
        G.tasks.resize(first_task[nth]);
        if (n_cuts) cut_of_task.assign(first_task[nth], -1);
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
                const int32_t cut = (t.type == T_GEMM) ? cut_for(i) : -1;
                if (cut >= 0) {
                    // early part: tmp = init -/+ sum over the early pairs; the final task starts from tmp
                    const ChainCut& cc = (*opt.chain_cuts)[cut];
                    const int32_t tmp = (int32_t)(n_ids_caller + cut);
                    Task e1 = t;
                    e1.out = tmp;
                    e1.n_pairs = cc.n_early;
                    IdInfo& tw = info[tmp];
                    tw.n_writers = 1; tw.n_readers = 1; tw.kind = OP_MUL; tw.first_writer = i;
                    G.task_of[tmp] = (int32_t)tid;
                    G.tasks[tid++] = e1;
                    t.n_pairs -= cc.n_early;
                    t.flags = (t.flags & (TF_NEGATE | TF_TRANSB)) | TF_INIT;
                    t.init = tmp;
                    cut_of_task[tid] = cut;
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
    int64_t nt = (int64_t)G.tasks.size();
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
                int32_t pb = t.pair_begin, np = t.n_pairs;
                if (n_cuts && cut_of_task[tid] >= 0) { pb = G.tasks[tid - 1].pair_begin; np += G.tasks[tid - 1].n_pairs; }
                if (k < np) { G.pairs[pb + k] = Pair{src[i], src2[i]}; pair_op[pb + k] = (int32_t)i; }
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
            const int32_t cut = n_cuts ? cut_of_task[t] : -1;
            if (n_cuts && cut < 0 && t + 1 < nt && cut_of_task[t + 1] >= 0) continue;   // early part: filled through its final task
            int32_t n = T.n_pairs, pb = T.pair_begin;
            if (cut >= 0) { pb = G.tasks[t - 1].pair_begin; n += G.tasks[t - 1].n_pairs; }
            if (fill[t] != n) { mismatch = 1; continue; }
            bool sorted = true;
            for (int32_t k = 1; k < n; k++) sorted = sorted && pair_op[pb + k - 1] < pair_op[pb + k];
            if (sorted && cut < 0) continue;
            // sort by op index (chains are short: up to a few hundred operands)
            std::vector<std::pair<int32_t, Pair>> tmp(n);
            for (int32_t k = 0; k < n; k++) tmp[k] = {pair_op[pb + k], G.pairs[pb + k]};
            if (!sorted) std::sort(tmp.begin(), tmp.end(), [](const auto& x, const auto& y) { return x.first < y.first; });
            if (cut < 0) {
                for (int32_t k = 0; k < n; k++) G.pairs[pb + k] = tmp[k].second;
            } else {
                // the early positions (in op order) first, then the rest (in op order)
                const ChainCut& cc = (*opt.chain_cuts)[cut];
                std::vector<char> early(n, 0);
                for (int32_t q : cc.early_pos) early[q] = 1;
                int32_t w = pb;
                for (int32_t k = 0; k < n; k++) if (early[k]) G.pairs[w++] = tmp[k].second;
                for (int32_t k = 0; k < n; k++) if (!early[k]) G.pairs[w++] = tmp[k].second;
            }
        }
        if (mismatch) return "internal: pair count mismatch";
        if (n_cuts) for (int64_t t = 0; t < nt; t++) G.chain_splits += cut_of_task[t] >= 0;
    }
    lap("pairs");
    // ---- multi-GPU: owners, mirrors of remote blocks and their fetch tasks ---------------------------
    // A task runs on the GPU that owns its result block.  A produced block that a GPU reads at least
    // mirror_min times from a peer is MIRRORED there: a fetch task (a T_SUB "copy": out = remote - 0)
    // owned by the reader copies it once over NVLink as soon as it is complete, and the reader's
    // tasks use the copy (the panel broadcast of the north star, pulled by the consumer).  Mirrors
    // are ordinary blocks (ids appended after the caller's) so slot recycling and segments cover them.
    int64_t nid = n_ids;                       // block ids including mirrors
    const int nown = std::max(1, opt.n_owners);
    G.n_owners = nown;
    G.owner_of.assign(n_ids, 0);
    if (nown > 1) {
        if (nown > MAX_GPUS) return "too many GPUs";
        if (opt.owner_of_id)
            for (int64_t id = 1; id < n_ids_caller; id++) {
                if (opt.owner_of_id[id] < 0 || opt.owner_of_id[id] >= nown) return "block owner out of range";
                G.owner_of[id] = opt.owner_of_id[id];
            }
        for (int64_t c = 0; c < n_cuts; c++) {       // a temporary lives where the block it becomes lives
            const int32_t o = (*opt.chain_cuts)[c].out_id;
            if (o > 0 && o < n_ids_caller) G.owner_of[n_ids_caller + c] = G.owner_of[o];
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
    // ---- pool slots and segments -------------------------------------------------------------------
    // Unlimited pool: every input / produced block gets its own slot, one segment.  Limited pool
    // (opt.max_slots): walk the tasks in order (a topological order: the op list is stage-sorted),
    // hand out free slots, and when none is left close the segment there; at a segment boundary
    // every block whose producer and readers all lie before it is dead and its slot returns to the
    // free list.  Inputs, kept blocks (L, U) and the diagonal inverses the solve uses are pinned.
    auto canon = [&](int32_t id) { return alias_to[id] ? alias_to[id] : id; };
    std::vector<int32_t> seg_of(nt, 0);
    G.recycled.assign(nid, 0);
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
            auto take = [&](int o) -> int32_t {
                if (!freelist[o].empty()) { int32_t sl = freelist[o].back(); freelist[o].pop_back(); return sl; }
                if (G.slots_per_owner[o] < opt.max_slots) return (int32_t)G.slots_per_owner[o]++;
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
    // ---- dependencies: distinct producer tasks of every source, inside the same segment ----------
    // (producers in earlier segments have finished before the launch starts)
    // predecessor lists in one flat array (capacity = operand count per task), sorted and made unique per task
    std::vector<int64_t> poff(nt + 1, 0);
    for (int64_t t = 0; t < nt; t++) poff[t + 1] = poff[t] + 2 * (int64_t)G.tasks[t].n_pairs + ((G.tasks[t].flags & TF_INIT) ? 1 : 0);
    BigVec<int32_t> pflat(poff[nt]);
    std::vector<int32_t> pcnt(nt, 0);
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
    // successor references; a group with exactly one predecessor task (unsplit) carries the "sole predecessor" bit
    G.succ_enc.resize(G.succ.size());
#pragma omp parallel for schedule(static)
    for (int64_t t = 0; t < (int64_t)G.tasks.size(); t++) {
        const Task& T = G.tasks[t];
        if (!task_is_leader(T)) continue;              // the slices of one task share their leader's list
        const bool single = task_group_size(T) == 1;
        for (int32_t e = T.succ_begin; e < T.succ_end; e++) {
            const Task& S = G.tasks[G.succ[e]];
            G.succ_enc[e] = make_task_ref(0, single && S.n_deps == 1, task_log2_slices(S), G.succ[e]);
        }
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
    // ---- chain analysis (diagnostics + input of a second compile, CompileOptions::analyze_chains) --------------
    // Under the cost model of model.h: fin[t] = earliest time a successor of t can start, as compiled (a task starts
    // when ALL operands are ready), and fin_e[t] = the same if every accumulation chain may be cut once into an
    // early part (the pairs that are ready first, plus the initial value) and a late part that starts from it.
    // A cut is proposed when it moves the task's own finish by cut_min_gain_us and the task has little slack.
    if (opt.analyze_chains && !G.tasks.empty()) {
        const ModelParams M;
        const int64_t n2 = (int64_t)G.tasks.size();
        const float o_in = (float)model_in_us(M), o_out = (float)model_out_us(M);
        auto stage_us = [&](const Task& T) -> float { return (float)model_stage_us(T, M); };
        std::vector<float> fin(n2, 0.f), fin_e(n2, 0.f);
        std::vector<int32_t> best_k(n2, 0);
        std::vector<std::pair<float, int32_t>> rs;   // (ready time, position) of the pairs of one chain
        for (int64_t t = 0; t < n2; t++) {
            const Task& T = G.tasks[t];
            if (!task_is_leader(T)) {
                const int64_t lead = t - ((T.flags >> TF_ROW0_SHIFT) & 3) / std::max(1, (T.flags >> TF_NROWS_SHIFT) & 7);
                fin[t] = fin[lead]; fin_e[t] = fin_e[lead];
                continue;
            }
            const float ts = stage_us(T);
            const int n = (T.type == T_GEMM) ? T.n_pairs : 1;
            auto ready = [&](int32_t id, const std::vector<float>& f) -> float {
                if (id <= 0) return 0.f;
                const int32_t p = G.task_of[id];
                return (p >= 0 && p != t) ? f[p] : 0.f;
            };
            float r_all = 0.f, ri = 0.f, ri_e = 0.f;
            if (T.flags & TF_INIT) { ri = ready(T.init, fin); ri_e = ready(T.init, fin_e); }
            rs.clear();
            for (int k = 0; k < T.n_pairs; k++) {
                const Pair& pr = G.pairs[T.pair_begin + k];
                r_all = std::max(r_all, std::max(ready(pr.a, fin), ready(pr.b, fin)));
                if (T.type == T_GEMM) rs.push_back({std::max(ready(pr.a, fin_e), ready(pr.b, fin_e)), k});
            }
            fin[t] = std::max(r_all, ri) + o_in + n * ts + o_out;
            if (T.type != T_GEMM) {
                float r = 0.f;
                for (int k = 0; k < T.n_pairs; k++) r = std::max(r, std::max(ready(G.pairs[T.pair_begin + k].a, fin_e), ready(G.pairs[T.pair_begin + k].b, fin_e)));
                fin_e[t] = r + o_in + n * ts + o_out;
                continue;
            }
            std::stable_sort(rs.begin(), rs.end(), [](const auto& x, const auto& y) { return x.first < y.first; });
            const float r_last = rs.back().first;
            float best = std::max(r_last, ri_e) + o_in + n * ts + o_out;
            int bk = 0;
            for (int k = 1; k < n; k++) {
                const float fe = std::max(rs[k - 1].first, ri_e) + o_in + k * ts + o_out;
                const float f = std::max(fe, r_last) + o_in + (n - k) * ts + o_out;
                if (f < best - (float)opt.cut_min_gain_us) { best = f; bk = k; }
            }
            fin_e[t] = best;
            best_k[t] = bk;
        }
        for (int64_t t = 0; t < n2; t++) { G.cp_us = std::max(G.cp_us, (double)fin[t]); G.cp_early_us = std::max(G.cp_early_us, (double)fin_e[t]); }
        // slack under the early-start times: a cut only matters on (near-)critical tasks
        std::vector<float> bot(n2, 0.f);     // longest path from the END of the task to the end of the factorisation
        for (int64_t t = n2 - 1; t >= 0; t--) {
            const Task& T = G.tasks[t];
            float m = 0.f;
            for (int32_t e = T.succ_begin; e < T.succ_end; e++) {
                const Task& S = G.tasks[G.succ[e]];
                m = std::max(m, bot[G.succ[e]] + o_in + ((S.type == T_GEMM) ? S.n_pairs : 1) * stage_us(S) + o_out);
            }
            bot[t] = m;
        }
        for (int64_t t = 0; t < n2; t++) {
            const Task& T = G.tasks[t];
            if (T.type != T_GEMM || !task_is_leader(T) || best_k[t] == 0) continue;
            if (G.cp_early_us - (fin_e[t] + bot[t]) > opt.cut_max_slack_us) continue;
            // recompute the order of the chain (cheap: only the chosen tasks)
            rs.clear();
            for (int k = 0; k < T.n_pairs; k++) {
                const Pair& pr = G.pairs[T.pair_begin + k];
                auto rd = [&](int32_t id) { if (id <= 0) return 0.f; const int32_t p = G.task_of[id]; return (p >= 0 && p != t) ? fin_e[p] : 0.f; };
                rs.push_back({std::max(rd(pr.a), rd(pr.b)), k});
            }
            std::stable_sort(rs.begin(), rs.end(), [](const auto& x, const auto& y) { return x.first < y.first; });
            ChainCut c;
            c.out_id = T.out;
            c.n_early = best_k[t];
            for (int k = 0; k < best_k[t]; k++) c.early_pos.push_back(rs[k].second);
            std::sort(c.early_pos.begin(), c.early_pos.end());
            G.cuts.push_back(std::move(c));
        }
        // Shared-operand pairs (diagnostics): two bulk GEMM tasks with the same left-operand sequence could
        // run as ONE task that loads every A block once (3 instead of 4 block loads per two products).  Candidates:
        // whole-block tasks with plenty of slack, the same flags and chain, ready at about the same time, and no
        // successor of the first in between (the pair would sit at the second task's place in the order).
        {
            struct Cand { uint64_t key; int32_t task; };
            std::vector<Cand> cand;
            for (int64_t t = 0; t < n2; t++) {
                const Task& T = G.tasks[t];
                if (T.type != T_GEMM || T.n_pairs < 2 || ((T.flags >> TF_NROWS_SHIFT) & 7) != 4) continue;
                if (G.cp_early_us - (fin_e[t] + bot[t]) < opt.dual_min_slack_us) continue;
                uint64_t h = 1469598103934665603ull ^ (uint64_t)(T.flags & (TF_NEGATE | TF_TRANSB | TF_INIT)) ^ ((uint64_t)T.n_pairs << 8) ^ ((uint64_t)G.task_owner[t] << 40);
                for (int k = 0; k < T.n_pairs; k++) { h ^= (uint64_t)(uint32_t)G.pairs[T.pair_begin + k].a; h *= 1099511628211ull; }
                cand.push_back({h, (int32_t)t});
            }
            std::sort(cand.begin(), cand.end(), [](const Cand& x, const Cand& y) { return x.key != y.key ? x.key < y.key : x.task < y.task; });
            int64_t pairs_formed = 0, covered = 0;
            for (size_t q = 0; q + 1 < cand.size();) {
                const int32_t t1 = cand[q].task, t2 = cand[q + 1].task;
                bool ok = cand[q].key == cand[q + 1].key && std::fabs(fin_e[t1] - fin_e[t2]) < 50.f;
                if (ok) {
                    const Task &A = G.tasks[t1], &B = G.tasks[t2];
                    for (int k = 0; k < A.n_pairs && ok; k++) ok = G.pairs[A.pair_begin + k].a == G.pairs[B.pair_begin + k].a;
                    for (int32_t e = A.succ_begin; e < A.succ_end && ok; e++) ok = G.succ[e] > t2;
                }
                if (ok) { pairs_formed++; covered += 2 * (int64_t)G.tasks[t1].n_pairs; q += 2; }
                else q++;
            }
            G.dual_pairs = pairs_formed;
            G.dual_covered_pairs = covered;
        }
        lap("chain analysis");
    }
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
                    D.succ.push_back(make_task_ref(G.task_owner[s2], (G.succ_enc[e] & TASK_SOLE_BIT) != 0, task_log2_slices(G.tasks[s2]), D.task_local[s2]));
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
