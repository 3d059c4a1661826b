// Driver: the reference's SOGLU::solveLU / decompose_solveLU (solver.cpp:50-184) on top of
// the C ABI of include/soglu.h.  The host does ordering + planning (bit-exact integer
// work), the GPU does BlockPlanner::calculate and BlockPlanner::solve.
#include "solver.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <new>
#include <stdexcept>
#include <vector>

#include "problem.h"

using soglu::Problem;

extern "C" {

static int upload_matrix(soglu_ctx* ctx, const Problem* p);

int soglu_load_problem(soglu_ctx* ctx, const soglu_problem* pp) {
    try {
    const Problem* p = reinterpret_cast<const Problem*>(pp);
    if (!ctx || !p) { soglu::set_error("bad argument"); return SOGLU_ERR_ARG; }
    const soglu::Plan& pl = p->plan;
    // input blocks: ids are 1..n_input in allocation order; their values go up as the sparse entry list
    const int64_t n_in = (int64_t)pl.inputs.size();
    std::vector<int32_t> in_ids(n_in);
    for (int64_t k = 0; k < n_in; k++) in_ids[k] = (int32_t)(k + 1);
    const int64_t n_ent = (int64_t)pl.entry_val.size();
    soglu::BigVec<int32_t> ent_in(n_ent);
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < n_ent; k++) ent_in[k] = pl.entry_block[k] - 1;
    int rc = soglu_set_blocks_sparse(ctx, pl.storage, n_in, in_ids.data(), n_ent, ent_in.data(), pl.entry_pos.data(), pl.entry_val.data());
    if (rc) return rc;
    const int64_t n = (int64_t)pl.ops.size();
    soglu::BigVec<int32_t> src(n), src2(n), res(n), res2(n), stg(n);
    soglu::BigVec<uint8_t> op(n);
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < n; k++) {
        const soglu::Op& o = pl.ops[k];
        src[k] = o.src; src2[k] = o.src2; res[k] = o.result; res2[k] = o.result2; stg[k] = o.stage; op[k] = o.op;
    }
    rc = soglu_set_graph(ctx, n, src.data(), src2.data(), op.data(), res.data(), res2.data(), stg.data(), pl.brow.data(), pl.bcol.data());
    if (rc) return rc;
    auto split = [](const std::vector<soglu::BlockRef>& v, std::vector<int32_t>& id, std::vector<int32_t>& r, std::vector<int32_t>& c) {
        id.resize(v.size()); r.resize(v.size()); c.resize(v.size());
        for (size_t k = 0; k < v.size(); k++) { id[k] = v[k].id; r[k] = v[k].brow; c[k] = v[k].bcol; }
    };
    std::vector<int32_t> li, lr, lc, ui, ur, uc;
    split(pl.L, li, lr, lc);
    split(pl.U, ui, ur, uc);
    rc = soglu_set_factors(ctx, (int64_t)li.size(), li.data(), lr.data(), lc.data(), (int64_t)ui.size(), ui.data(), ur.data(), uc.data(),
                           p->cfg.blockRows, p->symmetric ? 1 : 0);
    if (rc) return rc;
    return upload_matrix(ctx, p);
    } catch (const std::bad_alloc&) {
        soglu::set_error("out of host memory");
        return SOGLU_ERR_OOM;
    } catch (const std::exception& e) {
        soglu::set_error(std::string("internal error: ") + e.what());
        return SOGLU_ERR_ARG;
    }
}

// CSR of the permuted matrix, padded with the identity like the planner does (BlockPlanner.cpp:1520-1539); duplicates
// keep the last value, as in the block scatter (BlockPlanner.cpp:1510).  Uploaded once per loaded problem: the
// residual of the iterative refinement (soglu_solve_refined) is formed with it on the device.
static int upload_matrix(soglu_ctx* ctx, const Problem* p) {
    const int64_t n = p->n_ext, nnz = (int64_t)p->pv.size();
    std::vector<int64_t> rp(n + 1, 0);
    for (int64_t k = 0; k < nnz; k++) rp[p->pi[k] + 1]++;
    for (int64_t i = p->dim; i < n; i++) rp[i + 1]++;
    for (int64_t i = 0; i < n; i++) rp[i + 1] += rp[i];
    std::vector<int32_t> ci(rp[n]);
    std::vector<double> cv(rp[n]);
    std::vector<int64_t> pos(rp.begin(), rp.end() - 1);
    for (int64_t k = 0; k < nnz; k++) { ci[pos[p->pi[k]]] = p->pj[k]; cv[pos[p->pi[k]]++] = p->pv[k]; }
    for (int64_t i = p->dim; i < n; i++) { ci[pos[i]] = (int32_t)i; cv[pos[i]++] = 1.0; }
    return soglu_set_matrix(ctx, n, rp[n], rp.data(), ci.data(), cv.data());
}

int soglu_solve_problem(soglu_ctx* ctx, const soglu_problem* pp, const double* b, double* x, int refine, soglu_stats* out) {
    try {
    const Problem* p = reinterpret_cast<const Problem*>(pp);
    if (!ctx || !p || !x) { soglu::set_error("bad argument"); return SOGLU_ERR_ARG; }
    // (the CSR of the permuted matrix for the refinement's residual was uploaded by soglu_load_problem)
    std::vector<double> bext;
    const double* bp = p->b_perm.data();
    if (b) {   // permute + pad a caller-supplied rhs exactly like the problem's own (GPSOrder.cpp:435-447)
        bext.assign(p->n_ext, 1.0);
        for (int i = 0; i < p->dim; i++) bext[i] = b[p->ord.newOrder[i]];
        bp = bext.data();
    }
    std::vector<double> xext(p->n_ext);
    int rc = refine > 0 ? soglu_solve_refined(ctx, bp, xext.data(), refine, out) : soglu_solve(ctx, bp, xext.data(), out);
    if (rc) return rc;
    // un-permute (GOrder::reOrderResult, GPSOrder.cpp:41-53)
    for (int i = 0; i < p->dim; i++) x[i] = xext[p->ord.reverseOrder[i]];
    return SOGLU_OK;
    } catch (const std::bad_alloc&) {
        soglu::set_error("out of host memory");
        return SOGLU_ERR_OOM;
    } catch (const std::exception& e) {
        soglu::set_error(std::string("internal error: ") + e.what());
        return SOGLU_ERR_ARG;
    }
}

double* soglu_solveLU(int dim, int valcount, int symmetric, const int* index_i, const int* index_j, const double* vals, const double* b) {
    soglu_problem* prob = nullptr;
    if (soglu_problem_from_coo(dim, valcount, symmetric, index_i, index_j, vals, b, &prob)) {
        std::cout << "soglu error: " << soglu_last_error() << std::endl;
        return nullptr;
    }
    const Problem* p = reinterpret_cast<const Problem*>(prob);
    std::cout << p->log;
    soglu_ctx* ctx = nullptr;
    double* x = nullptr;
    soglu_stats fs, ss;
    do {
        // SOGLU_GPUS=N: shard the factorisation over GPUs 0..N-1 of this process (soglu_create with n_gpus > 1)
        const char* env_gpus = std::getenv("SOGLU_GPUS");
        const int n_gpus = env_gpus ? std::max(1, std::atoi(env_gpus)) : 1;
        if (soglu_create(&ctx, n_gpus, nullptr)) break;
        if (n_gpus > 1) std::cout << "sharded over " << n_gpus << " GPUs (2D block-cyclic block ownership)\n";
        if (soglu_load_problem(ctx, prob)) break;
        if (soglu_factor(ctx, &fs)) break;
        for (int64_t k = 0, n_bad = soglu_diag_warnings(ctx); k < n_bad && k < 8; k++) std::cout << " upper out of tolerance " << std::endl;   // (BlockPlanner.cpp:576; capped)
        std::cout << "kernel time: " << fs.seconds << "  (" << fs.flops / fs.seconds * 1e-9 << " GFLOP/s, " << fs.tasks << " tasks)\n";
        x = (double*)std::malloc(sizeof(double) * dim);
        if (soglu_solve_problem(ctx, prob, nullptr, x, 0, &ss)) { std::free(x); x = nullptr; break; }
        std::cout << "solve triangled :" << ss.seconds << "  (" << ss.bytes / ss.seconds * 1e-9 << " GB/s)\n";
        int nan = 0;
        for (int i = 0; i < dim; i++) nan += (x[i] != x[i]);
        if (nan) std::cout << "found NaN:  " << nan << std::endl;
    } while (0);
    if (!x) std::cout << "soglu error: " << soglu_last_error() << std::endl;
    if (ctx) soglu_destroy(ctx);
    soglu_problem_free(prob);
    return x;
}

}  // extern "C"

namespace SOGLU {
int iniData() { return 0; }
double* solveLU(int dim, int valcount, bool symmetric, int* index_i, int* index_j, double* vals, double* b) {
    return soglu_solveLU(dim, valcount, symmetric ? 1 : 0, index_i, index_j, vals, b);
}
}  // namespace SOGLU

// ---- host-only view of the task compiler (no GPU needed): used by the CPU tests -------------------
#include "../device/tasks.h"
#include <chrono>
extern "C" int soglu_debug_compile(const soglu_problem* pp, int fuse_sub, int fuse_inv, int split, int64_t max_slots, int64_t* out, int n_out);
// same with a process grid: out[16..] = per-owner {tasks, slots, mirrors} triples
extern "C" int soglu_debug_compile_dist(const soglu_problem* pp, int64_t max_slots, int pr, int pc, int nb, int64_t* out, int n_out) {
    const Problem* p = reinterpret_cast<const Problem*>(pp);
    const int world = pr * pc;
    if (!p || !out || n_out < 16 + 3 * world) { soglu::set_error("bad argument"); return SOGLU_ERR_ARG; }
    const soglu::Plan& pl = p->plan;
    const int64_t n = (int64_t)pl.ops.size();
    std::vector<int32_t> src(n), src2(n), res(n), res2(n);
    std::vector<uint8_t> op(n);
    for (int64_t k = 0; k < n; k++) { const soglu::Op& o = pl.ops[k]; src[k] = o.src; src2[k] = o.src2; res[k] = o.result; res2[k] = o.result2; op[k] = o.op; }
    std::vector<int32_t> in_ids(pl.inputs.size()), keep;
    for (size_t k = 0; k < in_ids.size(); k++) in_ids[k] = (int32_t)(k + 1);
    for (const auto& r : pl.L) keep.push_back(r.id);
    for (const auto& r : pl.U) keep.push_back(r.id);
    std::vector<int8_t> owners(pl.storage, 0);
    for (int64_t id = 1; id < pl.storage; id++)
        if (pl.brow[id] >= 0 && pl.bcol[id] >= 0) owners[id] = (int8_t)(((pl.brow[id] / nb) % pr) * pc + ((pl.bcol[id] / nb) % pc));
    soglu::CompileOptions co;
    co.max_slots = max_slots; co.owner_of_id = owners.data(); co.n_owners = world;
    soglu::TaskGraph G;
    std::string err = soglu::compile_tasks(pl.storage, (int64_t)in_ids.size(), in_ids.data(), n, src.data(), src2.data(), op.data(), res.data(), res2.data(), keep, co, G);
    if (!err.empty()) { soglu::set_error(err); return SOGLU_ERR_GRAPH; }
    // deps = sum of the group counters (leaders); succ = decrements at run time (every slice walks its task's list)
    int64_t deps = 0, releases = 0;
    for (const soglu::Task& t : G.tasks) { if (soglu::task_is_leader(t)) deps += t.n_deps; releases += t.succ_end - t.succ_begin; }
    int64_t v[16] = {(int64_t)G.tasks.size(), (int64_t)G.pairs.size(), releases, (int64_t)G.initial.size(), G.n_slots, G.n_levels,
                     G.fused_subs, G.fused_invs, G.aliased_invs, G.split_tasks, (int64_t)G.seg_begin.size() - 1, deps, 0, 0, 0, (int64_t)(G.flops * 1e-6)};
    for (int i = 0; i < 16; i++) out[i] = v[i];
    int64_t remote_edges = 0;
    for (int r = 0; r < world; r++) {
        soglu::DistLayout D;
        err = soglu::localize_tasks(G, r, D);
        if (!err.empty()) { soglu::set_error(err); return SOGLU_ERR_GRAPH; }
        out[16 + 3 * r] = (int64_t)D.tasks.size(); out[17 + 3 * r] = G.slots_per_owner[r]; out[18 + 3 * r] = D.mirrored;
        remote_edges += D.remote_edges;
        int64_t dd = 0, rr = 0;
        for (const soglu::Task& t : D.tasks) { if (soglu::task_is_leader(t)) dd += t.n_deps; rr += t.succ_end - t.succ_begin; }
        out[12] += dd; out[13] += rr; out[14] += D.remote_operands;
    }
    return SOGLU_OK;
}
extern "C" int soglu_debug_compile(const soglu_problem* pp, int fuse_sub, int fuse_inv, int split, int64_t max_slots, int64_t* out, int n_out) {
    const Problem* p = reinterpret_cast<const Problem*>(pp);
    if (!p || !out || n_out < 16) { soglu::set_error("bad argument"); return SOGLU_ERR_ARG; }
    const soglu::Plan& pl = p->plan;
    const int64_t n = (int64_t)pl.ops.size();
    std::vector<int32_t> src(n), src2(n), res(n), res2(n);
    std::vector<uint8_t> op(n);
    for (int64_t k = 0; k < n; k++) { const soglu::Op& o = pl.ops[k]; src[k] = o.src; src2[k] = o.src2; res[k] = o.result; res2[k] = o.result2; op[k] = o.op; }
    std::vector<int32_t> in_ids(pl.inputs.size()), keep;
    for (size_t k = 0; k < in_ids.size(); k++) in_ids[k] = (int32_t)(k + 1);
    for (const auto& r : pl.L) keep.push_back(r.id);
    for (const auto& r : pl.U) keep.push_back(r.id);
    soglu::CompileOptions co;
    co.fuse_sub = fuse_sub != 0; co.fuse_inv = fuse_inv != 0; co.split_narrow = split; co.max_slots = max_slots;
    soglu::TaskGraph G;
    auto t0 = std::chrono::steady_clock::now();
    std::string err = soglu::compile_tasks(pl.storage, (int64_t)in_ids.size(), in_ids.data(), n, src.data(), src2.data(), op.data(), res.data(), res2.data(), keep, co, G);
    double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (!err.empty()) { soglu::set_error(err); return SOGLU_ERR_GRAPH; }
    int64_t deps = 0, releases = 0, maxdeps = 0, gemm = 0, lu = 0, subs = 0;
    for (const soglu::Task& t : G.tasks) {
        if (soglu::task_is_leader(t)) deps += t.n_deps;
        releases += t.succ_end - t.succ_begin;
        maxdeps = std::max<int64_t>(maxdeps, t.n_deps);
        gemm += t.type == soglu::T_GEMM; lu += t.type == soglu::T_LU; subs += t.type == soglu::T_SUB;
    }
    int64_t v[16] = {(int64_t)G.tasks.size(), (int64_t)G.pairs.size(), releases, (int64_t)G.initial.size(), G.n_slots, G.n_levels,
                     G.fused_subs, G.fused_invs, G.aliased_invs, G.split_tasks, (int64_t)G.seg_begin.size() - 1, deps, maxdeps, gemm, lu,
                     (int64_t)(dt * 1e6)};
    for (int i = 0; i < 16; i++) out[i] = v[i];
    (void)subs;
    return SOGLU_OK;
}

// Host simulation of the executor's release protocol on a compiled graph (one segment, optional process grid):
// tasks are taken from the ready set in a seeded random order; a finishing task decrements the counter of every
// successor GROUP LEADER once and the whole group becomes ready when it reaches zero -- exactly what the CUDA
// executor does with its localized per-GPU arrays.  Checks that every task runs exactly once and that each operand
// block is complete (all of its writers done) when a reader starts.  out[0] = tasks run, out[1] = violations.
extern "C" int soglu_debug_simulate(const soglu_problem* pp, int split, int pr, int pc, int nb, uint64_t seed, int64_t* out) {
    const Problem* p = reinterpret_cast<const Problem*>(pp);
    const int world = pr * pc;
    if (!p || !out || world < 1 || world > soglu::MAX_GPUS) { soglu::set_error("bad argument"); return SOGLU_ERR_ARG; }
    const soglu::Plan& pl = p->plan;
    const int64_t n = (int64_t)pl.ops.size();
    std::vector<int32_t> src(n), src2(n), res(n), res2(n);
    std::vector<uint8_t> op(n);
    for (int64_t k = 0; k < n; k++) { const soglu::Op& o = pl.ops[k]; src[k] = o.src; src2[k] = o.src2; res[k] = o.result; res2[k] = o.result2; op[k] = o.op; }
    std::vector<int32_t> in_ids(pl.inputs.size()), keep;
    for (size_t k = 0; k < in_ids.size(); k++) in_ids[k] = (int32_t)(k + 1);
    for (const auto& r : pl.L) keep.push_back(r.id);
    for (const auto& r : pl.U) keep.push_back(r.id);
    std::vector<int8_t> owners(pl.storage, 0);
    if (world > 1)
        for (int64_t id = 1; id < pl.storage; id++)
            if (pl.brow[id] >= 0 && pl.bcol[id] >= 0) owners[id] = (int8_t)(((pl.brow[id] / nb) % pr) * pc + ((pl.bcol[id] / nb) % pc));
    soglu::CompileOptions co;
    co.split_narrow = split;
    if (world > 1) { co.owner_of_id = owners.data(); co.n_owners = world; }
    soglu::TaskGraph G;
    std::string err = soglu::compile_tasks(pl.storage, (int64_t)in_ids.size(), in_ids.data(), n, src.data(), src2.data(), op.data(), res.data(), res2.data(), keep, co, G);
    if (!err.empty()) { soglu::set_error(err); return SOGLU_ERR_GRAPH; }
    // per-GPU arrays, as uploaded
    std::vector<soglu::DistLayout> D(world);
    for (int r = 0; r < world; r++) {
        err = soglu::localize_tasks(G, r, D[r]);
        if (!err.empty()) { soglu::set_error(err); return SOGLU_ERR_GRAPH; }
    }
    int64_t max_slots = 0;
    for (int r = 0; r < world; r++) max_slots = std::max<int64_t>(max_slots, G.slots_per_owner[r]);
    auto key = [&](int32_t ref) { return (int64_t)((uint32_t)ref >> soglu::REF_SHIFT) * max_slots + (ref & soglu::REF_MASK); };
    std::vector<int32_t> writers((size_t)world * max_slots, 0);
    auto for_outs = [&](const soglu::Task& T, auto&& fn) {
        fn(T.out);
        if (T.type == soglu::T_LU) fn(T.out2);
        if (T.type == soglu::T_LU || T.type == soglu::T_LLT) { if (T.flags & soglu::TF_LINV) fn(T.init); if (T.flags & soglu::TF_UINV) fn(T.out4); }
    };
    for (int r = 0; r < world; r++)
        for (const soglu::Task& T : D[r].tasks) for_outs(T, [&](int32_t ref) { writers[key(ref)]++; });
    std::vector<std::vector<int32_t>> dep(world), runs(world);
    std::vector<std::pair<int8_t, int32_t>> ready;
    for (int r = 0; r < world; r++) {
        dep[r].resize(D[r].tasks.size());
        runs[r].assign(D[r].tasks.size(), 0);
        for (size_t t = 0; t < D[r].tasks.size(); t++) dep[r][t] = D[r].tasks[t].n_deps;
        for (int32_t t : D[r].initial) ready.push_back({(int8_t)r, t});
    }
    uint64_t rng = seed * 6364136223846793005ull + 1442695040888963407ull;
    int64_t done = 0, bad = 0;
    if (seed == 0) {
        // The executor's own discipline (executor.cu): every GPU has a few workers, each claims the next task of its
        // GPU IN TASK ORDER and waits until the group's counter is zero; a finished task counts every successor group
        // down by one.  A sweep over all workers without progress while tasks remain = the order would deadlock.
        const int workers = 3;
        std::vector<int32_t> next(world, 0);
        std::vector<std::vector<int32_t>> held(world, std::vector<int32_t>(workers, -1));
        auto leader = [&](int r, int32_t t) {
            const soglu::Task& T = D[r].tasks[t];
            const int rows16 = (T.flags >> soglu::TF_NROWS_SHIFT) & 7;
            return (T.type == soglu::T_GEMM && rows16 > 0 && rows16 < 4) ? t - ((T.flags >> soglu::TF_ROW0_SHIFT) & 3) / rows16 : t;
        };
        int64_t total0 = 0;
        for (int r = 0; r < world; r++) total0 += (int64_t)D[r].tasks.size();
        bool progress = true;
        while (progress) {
            progress = false;
            for (int r = 0; r < world; r++)
                for (int w = 0; w < workers; w++) {
                    int32_t& h = held[r][(w + (int)(done % workers)) % workers];
                    if (h < 0 && next[r] < (int32_t)D[r].tasks.size()) { h = next[r]++; progress = true; }
                    if (h < 0 || dep[r][leader(r, h)] > 0) continue;
                    const soglu::Task& T = D[r].tasks[h];
                    if (runs[r][h]++) bad++;
                    for (int32_t k = 0; k < T.n_pairs; k++) {
                        const soglu::Pair& pq = D[r].pairs[T.pair_begin + k];
                        if (writers[key(pq.a)] != 0) bad++;
                        if ((T.type == soglu::T_GEMM || T.type == soglu::T_SUB) && writers[key(pq.b)] != 0) bad++;
                    }
                    if ((T.flags & soglu::TF_INIT) && writers[key(T.init)] != 0) bad++;
                    for_outs(T, [&](int32_t ref) { writers[key(ref)]--; });
                    for (int32_t e = T.succ_begin; e < T.succ_end; e++) {
                        const int32_t ref = D[r].succ[e];
                        const int o = (uint32_t)ref >> soglu::REF_SHIFT, nx = ref & soglu::TASK_LOCAL_MASK;
                        if (--dep[o][nx] < 0) bad++;
                    }
                    done++;
                    h = -1;
                    progress = true;
                }
        }
        if (done != total0) bad++;      // stuck
        ready.clear();
    }
    while (!ready.empty()) {
        rng = rng * 6364136223846793005ull + 1442695040888963407ull;
        const size_t pick = (size_t)((rng >> 33) % ready.size());
        const auto cur = ready[pick];
        ready[pick] = ready.back();
        ready.pop_back();
        const int r = cur.first;
        const soglu::Task& T = D[r].tasks[cur.second];
        if (runs[r][cur.second]++) bad++;
        for (int32_t k = 0; k < T.n_pairs; k++) {
            const soglu::Pair& pq = D[r].pairs[T.pair_begin + k];
            if (writers[key(pq.a)] != 0) bad++;
            if ((T.type == soglu::T_GEMM || T.type == soglu::T_SUB) && writers[key(pq.b)] != 0) bad++;
        }
        if ((T.flags & soglu::TF_INIT) && writers[key(T.init)] != 0) bad++;
        for_outs(T, [&](int32_t ref) { writers[key(ref)]--; });
        done++;
        for (int32_t e = T.succ_begin; e < T.succ_end; e++) {
            const int32_t ref = D[r].succ[e];
            const int o = (uint32_t)ref >> soglu::REF_SHIFT, nx = ref & soglu::TASK_LOCAL_MASK, g = 1 << ((ref >> soglu::TASK_SPLIT_SHIFT) & 3);
            if (--dep[o][nx] == 0)
                for (int q = 0; q < g; q++) ready.push_back({(int8_t)o, nx + q});
        }
    }
    int64_t total = 0;
    for (int r = 0; r < world; r++) {
        total += (int64_t)D[r].tasks.size();
        for (int32_t c : runs[r]) if (c != 1) bad++;
    }
    out[0] = done; out[1] = bad; out[2] = total;
    return SOGLU_OK;
}

// FNV-1a over everything compile_tasks produces (regression guard for refactorings of the compiler: the task
// order, operand order and slot assignment decide the rounding of the result, so they must not drift)
extern "C" uint64_t soglu_debug_graph_hash(const soglu_problem* pp, int split, int64_t max_slots, int pr, int pc, int nb) {
    const Problem* p = reinterpret_cast<const Problem*>(pp);
    const int world = pr * pc;
    if (!p || world < 1 || world > soglu::MAX_GPUS) return 0;
    const soglu::Plan& pl = p->plan;
    const int64_t n = (int64_t)pl.ops.size();
    std::vector<int32_t> src(n), src2(n), res(n), res2(n);
    std::vector<uint8_t> op(n);
    for (int64_t k = 0; k < n; k++) { const soglu::Op& o = pl.ops[k]; src[k] = o.src; src2[k] = o.src2; res[k] = o.result; res2[k] = o.result2; op[k] = o.op; }
    std::vector<int32_t> in_ids(pl.inputs.size()), keep;
    for (size_t k = 0; k < in_ids.size(); k++) in_ids[k] = (int32_t)(k + 1);
    for (const auto& r : pl.L) keep.push_back(r.id);
    for (const auto& r : pl.U) keep.push_back(r.id);
    std::vector<int8_t> owners(pl.storage, 0);
    if (world > 1)
        for (int64_t id = 1; id < pl.storage; id++)
            if (pl.brow[id] >= 0 && pl.bcol[id] >= 0) owners[id] = (int8_t)(((pl.brow[id] / nb) % pr) * pc + ((pl.bcol[id] / nb) % pc));
    soglu::CompileOptions co;
    co.split_narrow = split; co.max_slots = max_slots;
    if (world > 1) { co.owner_of_id = owners.data(); co.n_owners = world; }
    soglu::TaskGraph G;
    if (!soglu::compile_tasks(pl.storage, (int64_t)in_ids.size(), in_ids.data(), n, src.data(), src2.data(), op.data(), res.data(), res2.data(), keep, co, G).empty()) return 0;
    uint64_t h = 1469598103934665603ull;
    auto mix = [&](const void* d, size_t bytes) { const unsigned char* c = (const unsigned char*)d; for (size_t i = 0; i < bytes; i++) { h ^= c[i]; h *= 1099511628211ull; } };
    mix(G.tasks.data(), G.tasks.size() * sizeof(soglu::Task));
    mix(G.pairs.data(), G.pairs.size() * sizeof(soglu::Pair));
    mix(G.succ.data(), G.succ.size() * 4);
    mix(G.succ_enc.data(), G.succ_enc.size() * 4);
    mix(G.initial.data(), G.initial.size() * 4);
    mix(G.slot_of.data(), G.slot_of.size() * 4);
    mix(G.task_of.data(), G.task_of.size() * 4);
    mix(G.seg_begin.data(), G.seg_begin.size() * 4);
    mix(G.seg_init.data(), G.seg_init.size() * 4);
    mix(G.recycled.data(), G.recycled.size());
    mix(G.owner_of.data(), G.owner_of.size());
    mix(G.task_owner.data(), G.task_owner.size());
    return h;
}

// compile a raw op list (the arrays of soglu_set_blocks / soglu_set_graph / soglu_set_factors) on the host only:
// lets the CPU tests reach every validation error of the task compiler.  out = {tasks, pairs, slots, segments}.
extern "C" int soglu_debug_compile_raw(int64_t n_ids, int64_t n_input, const int32_t* input_ids, int64_t n_ops, const int32_t* src,
                                       const int32_t* src2, const uint8_t* op, const int32_t* result, const int32_t* result2,
                                       int64_t n_keep, const int32_t* keep_ids, int64_t max_slots, int64_t* out) {
    std::vector<int32_t> keep(keep_ids, keep_ids + (keep_ids ? n_keep : 0));
    soglu::CompileOptions co;
    co.max_slots = max_slots;
    soglu::TaskGraph G;
    std::string err = soglu::compile_tasks(n_ids, n_input, input_ids, n_ops, src, src2, op, result, result2, keep, co, G);
    if (!err.empty()) { soglu::set_error(err); return SOGLU_ERR_GRAPH; }
    if (out) { out[0] = (int64_t)G.tasks.size(); out[1] = (int64_t)G.pairs.size(); out[2] = G.n_slots; out[3] = (int64_t)G.seg_begin.size() - 1; }
    return SOGLU_OK;
}
