// Driver: the reference's SOGLU::solveLU / decompose_solveLU (solver.cpp:50-184) on top of
// the C ABI of include/soglu.h.  The host does ordering + planning (bit-exact integer
// work), the GPU does BlockPlanner::calculate and BlockPlanner::solve.
#include "solver.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <vector>

#include "problem.h"

using soglu::Problem;

extern "C" {

int soglu_load_problem(soglu_ctx* ctx, const soglu_problem* pp) {
    const Problem* p = reinterpret_cast<const Problem*>(pp);
    if (!ctx || !p) { soglu::set_error("bad argument"); return SOGLU_ERR_ARG; }
    const soglu::Plan& pl = p->plan;
    // input blocks: ids are 1..n_input in allocation order and input_vals is indexed by id-1
    const int64_t n_in = (int64_t)pl.inputs.size();
    std::vector<int32_t> in_ids(n_in);
    for (int64_t k = 0; k < n_in; k++) in_ids[k] = (int32_t)(k + 1);
    int rc = soglu_set_blocks(ctx, pl.storage, n_in, in_ids.data(), pl.input_vals.data());
    if (rc) return rc;
    const int64_t n = (int64_t)pl.ops.size();
    std::vector<int32_t> src(n), src2(n), res(n), res2(n), stg(n);
    std::vector<uint8_t> op(n);
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
    return soglu_set_factors(ctx, (int64_t)li.size(), li.data(), lr.data(), lc.data(), (int64_t)ui.size(), ui.data(), ur.data(), uc.data(),
                             p->cfg.blockRows, p->symmetric ? 1 : 0);
}

int soglu_solve_problem(soglu_ctx* ctx, const soglu_problem* pp, const double* b, double* x, int refine, soglu_stats* out) {
    const Problem* p = reinterpret_cast<const Problem*>(pp);
    if (!ctx || !p || !x) { soglu::set_error("bad argument"); return SOGLU_ERR_ARG; }
    if (refine > 0) {
        // CSR of the permuted matrix, padded with the identity like the planner does (BlockPlanner.cpp:1520-1539);
        // duplicates keep the last value, as in the block scatter (BlockPlanner.cpp:1510)
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
        int rcm = soglu_set_matrix(ctx, n, rp[n], rp.data(), ci.data(), cv.data());
        if (rcm) return rcm;
    }
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
}

double* soglu_solveLU(int dim, int valcount, int symmetric, const int* index_i, const int* index_j, const double* vals, const double* b) {
    soglu_problem* prob = nullptr;
    if (soglu_problem_from_coo(dim, valcount, symmetric, index_i, index_j, vals, b, &prob)) return nullptr;
    const Problem* p = reinterpret_cast<const Problem*>(prob);
    std::cout << p->log;
    soglu_ctx* ctx = nullptr;
    double* x = nullptr;
    soglu_stats fs, ss;
    do {
        if (soglu_create(&ctx, 1, nullptr)) break;
        if (soglu_load_problem(ctx, prob)) break;
        if (soglu_factor(ctx, &fs)) break;
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
    int64_t deps = 0;
    for (const soglu::Task& t : G.tasks) deps += t.n_deps;
    int64_t v[16] = {(int64_t)G.tasks.size(), (int64_t)G.pairs.size(), (int64_t)G.succ.size(), (int64_t)G.initial.size(), G.n_slots, G.n_levels,
                     G.fused_subs, G.fused_invs, G.aliased_invs, G.split_tasks, (int64_t)G.seg_begin.size() - 1, deps, 0, 0, 0, (int64_t)(G.flops * 1e-6)};
    for (int i = 0; i < 16; i++) out[i] = v[i];
    int64_t remote_edges = 0;
    for (int r = 0; r < world; r++) {
        soglu::DistLayout D;
        err = soglu::localize_tasks(G, r, D);
        if (!err.empty()) { soglu::set_error(err); return SOGLU_ERR_GRAPH; }
        out[16 + 3 * r] = (int64_t)D.tasks.size(); out[17 + 3 * r] = G.slots_per_owner[r]; out[18 + 3 * r] = D.mirrored;
        remote_edges += D.remote_edges;
        int64_t dd = 0; for (const soglu::Task& t : D.tasks) dd += t.n_deps;
        out[12] += dd; out[13] += (int64_t)D.succ.size(); out[14] += D.remote_operands;
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
    int64_t deps = 0, maxdeps = 0, gemm = 0, lu = 0, subs = 0;
    for (const soglu::Task& t : G.tasks) {
        deps += t.n_deps; maxdeps = std::max<int64_t>(maxdeps, t.n_deps);
        gemm += t.type == soglu::T_GEMM; lu += t.type == soglu::T_LU; subs += t.type == soglu::T_SUB;
    }
    int64_t v[16] = {(int64_t)G.tasks.size(), (int64_t)G.pairs.size(), (int64_t)G.succ.size(), (int64_t)G.initial.size(), G.n_slots, G.n_levels,
                     G.fused_subs, G.fused_invs, G.aliased_invs, G.split_tasks, (int64_t)G.seg_begin.size() - 1, deps, maxdeps, gemm, lu,
                     (int64_t)(dt * 1e6)};
    for (int i = 0; i < 16; i++) out[i] = v[i];
    (void)subs;
    return SOGLU_OK;
}
