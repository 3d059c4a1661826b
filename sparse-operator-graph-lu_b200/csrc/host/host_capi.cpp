// Section B of include/soglu.h: the planned problem as an opaque handle.
// Mirrors the front half of SOGLU::solveLU (solver.cpp:121-163) and decompose_solveLU
// up to the point where the reference calls BlockPlanner::calculate (solver.cpp:50-100).
#include "problem.h"
#ifdef _OPENMP
#include <omp.h>
#endif

#include <chrono>
#include <new>
#include <stdexcept>
#include <cstdio>
#include <cstring>
#include <sstream>

namespace soglu {

thread_local std::string g_last_error;
void set_error(const std::string& s) { g_last_error = s; }

static double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int prepare_problem(Problem& P, int dim, int64_t nnz, bool symmetric, const int* ii, const int* jj, const double* vv,
                    const double* b) {
    double t0 = now_s();
    P.dim = dim;
    P.symmetric = symmetric;
    P.a.n = dim;
    P.a.symmetric = symmetric;
    P.a.i.assign(ii, ii + nnz);
    P.a.j.assign(jj, jj + nnz);
    P.a.v.assign(vv, vv + nnz);
    P.b.assign(dim, 1.0);
    if (b) P.b.assign(b, b + dim);

    // mirror symmetric entries right after their source entry (solver.cpp:136-149)
    std::vector<int>& pi = P.pi;
    std::vector<int>& pj = P.pj;
    std::vector<double>& pv = P.pv;
    pi.clear(); pj.clear(); pv.clear();
    pi.reserve(symmetric ? 2 * nnz : nnz); pj.reserve(pi.capacity()); pv.reserve(pi.capacity());
    for (int64_t k = 0; k < nnz; k++) {
        if (ii[k] < 0 || ii[k] >= dim || jj[k] < 0 || jj[k] >= dim) { set_error("COO index out of range"); return SOGLU_ERR_ARG; }
        pi.push_back(ii[k]); pj.push_back(jj[k]); pv.push_back(vv[k]);
        if (symmetric && ii[k] != jj[k]) { pi.push_back(jj[k]); pj.push_back(ii[k]); pv.push_back(vv[k]); }
    }
    P.cfg.set(dim);
    P.n_ext = P.cfg.blockRows * P.cfg.blockSize;
    // rhs padded with 1.0 (solver.cpp:155-161); gps_reorder rebuilds it permuted
    P.b_perm = P.b;
    gps_reorder(dim, pi, pj, P.b_perm, P.cfg, P.ord);
    sort_in_block(dim, pi, pj, pv, P.b_perm, P.ord);
    double t1 = now_s();
    P.t_reorder = t1 - t0;
    std::ostringstream lg;
    lg << "GGPS reorder: levels: " << P.ord.levels << " bandwidth: " << P.ord.width << " last level count: "
       << P.ord.lastLevelCount << " total accounted: " << P.ord.accounted << " start from " << P.ord.startNode << "\n";
    lg << "re Order time: " << P.t_reorder << "\n";
    int rc = build_plan(P.cfg, symmetric, pi, pj, pv, P.plan, true);
    P.t_plan = now_s() - t1;
    P.log = lg.str() + P.plan.log;
    if (rc) { set_error(P.plan.log); return SOGLU_ERR_PLAN; }
    {
        std::ostringstream l2;
        l2 << "plan time: " << P.t_plan << "\n";
        P.log += l2.str();
    }
    P.flops = factor_flops(P.plan.ops);
    return SOGLU_OK;
}

}  // namespace soglu

using soglu::Problem;

extern "C" {

const char* soglu_last_error(void) { return soglu::g_last_error.c_str(); }
int soglu_abi_version(void) { return SOGLU_ABI_VERSION; }

int soglu_problem_from_coo(int32_t dim, int64_t nnz, int symmetric, const int32_t* index_i, const int32_t* index_j,
                           const double* vals, const double* b, soglu_problem** out) {
    if (!out || dim <= 0 || nnz < 0 || !index_i || !index_j || !vals) { soglu::set_error("bad argument"); return SOGLU_ERR_ARG; }
    Problem* P = nullptr;
    try {
        P = new Problem();
        int rc = soglu::prepare_problem(*P, dim, nnz, symmetric != 0, index_i, index_j, vals, b);
        if (rc) { delete P; return rc; }
    } catch (const std::bad_alloc&) {     // no exception may cross the C ABI
        delete P;
        soglu::set_error("out of host memory while planning");
        return SOGLU_ERR_OOM;
    } catch (const std::exception& e) {
        delete P;
        soglu::set_error(std::string("internal error: ") + e.what());
        return SOGLU_ERR_PLAN;
    }
    *out = reinterpret_cast<soglu_problem*>(P);
    return SOGLU_OK;
}

int soglu_problem_from_mtx(const char* path, soglu_problem** out) {
    if (!path || !out) { soglu::set_error("bad argument"); return SOGLU_ERR_ARG; }
    std::string fname(path);
    size_t pos = fname.find(".mtx");
    if (pos == std::string::npos) { soglu::set_error("usage: ./solve filename.mtx"); return SOGLU_ERR_ARG; }
    soglu::Coo a;
    std::vector<double> b;
    try {
        if (soglu::read_mtx(fname, a) == 0) { soglu::set_error("Can not open file"); return SOGLU_ERR_IO; }
        soglu::read_array(fname.substr(0, pos) + "_b.mtx", a.n, b);
    } catch (const std::exception& e) {
        soglu::set_error(std::string("reading the matrix failed: ") + e.what());
        return SOGLU_ERR_IO;
    }
    return soglu_problem_from_coo(a.n, (int64_t)a.v.size(), a.symmetric, a.i.data(), a.j.data(), a.v.data(), b.data(), out);
}

void soglu_problem_free(soglu_problem* p) { delete reinterpret_cast<Problem*>(p); }

int64_t soglu_problem_size(const soglu_problem* pp, const char* what) {
    const Problem* p = reinterpret_cast<const Problem*>(pp);
    if (!p || !what) return -1;
    std::string w(what);
    if (w == "dim") return p->dim;
    if (w == "nnz") return (int64_t)p->a.v.size();
    if (w == "nnz_expanded") return (int64_t)p->pv.size();
    if (w == "n_ext") return p->n_ext;
    if (w == "block_rows") return p->cfg.blockRows;
    if (w == "block_rows_l2") return p->cfg.blockRowsL2;
    if (w == "block_size_l2") return p->cfg.blockSizeL2;
    if (w == "storage") return p->plan.storage;
    if (w == "n_ops") return (int64_t)p->plan.ops.size();
    if (w == "fine_emitted") return p->plan.fine_emitted;
    if (w == "n_input") return (int64_t)p->plan.inputs.size();
    if (w == "n_entries") return (int64_t)p->plan.entry_val.size();
    if (w == "n_L") return (int64_t)p->plan.L.size();
    if (w == "n_U") return (int64_t)p->plan.U.size();
    if (w == "coarse_ops") return (int64_t)p->plan.coarse_ops.size();
    if (w == "coarse_emitted") return p->plan.coarse_emitted;
    if (w == "coarse_storage") return p->plan.coarse_storage;
    if (w == "symmetric") return p->symmetric;
    if (w == "gps_levels") return p->ord.levels;
    if (w == "gps_width") return p->ord.width;
    if (w == "gps_start") return p->ord.startNode;
    if (w == "gps_last") return p->ord.lastLevelCount;
    if (w == "max_stage") return p->plan.ops.empty() ? 0 : p->plan.ops.back().stage;
    return -1;
}

static void pack_ops(const soglu::OpVec& ops, int32_t* out) {
#pragma omp parallel for schedule(static)
    for (size_t k = 0; k < ops.size(); k++) {
        const soglu::Op& o = ops[k];
        int32_t* r = out + 8 * k;
        r[0] = o.op; r[1] = o.src; r[2] = o.src2; r[3] = o.result; r[4] = o.result2; r[5] = o.stage; r[6] = o.group; r[7] = o.seq;
    }
}
static void pack_refs(const std::vector<soglu::BlockRef>& v, int32_t* out) {
    for (size_t k = 0; k < v.size(); k++) { out[3 * k] = v[k].id; out[3 * k + 1] = v[k].brow; out[3 * k + 2] = v[k].bcol; }
}

int soglu_problem_get_i32(const soglu_problem* pp, const char* what, int32_t* out) {
    const Problem* p = reinterpret_cast<const Problem*>(pp);
    if (!p || !what || !out) { soglu::set_error("bad argument"); return SOGLU_ERR_ARG; }
    std::string w(what);
    auto cp = [&](const std::vector<int32_t>& v) { std::memcpy(out, v.data(), v.size() * sizeof(int32_t)); return (int)SOGLU_OK; };
    if (w == "perm_new2old") return cp(p->ord.newOrder);
    if (w == "perm_old2new") return cp(p->ord.reverseOrder);
    if (w == "stage") return cp(p->plan.stage);
    if (w == "laststage") return cp(p->plan.laststage);
    if (w == "block_row") return cp(p->plan.brow);
    if (w == "block_col") return cp(p->plan.bcol);
    if (w == "entry_block") { std::memcpy(out, p->plan.entry_block.data(), p->plan.entry_block.size() * 4); return SOGLU_OK; }
    if (w == "entry_pos") { std::memcpy(out, p->plan.entry_pos.data(), p->plan.entry_pos.size() * 4); return SOGLU_OK; }
    if (w == "perm_i") return cp(p->pi);
    if (w == "perm_j") return cp(p->pj);
    if (w == "ops") { pack_ops(p->plan.ops, out); return SOGLU_OK; }
    if (w == "coarse_ops") { pack_ops(p->plan.coarse_ops, out); return SOGLU_OK; }
    if (w == "inputs") { pack_refs(p->plan.inputs, out); return SOGLU_OK; }
    if (w == "L") { pack_refs(p->plan.L, out); return SOGLU_OK; }
    if (w == "U") { pack_refs(p->plan.U, out); return SOGLU_OK; }
    soglu::set_error("unknown array: " + w);
    return SOGLU_ERR_ARG;
}

int soglu_problem_get_f64(const soglu_problem* pp, const char* what, double* out) {
    const Problem* p = reinterpret_cast<const Problem*>(pp);
    if (!p || !what || !out) { soglu::set_error("bad argument"); return SOGLU_ERR_ARG; }
    std::string w(what);
    auto cp = [&](const std::vector<double>& v) { std::memcpy(out, v.data(), v.size() * sizeof(double)); return (int)SOGLU_OK; };
    if (w == "b") return cp(p->b);
    if (w == "b_perm") return cp(p->b_perm);
    if (w == "input_vals") {   // dense view of the entry list, built on demand
        soglu::BlockValues V;
        if (!p->plan.dense_inputs(V)) { soglu::set_error("out of host memory"); return SOGLU_ERR_OOM; }
        std::memcpy(out, V.data(), V.size() * sizeof(double));
        return SOGLU_OK;
    }
    if (w == "entry_val") { std::memcpy(out, p->plan.entry_val.data(), p->plan.entry_val.size() * sizeof(double)); return SOGLU_OK; }
    if (w == "flops") { out[0] = p->flops; return SOGLU_OK; }
    if (w == "t_reorder") { out[0] = p->t_reorder; return SOGLU_OK; }
    if (w == "t_plan") { out[0] = p->t_plan; return SOGLU_OK; }
    soglu::set_error("unknown array: " + w);
    return SOGLU_ERR_ARG;
}

const char* soglu_problem_log(const soglu_problem* pp) {
    const Problem* p = reinterpret_cast<const Problem*>(pp);
    return p ? p->log.c_str() : "";
}

void soglu_free(void* p) { std::free(p); }

int soglu_set_host_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n > 0 ? n : omp_get_num_procs());
    return omp_get_max_threads();
#else
    (void)n;
    return 1;
#endif
}

// Synthetic stencil matrices: natural lexicographic numbering with x
// fastest, rows emitted in order with ascending columns, %.17g, rhs 1 + 0.25*(i mod 7).
int soglu_write_stencil_mtx(const char* kind, int nx, int ny, int nz, int symmetric, const char* path) {
    if (!kind || !path || nx <= 0) { soglu::set_error("bad argument"); return SOGLU_ERR_ARG; }
    std::string k(kind), fname(path);
    size_t pos = fname.find(".mtx");
    if (pos == std::string::npos) { soglu::set_error("path must contain .mtx"); return SOGLU_ERR_ARG; }
    int dims;
    double diag;
    bool corners;
    if (k == "lap2d") { dims = 2; diag = 4; corners = false; }
    else if (k == "nine2d") { dims = 2; diag = 8; corners = true; }
    else if (k == "lap3d") { dims = 3; diag = 6; corners = false; }
    else { soglu::set_error("unknown stencil kind"); return SOGLU_ERR_ARG; }
    if (ny <= 0) ny = nx;
    if (dims == 2) nz = 1; else if (nz <= 0) nz = nx;
    const int64_t n = (int64_t)nx * ny * nz;
    // neighbour offsets in ascending linear-index order
    struct Off { int dx, dy, dz; };
    std::vector<Off> offs;
    for (int dz = -1; dz <= 1; dz++)
        for (int dy = -1; dy <= 1; dy++)
            for (int dx = -1; dx <= 1; dx++) {
                int nzc = (dx != 0) + (dy != 0) + (dz != 0);
                if (dims == 2 && dz != 0) continue;
                if (nzc == 0) { offs.push_back({0, 0, 0}); continue; }
                if (corners ? (dims == 2) : (nzc == 1)) offs.push_back({dx, dy, dz});
            }
    int64_t nnz = 0;
    for (int pass = 0; pass < 2; pass++) {
        FILE* fp = nullptr;
        if (pass == 1) {
            fp = std::fopen(fname.c_str(), "w");
            if (!fp) { soglu::set_error("cannot write " + fname); return SOGLU_ERR_IO; }
            std::fprintf(fp, "%%%%MatrixMarket matrix coordinate real %s\n%lld %lld %lld\n", symmetric ? "symmetric" : "general",
                         (long long)n, (long long)n, (long long)nnz);
        }
        for (int z = 0; z < nz; z++)
            for (int y = 0; y < ny; y++)
                for (int x = 0; x < nx; x++) {
                    int64_t i = ((int64_t)z * ny + y) * nx + x;
                    for (const Off& o : offs) {
                        int xx = x + o.dx, yy = y + o.dy, zz = z + o.dz;
                        if (xx < 0 || xx >= nx || yy < 0 || yy >= ny || zz < 0 || zz >= nz) continue;
                        int64_t j = ((int64_t)zz * ny + yy) * nx + xx;
                        if (symmetric && j > i) continue;
                        if (pass == 0) nnz++;
                        else std::fprintf(fp, "%lld %lld %.17g\n", (long long)(i + 1), (long long)(j + 1), i == j ? diag : -1.0);
                    }
                }
        if (fp) std::fclose(fp);
    }
    FILE* fb = std::fopen((fname.substr(0, pos) + "_b.mtx").c_str(), "w");
    if (!fb) { soglu::set_error("cannot write rhs file"); return SOGLU_ERR_IO; }
    std::fprintf(fb, "%%%%MatrixMarket matrix array real general\n%lld 1\n", (long long)n);
    for (int64_t i = 0; i < n; i++) std::fprintf(fb, "%.17g\n", 1.0 + 0.25 * (double)(i % 7));
    std::fclose(fb);
    return SOGLU_OK;
}

}  // extern "C"

// the raw MatrixMarket reader (tests): entries in file order, meta = {dimension, symmetric}; returns the entry count
extern "C" int64_t soglu_debug_read_mtx(const char* path, int64_t cap, int32_t* i, int32_t* j, double* v, int64_t* meta) {
    soglu::Coo c;
    const long n = soglu::read_mtx(path ? path : "", c);
    if (meta) { meta[0] = c.n; meta[1] = c.symmetric ? 1 : 0; }
    for (long k = 0; k < n && k < cap; k++) { i[k] = c.i[k]; j[k] = c.j[k]; v[k] = c.v[k]; }
    return n;
}
