// TEST INFRASTRUCTURE ONLY -- never linked into the product library.
//
// Harness around the UNMODIFIED reference objects (built by oracle/Makefile from
// /root/reference/*.cpp into oracle/_ref/).  It runs the reference's own public
// entry point SOGLU::solveLU (solver.h:24) and captures, through GNU ld
// `--wrap` hooks on three BlockPlanner entry points, the intermediate state the
// parity tests need:
//
//   copyOperatorL2 (BlockPlanner.cpp:1332)  -> coarse (L2) op list after blockPlan
//   calculate      (BlockPlanner.cpp:376)   -> fine op list, stage/laststage,
//                                              input blocks (dense 64x64)
//   solve          (BlockPlanner.cpp:834)   -> L/U leaf coordinates, factor blocks,
//                                              permuted padded rhs
//
// and afterwards x (returned by solveLU, NOT the _x.mtx file -- main.cpp:61 writes
// data::b) and the permutation (GOrder::newOrder / reverseOrder, GPSOrder.cpp:448).
//
// usage: ref_harness <file.mtx> <outdir> [--blocks] [--lean]
//   --blocks     also dump dense values of every input block and every L/U block
//   --lean       timing + x only: skip the op-list / stage / factor-coordinate dumps (4.7 GB of op records
//                at 100^3); what bench.py uses for the CPU baseline on the full-size configs
// All dumps are raw little-endian arrays; meta.txt lists the counts.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <string>
#include <vector>
#include <chrono>
#include <iostream>
#include <memory>
#include <algorithm>

#include "operation.h"
#include "matrix.h"
#include "data.h"
#include "memutil.h"
#include "BlockPlanner.h"
#include "GPSOrder.h"
#include "config.h"
#include "mtx.h"
#include "solver.h"

using namespace SOGLU;

static std::string g_out;
static bool g_blocks = false, g_lean = false;
static FILE* g_meta = nullptr;
static double g_t_factor = 0, g_t_solve = 0;

static void write_raw(const std::string& name, const void* p, size_t bytes) {
    std::string f = g_out + "/" + name;
    FILE* fp = fopen(f.c_str(), "wb");
    if (!fp) { perror(f.c_str()); exit(2); }
    if (bytes) fwrite(p, 1, bytes, fp);
    fclose(fp);
}

static void dump_graph(const char* name) {
    size_t n = data::graph.size();
    std::vector<int32_t> rec(n * 8);
    for (size_t i = 0; i < n; i++) {
        const operation* o = data::graph[i];
        int32_t* r = &rec[i * 8];
        r[0] = (int)o->op; r[1] = o->src; r[2] = o->src2; r[3] = o->result; r[4] = o->result2;
        r[5] = o->stage;   r[6] = o->groupNum; r[7] = o->sequenceNum;
    }
    write_raw(name, rec.data(), rec.size() * sizeof(int32_t));
}

// dense 64x64 view of a 64x72 reference block, masked by the per-row detail bitmap
// (layout: const.h:19-34; same masking rule as mat_clean, MatrixStdDouble.cpp:161-190)
static void block_to_dense(const double* blk, double* out) {
    const uint16_t* det = (const uint16_t*)(blk + DETAILOFFSET);
    for (int r = 0; r < 64; r++) {
        uint16_t m = det[DETAILSKIPSHORT * (r / 32) + r % 32];
        for (int c = 0; c < 64; c++)
            out[r * 64 + c] = (m & (1u << (c / 8))) ? blk[r * BLOCKCOL + c] : 0.0;
    }
}

struct Leaf { int32_t id, brow, bcol; };
static void walk(matrix* m, int r0, int c0, int n, std::vector<Leaf>& out) {
    if (!m) return;
    if (m->level == 0) {
        if (m->blockindex > 0) out.push_back({(int32_t)m->blockindex, r0, c0});
        return;
    }
    int h = n / 2;
    for (int q = 0; q < 4; q++) walk(m->submatrix[q], r0 + (q >> 1) * h, c0 + (q & 1) * h, h, out);
}

static void dump_leaf_blocks(const char* name, const std::vector<Leaf>& lv) {
    std::string f = g_out + "/" + name;
    FILE* fp = fopen(f.c_str(), "wb");
    std::vector<double> d(4096);
    for (const Leaf& l : lv) {
        const double* p = data::blockstorage[l.id];
        if (p) block_to_dense(p, d.data()); else std::fill(d.begin(), d.end(), 0.0);
        fwrite(d.data(), sizeof(double), 4096, fp);
    }
    fclose(fp);
}

extern "C" {
// ---- hook 1: coarse plan is complete when copyOperatorL2 is entered (solver.cpp:90-92)
void __real__ZN5SOGLU12BlockPlanner14copyOperatorL2EPNS_6matrixES2_S2_S2_S2_i(matrix*, matrix*, matrix*, matrix*, matrix*, int);
void __wrap__ZN5SOGLU12BlockPlanner14copyOperatorL2EPNS_6matrixES2_S2_S2_S2_i(matrix* a, matrix* l, matrix* u, matrix* l2, matrix* u2, int n) {
    if (!g_lean) dump_graph("ops_coarse.i32");
    fprintf(g_meta, "coarse_ops %zu\ncoarse_storage %d\ncoarse_block_rows %d\ncoarse_block_size %d\n",
            data::graph.size(), data::storageCount, data::blockRows, data::blockSize);
    __real__ZN5SOGLU12BlockPlanner14copyOperatorL2EPNS_6matrixES2_S2_S2_S2_i(a, l, u, l2, u2, n);
}

// ---- hook 2: fine plan is complete when calculate is entered (solver.cpp:106)
void __real__ZN5SOGLU12BlockPlanner9calculateEv();
void __wrap__ZN5SOGLU12BlockPlanner9calculateEv() {
    std::vector<Leaf> in;
    walk(data::blocks, 0, 0, data::blockRows, in);
    if (!g_lean) {
        dump_graph("ops_fine.i32");
        write_raw("stage.i32", data::stage.get(), sizeof(int) * data::storageCount);
        write_raw("laststage.i32", data::laststage.get(), sizeof(int) * data::storageCount);
        write_raw("inputs.i32", in.data(), in.size() * sizeof(Leaf));
    }
    if (g_blocks) dump_leaf_blocks("inputs.f64", in);
    fprintf(g_meta, "fine_ops %zu\nstorage %d\nblock_rows %d\nn_input %zu\nmsize %d\nsymmetric %d\n",
            data::graph.size(), data::storageCount, data::blockRows, in.size(), data::mSize, (int)data::symmetric);
    auto t0 = std::chrono::steady_clock::now();
    __real__ZN5SOGLU12BlockPlanner9calculateEv();
    g_t_factor = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

// ---- hook 3: factors are complete when solve is entered (solver.cpp:115)
void __real__ZN5SOGLU12BlockPlanner5solveEPNS_6matrixES2_Pdi(matrix*, matrix*, double*, int);
void __wrap__ZN5SOGLU12BlockPlanner5solveEPNS_6matrixES2_Pdi(matrix* bl, matrix* bu, double* b, int n) {
    std::vector<Leaf> L, U;
    walk(bl, 0, 0, data::blockRows, L);
    walk(bu, 0, 0, data::blockRows, U);
    if (!g_lean) {
        write_raw("L.i32", L.data(), L.size() * sizeof(Leaf));
        write_raw("U.i32", U.data(), U.size() * sizeof(Leaf));
        write_raw("b_perm.f64", b, sizeof(double) * n);
    }
    if (g_blocks) { dump_leaf_blocks("L.f64", L); dump_leaf_blocks("U.f64", U); }
    fprintf(g_meta, "n_L %zu\nn_U %zu\nn_ext %d\n", L.size(), U.size(), n);
    auto t0 = std::chrono::steady_clock::now();
    __real__ZN5SOGLU12BlockPlanner5solveEPNS_6matrixES2_Pdi(bl, bu, b, n);
    g_t_solve = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (!g_lean) write_raw("x_perm.f64", data::x, sizeof(double) * data::mSize);
}
}

int main(int argc, char** argv) {
    if (argc < 3) { fprintf(stderr, "usage: ref_harness file.mtx outdir [--blocks]\n"); return 1; }
    std::string fname = argv[1];
    g_out = argv[2];
    for (int i = 3; i < argc; i++) {
        if (!strcmp(argv[i], "--blocks")) g_blocks = true;
        if (!strcmp(argv[i], "--lean")) g_lean = true;
    }
    g_meta = fopen((g_out + "/meta.txt").c_str(), "w");
    if (!g_meta) { perror("meta.txt"); return 2; }

    iniData();
    std::string base = fname.substr(0, fname.find(".mtx"));
    if (mtx::readMTX(fname) == 0) return 3;
    mtx::readArray(base + "_b.mtx", mtx::mSize);
    if (!g_lean) {
        write_raw("coo_i.i32", mtx::indexi, sizeof(int) * mtx::valcount);
        write_raw("coo_j.i32", mtx::indexj, sizeof(int) * mtx::valcount);
        write_raw("coo_v.f64", mtx::vals, sizeof(double) * mtx::valcount);
        write_raw("b.f64", mtx::b, sizeof(double) * mtx::mSize);
    }

    auto t0 = std::chrono::steady_clock::now();
    double* x = solveLU(mtx::mSize, mtx::valcount, mtx::symmetric, mtx::indexi, mtx::indexj, mtx::vals, mtx::b);
    double t_total = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    double err = mtx::checkResult(x);

    write_raw("x.f64", x, sizeof(double) * mtx::mSize);
    if (!g_lean) {
        write_raw("perm_new2old.i32", GOrder::newOrder, sizeof(int) * mtx::mSize);
        write_raw("perm_old2new.i32", GOrder::reverseOrder, sizeof(int) * mtx::mSize);
    }
    fprintf(g_meta, "dim %d\nnnz %d\nfile_symmetric %d\nmax_rhs_error %.17g\nt_factor %.6f\nt_solve %.6f\nt_total %.6f\n",
            mtx::mSize, mtx::valcount, (int)mtx::symmetric, err, g_t_factor, g_t_solve, t_total);
    fclose(g_meta);
    printf("HARNESS factor_s %.6f solve_s %.6f total_s %.6f max_rhs_error %.6g\n", g_t_factor, g_t_solve, t_total, err);
    return 0;
}
