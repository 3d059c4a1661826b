// TEST INFRASTRUCTURE ONLY -- the reference-side binding of INTEGRATION.md section 2, as real code.
//
// Compiled against the UNMODIFIED reference headers and linked with the unmodified reference objects (oracle/_ref/*.o)
// plus libsoglu_b200.so.  GNU ld --wrap redirects the reference's two hot-path calls in decompose_solveLU
// (solver.cpp:106 BlockPlanner::calculate(), solver.cpp:115 BlockPlanner::solve(bl2, bu2, b, n)) to the adapter, and a
// third hook on copyOperatorL2 (solver.cpp:90-92) only remembers the L/U quadtrees, which calculate() itself does not
// receive.  Everything else -- mtx reader, GPS ordering, both planner passes, iniBlockStorage, the un-permutation of
// x -- is the reference's own code, so tests/test_gpu_parity.py::test_reference_side_adapter proves the C ABI from the
// reference's side: reference front-end + this library's factor / solve must give the reference's x.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <string>
#include <vector>

#include "../include/soglu.h"
#include "operation.h"
#include "matrix.h"
#include "data.h"
#include "memutil.h"
#include "mtx.h"
#include "config.h"
#include "solver.h"

using namespace SOGLU;

// leaves of a block quadtree with their block coordinates: halving n from data::blockRows, exactly as lowerSolver does
// (BlockPlanner.cpp:763-767)
static void leaves(matrix* m, int r0, int c0, int n, std::vector<int32_t>& id, std::vector<int32_t>& br, std::vector<int32_t>& bc) {
    if (!m) return;
    if (m->level == 0) { if (m->blockindex > 0) { id.push_back((int32_t)m->blockindex); br.push_back(r0); bc.push_back(c0); } return; }
    for (int q = 0; q < 4; q++) leaves(m->submatrix[q], r0 + (q >> 1) * n / 2, c0 + (q & 1) * n / 2, n / 2, id, br, bc);
}

static soglu_ctx* g_ctx;
static matrix *g_bl2, *g_bu2;

static void die(const char* what) { std::cout << what << ": " << soglu_last_error() << std::endl; std::exit(4); }

// replaces BlockPlanner::calculate() at solver.cpp:106
static void soglu_calculate(matrix* bl2, matrix* bu2) {
    if (soglu_create(&g_ctx, 1, nullptr)) die("soglu_create");
    // input blocks: unpack the 64x72 layout (const.h:19-34) to dense 64x64, masking by the detail bitmap
    std::vector<int32_t> in_id, in_r, in_c;
    leaves(data::blocks, 0, 0, data::blockRows, in_id, in_r, in_c);
    std::vector<double> dense(in_id.size() * 4096);
    for (size_t k = 0; k < in_id.size(); k++) {
        const double* b = data::blockstorage[in_id[k]];
        const uint16_t* det = (const uint16_t*)(b + DETAILOFFSET);
        for (int r = 0; r < 64; r++)
            for (int c = 0; c < 64; c++)
                dense[k * 4096 + r * 64 + c] = ((det[DETAILSKIPSHORT * (r / 32) + r % 32] >> (c / 8)) & 1) ? b[r * BLOCKCOL + c] : 0.0;
    }
    if (soglu_set_blocks(g_ctx, data::storageCount, (int64_t)in_id.size(), in_id.data(), dense.data())) die("soglu_set_blocks");
    // operation list (operation.h:37-52) as parallel arrays, in data::graph order
    const size_t n = data::graph.size();
    std::vector<int32_t> s(n), s2(n), r(n), r2(n), st(n);
    std::vector<uint8_t> op(n);
    for (size_t i = 0; i < n; i++) {
        const operation* o = data::graph[i];
        s[i] = o->src; s2[i] = o->src2; r[i] = o->result; r2[i] = o->result2; st[i] = o->stage; op[i] = (uint8_t)o->op;
    }
    if (soglu_set_graph(g_ctx, (int64_t)n, s.data(), s2.data(), op.data(), r.data(), r2.data(), st.data(), nullptr, nullptr)) die("soglu_set_graph");
    std::vector<int32_t> li, lr, lc, ui, ur, uc;
    leaves(bl2, 0, 0, data::blockRows, li, lr, lc);
    leaves(bu2, 0, 0, data::blockRows, ui, ur, uc);
    if (soglu_set_factors(g_ctx, (int64_t)li.size(), li.data(), lr.data(), lc.data(), (int64_t)ui.size(), ui.data(), ur.data(), uc.data(),
                          data::blockRows, data::symmetric ? 1 : 0)) die("soglu_set_factors");
    soglu_stats fs;
    if (soglu_factor(g_ctx, &fs)) die("soglu_factor");
    std::printf("ADAPTER factor_s %.6f tasks %lld launches %lld\n", fs.seconds, (long long)fs.tasks, (long long)fs.kernel_launches);
}

// replaces BlockPlanner::solve() at solver.cpp:115
static void soglu_solve_adapter(double* b, int n) {
    std::vector<double> x(n);
    soglu_stats ss;
    if (soglu_solve(g_ctx, b, x.data(), &ss)) die("soglu_solve");
    data::x = (double*)memutil::getSmallMem(1, sizeof(double) * data::mSize);     // same ownership as BlockPlanner.cpp:850
    for (int i = 0; i < data::mSize; i++) data::x[i] = x[i];
    std::printf("ADAPTER solve_s %.6f\n", ss.seconds);
    soglu_destroy(g_ctx);
}

extern "C" {
void __real__ZN5SOGLU12BlockPlanner14copyOperatorL2EPNS_6matrixES2_S2_S2_S2_i(matrix*, matrix*, matrix*, matrix*, matrix*, int);
void __wrap__ZN5SOGLU12BlockPlanner14copyOperatorL2EPNS_6matrixES2_S2_S2_S2_i(matrix* a, matrix* l, matrix* u, matrix* l2, matrix* u2, int n) {
    g_bl2 = l2; g_bu2 = u2;
    __real__ZN5SOGLU12BlockPlanner14copyOperatorL2EPNS_6matrixES2_S2_S2_S2_i(a, l, u, l2, u2, n);
}
void __wrap__ZN5SOGLU12BlockPlanner9calculateEv() { soglu_calculate(g_bl2, g_bu2); }
void __wrap__ZN5SOGLU12BlockPlanner5solveEPNS_6matrixES2_Pdi(matrix*, matrix*, double* b, int n) { soglu_solve_adapter(b, n); }
}

// usage: ref_adapter <file.mtx> <out_x.f64>   (the reference's main.cpp flow, x written as raw doubles)
int main(int argc, char** argv) {
    if (argc < 3) { std::fprintf(stderr, "usage: ref_adapter file.mtx out_x.f64\n"); return 1; }
    const std::string fname = argv[1];
    iniData();
    if (mtx::readMTX(fname) == 0) return 3;
    mtx::readArray(fname.substr(0, fname.find(".mtx")) + "_b.mtx", mtx::mSize);
    double* x = solveLU(mtx::mSize, mtx::valcount, mtx::symmetric, mtx::indexi, mtx::indexj, mtx::vals, mtx::b);
    const double err = mtx::checkResult(x);
    FILE* fp = std::fopen(argv[2], "wb");
    if (!fp) return 2;
    std::fwrite(x, sizeof(double), mtx::mSize, fp);
    std::fclose(fp);
    std::printf("ADAPTER max_rhs_error %.6g\n", err);
    return 0;
}
