// TEST INFRASTRUCTURE (not part of the product): executes a compiled TaskGraph of the persistent executor on the
// host with plain dense 64x64 loops, one task after the other in task order.  It checks, without a GPU, that what
// the task compiler emits (fused sub / inverse tasks, aliased inverses, row slices, pool slot recycling, chain cuts)
// still computes the factors the reference's operation list defines -- the tests compare the L and U blocks it
// returns with the oracle's.  Block semantics as in the executor (csrc/device/executor.cu) and the reference:
// lu without pivoting, unit L, |u_kk| < 1e-9 clamped sign-preserving (ludcmpSimple, MatrixStdDouble.cpp:2711-2784);
// llt with the pivot < 1e-20 clamp (lltdcmpSimple, 2629-2668); inverses by substitution (2787-2866).
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "../../sparse-operator-graph-lu_b200/csrc/device/tasks.h"

using namespace soglu;

namespace {
constexpr int N = 64, NN = N * N;

void lu_nopivot(const double* a, double* l, double* u) {
    double w[NN];
    std::memcpy(w, a, sizeof w);
    for (int k = 0; k < N; k++) {
        double p = w[k * N + k];
        if (p < 1e-9 && p > -1e-9) p = (p < 0) ? -1e-9 : 1e-9;
        w[k * N + k] = p;
        for (int i = k + 1; i < N; i++) {
            const double m = w[i * N + k] / p;
            w[i * N + k] = m;
            for (int j = k + 1; j < N; j++) w[i * N + j] -= m * w[k * N + j];
        }
    }
    for (int i = 0; i < N; i++)
        for (int j = 0; j < N; j++) {
            l[i * N + j] = j < i ? w[i * N + j] : (j == i ? 1.0 : 0.0);
            u[i * N + j] = j >= i ? w[i * N + j] : 0.0;
        }
}
void llt(const double* a, double* l) {
    std::memset(l, 0, NN * sizeof(double));
    for (int k = 0; k < N; k++) {
        double p = a[k * N + k];
        for (int j = 0; j < k; j++) p -= l[k * N + j] * l[k * N + j];
        if (p < 1e-20) p = 1e-20;
        const double d = std::sqrt(p);
        l[k * N + k] = d;
        for (int i = k + 1; i < N; i++) {
            double s = a[i * N + k];
            for (int j = 0; j < k; j++) s -= l[i * N + j] * l[k * N + j];
            l[i * N + k] = s / d;
        }
    }
}
void inv_lower(const double* l, double* y) {     // L Y = I, general diagonal
    std::memset(y, 0, NN * sizeof(double));
    for (int c = 0; c < N; c++)
        for (int i = c; i < N; i++) {
            double s = (i == c) ? 1.0 : 0.0;
            for (int k = c; k < i; k++) s -= l[i * N + k] * y[k * N + c];
            y[i * N + c] = s / l[i * N + i];
        }
}
void inv_upper(const double* u, double* y) {     // U Y = I
    std::memset(y, 0, NN * sizeof(double));
    for (int c = 0; c < N; c++)
        for (int i = c; i >= 0; i--) {
            double s = (i == c) ? 1.0 : 0.0;
            for (int k = i + 1; k <= c; k++) s -= u[i * N + k] * y[k * N + c];
            y[i * N + c] = s / u[i * N + i];
        }
}
}  // namespace

// split: bit 0 = row slices in narrow levels, bit 1 = also near-critical tasks (split_slack), bit 2 = 8 SMs, bit 3 = no static (latest-start) order.
// mode: 0 = default order key, m > 0 = order_alpha (m - 1) %.  max_slots > 0 forces slot recycling.  keep_out: n_keep dense 64x64 blocks.  stats = {tasks, segments, slots, chain cuts applied, row-split tasks}.
// grid = {pr, pc, nb} with brow / bcol per block id: the graph is compiled for pr*pc owners (2D block-cyclic squares of
// nb blocks, mirrors of remote blocks filled by fetch tasks) and every owner gets its own pool here.
extern "C" int emu_run(int64_t n_ids, int64_t n_input, const int32_t* input_ids, const double* input_dense, int64_t n_ops, const int32_t* src,
                       const int32_t* src2, const uint8_t* op, const int32_t* result, const int32_t* result2, int64_t n_keep,
                       const int32_t* keep_ids, int mode, double cut_max_slack_us, int64_t max_slots, int split, uint64_t order_seed, double* keep_out, int64_t* stats,
                       char* err_out, int err_len, const int32_t* grid, const int32_t* brow, const int32_t* bcol) {
    std::vector<int32_t> keep(keep_ids, keep_ids + n_keep);
    CompileOptions co;
    co.max_slots = max_slots;
    co.split_narrow = split & 1;
    if (split & 2) co.split_slack_us = 100.0;     // also split near-critical GEMM tasks of wide levels
    if (split & 8) co.static_order = false;       // keep the tasks in the order of the operation list
    if (split & 4) co.n_sms = 8;                  // pretend the GPU is small: the small test cases get wide levels too
    TaskGraph G;
    auto fail = [&](const std::string& e) { std::snprintf(err_out, err_len, "%s", e.c_str()); return 1; };
    std::vector<int8_t> owners;
    const int world = grid ? grid[0] * grid[1] : 1;
    if (world > 1) {
        owners.assign(n_ids, 0);
        for (int64_t id = 1; id < n_ids; id++)
            if (brow[id] >= 0 && bcol[id] >= 0) owners[id] = (int8_t)(((brow[id] / grid[2]) % grid[0]) * grid[1] + ((bcol[id] / grid[2]) % grid[1]));
        co.owner_of_id = owners.data();
        co.n_owners = world;
    }
    if (mode > 0) co.order_alpha = (mode - 1) / 100.0;
    (void)cut_max_slack_us;     // (kept in the signature: the chain-cut compile mode it belonged to was measured and removed)
    std::string err = compile_tasks(n_ids, n_input, input_ids, n_ops, src, src2, op, result, result2, keep, co, G);
    if (!err.empty()) return fail(err);
    // one pool per owner; slot 0 of each stays its zero block
    int64_t stride = 0;
    for (int64_t sl : G.slots_per_owner) stride = std::max(stride, sl);
    std::vector<double> pool((size_t)stride * world * NN, 0.0);
    auto blk = [&](int32_t ref) { return pool.data() + ((size_t)((uint32_t)ref >> REF_SHIFT) * stride + (size_t)(ref & REF_MASK)) * NN; };
    auto ref_of = [&](int32_t id) { return make_ref(G.owner_of[id], G.slot_of[id]); };
    for (int64_t k = 0; k < n_input; k++) std::memcpy(blk(ref_of(input_ids[k])), input_dense + k * NN, NN * sizeof(double));
    std::vector<double> acc(NN);
    // Execution order: the task order (seed 0), or -- like the executor -- whatever the dependency counters allow:
    // per segment, a seeded random pick from the ready set; a finishing task decrements every successor group once and
    // a group's slices become ready together.  A missing dependency edge then shows up as a wrong factor.
    std::vector<int32_t> order;
    order.reserve(G.tasks.size());
    if (order_seed == 0) {
        for (size_t t = 0; t < G.tasks.size(); t++) order.push_back((int32_t)t);
    } else {
        uint64_t rng = order_seed * 6364136223846793005ull + 1442695040888963407ull;
        std::vector<int32_t> dep(G.tasks.size());
        for (size_t t = 0; t < G.tasks.size(); t++) dep[t] = G.tasks[t].n_deps;
        for (size_t sg = 0; sg + 1 < G.seg_begin.size(); sg++) {
            std::vector<int32_t> ready(G.initial.begin() + G.seg_init[sg], G.initial.begin() + G.seg_init[sg + 1]);
            size_t ran = 0;
            while (!ready.empty()) {
                rng = rng * 6364136223846793005ull + 1442695040888963407ull;
                const size_t pick = (size_t)((rng >> 33) % ready.size());
                const int32_t t = ready[pick];
                ready[pick] = ready.back();
                ready.pop_back();
                order.push_back(t);
                ran++;
                for (int32_t e = G.tasks[t].succ_begin; e < G.tasks[t].succ_end; e++) {
                    const int32_t nx = G.succ[e];
                    if (--dep[nx] == 0)
                        for (int q = 0, g = task_group_size(G.tasks[nx]); q < g; q++) ready.push_back(nx + q);
                }
            }
            if (ran != (size_t)(G.seg_begin[sg + 1] - G.seg_begin[sg])) return fail("dependency counters do not release every task of a segment");
        }
    }
    for (const int32_t t : order) {
        const Task& T = G.tasks[t];
        const Pair* P = G.pairs.data() + T.pair_begin;
        switch (T.type) {
            case T_GEMM: {
                const int r0 = 16 * ((T.flags >> TF_ROW0_SHIFT) & 3), r1 = r0 + 16 * ((T.flags >> TF_NROWS_SHIFT) & 7);
                std::fill(acc.begin(), acc.end(), 0.0);
                for (int p = 0; p < T.n_pairs; p++) {
                    const double* a = blk(P[p].a);
                    const double* b = blk(P[p].b);
                    for (int i = r0; i < r1; i++)
                        for (int k = 0; k < N; k++) {
                            const double av = a[i * N + k];
                            if (T.flags & TF_TRANSB) { for (int j = 0; j < N; j++) acc[i * N + j] += av * b[j * N + k]; }
                            else { for (int j = 0; j < N; j++) acc[i * N + j] += av * b[k * N + j]; }
                        }
                }
                double* out = blk(T.out);
                const double* ini = (T.flags & TF_INIT) ? blk(T.init) : nullptr;
                for (int i = r0; i < r1; i++)
                    for (int j = 0; j < N; j++) {
                        const double v = (T.flags & TF_NEGATE) ? -acc[i * N + j] : acc[i * N + j];
                        out[i * N + j] = ini ? ini[i * N + j] + v : v;
                    }
                break;
            }
            case T_SUB: {
                const double* a = blk(P[0].a);
                const double* b = blk(P[0].b);
                double* out = blk(T.out);
                for (int i = 0; i < NN; i++) out[i] = a[i] - b[i];
                break;
            }
            case T_LU: {
                std::vector<double> l(NN), u(NN);
                lu_nopivot(blk(P[0].a), l.data(), u.data());
                std::memcpy(blk(T.out), l.data(), NN * sizeof(double));
                std::memcpy(blk(T.out2), u.data(), NN * sizeof(double));
                if (T.flags & TF_LINV) inv_lower(l.data(), blk(T.init));
                if (T.flags & TF_UINV) inv_upper(u.data(), blk(T.out4));
                break;
            }
            case T_LLT: {
                std::vector<double> l(NN);
                llt(blk(P[0].a), l.data());
                std::memcpy(blk(T.out), l.data(), NN * sizeof(double));
                if (T.flags & TF_LINV) inv_lower(l.data(), blk(T.init));
                break;
            }
            case T_LOWERINV: { std::vector<double> y(NN); inv_lower(blk(P[0].a), y.data()); std::memcpy(blk(T.out), y.data(), NN * sizeof(double)); break; }
            case T_UPPERINV: { std::vector<double> y(NN); inv_upper(blk(P[0].a), y.data()); std::memcpy(blk(T.out), y.data(), NN * sizeof(double)); break; }
            default: return fail("unknown task type");
        }
    }
    for (int64_t k = 0; k < n_keep; k++) {
        if (G.recycled[keep_ids[k]]) return fail("a kept block was recycled");
        std::memcpy(keep_out + k * NN, blk(ref_of(keep_ids[k])), NN * sizeof(double));
    }
    stats[0] = (int64_t)G.tasks.size(); stats[1] = (int64_t)G.seg_begin.size() - 1; stats[2] = stride; stats[3] = 0; stats[4] = G.split_tasks;
    stats[5] = 0;
    for (int64_t m : G.mirrors_per_owner) stats[5] += m;
    return 0;
}
