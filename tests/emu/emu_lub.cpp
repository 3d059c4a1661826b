// TEST INFRASTRUCTURE: host run of the blocked diagonal-block kernel (csrc/device/lu_blocked.cuh).  The device code
// is compiled unchanged for the host and executed by 256 real threads, one per CUDA math thread: __syncwarp and the
// CTA barrier are pthread barriers, warp shuffles and the m8n8k4 DMMA are emulated through per-warp exchange arrays
// with the PTX fragment layout (lane 4g+t holds A[g][t], B[t][g], C[g][2t..2t+1]); the arrive-only side of barrier Y
// and the shared-memory pivot counter (release / acquire) are host atomics.  Checked against a plain
// no-pivoting LU with the reference's pivot clamp (MatrixStdDouble.cpp:2745) and against L^-1 L = I, U U^-1 = I.
#define SOGLU_LUB_HOST 1
#include <pthread.h>

#include <atomic>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>

#include "../../sparse-operator-graph-lu_b200/csrc/device/lu_blocked.cuh"

namespace {
std::vector<double>* g_scr_ptr = nullptr;
pthread_barrier_t g_cta, g_warp[8];
double g_xa[8][32], g_xb[8][32];
thread_local int tl_ct = 0;
thread_local unsigned long long tl_rng = 88172645463325252ull;
// random yields shake the interleaving of the 256 host threads, so that an access that is only safe because of
// lock-step execution (a missing barrier in the device code) shows up as a wrong result here
inline void jitter() {
    tl_rng ^= tl_rng << 13; tl_rng ^= tl_rng >> 7; tl_rng ^= tl_rng << 17;
    if ((tl_rng & 15) == 0) std::this_thread::yield();
}
}  // namespace

namespace soglu { namespace lub { namespace hw {
void sync_warp() { jitter(); pthread_barrier_wait(&g_warp[tl_ct >> 5]); jitter(); }
void sync_math() { jitter(); pthread_barrier_wait(&g_cta); jitter(); }
double shfl(double v, int src) {
    const int w = tl_ct >> 5, l = tl_ct & 31;
    g_xa[w][l] = v;
    sync_warp();
    const double r = g_xa[w][src];
    sync_warp();
    return r;
}
void dmma(double& c0, double& c1, double a, double b) {
    const int w = tl_ct >> 5, l = tl_ct & 31, g = l >> 2, t = l & 3;
    g_xa[w][l] = a;
    g_xb[w][l] = b;
    sync_warp();
    for (int k = 0; k < 4; k++) {
        c0 = std::fma(g_xa[w][4 * g + k], g_xb[w][4 * (2 * t) + k], c0);
        c1 = std::fma(g_xa[w][4 * g + k], g_xb[w][4 * (2 * t + 1) + k], c1);
    }
    sync_warp();
}
double rcp(double x) { return 1.0 / x; }
double neg_rcp(double x) { return -1.0 / x; }
// barrier Y: 32 threads (warp 0) arrive and go on, 224 wait for all 256 (PTX bar.arrive / bar.sync on one barrier)
static std::atomic<int> y_count{0}, y_gen{0};
static bool y_enter() {
    if (y_count.fetch_add(1) + 1 == 256) { y_count.store(0); y_gen.fetch_add(1); return true; }
    return false;
}
void arrive_y() { jitter(); y_enter(); }
void sync_y() {
    jitter();
    const int g = y_gen.load();
    if (!y_enter()) while (y_gen.load() == g) std::this_thread::yield();
    jitter();
}
// an mbarrier with one arrival per phase = a counter of completed phases; the phase with parity p has completed
// iff (count & 1) != p, which is also what the hardware answers on a fresh barrier asked for parity 1
void pivot_init(unsigned long long* b) { __atomic_store_n(b, 0ull, __ATOMIC_RELAXED); }
void pivot_init_done() {}
void pivot_signal(unsigned long long* b) { jitter(); __atomic_fetch_add(b, 1ull, __ATOMIC_RELEASE); }
bool pivot_ready(const unsigned long long* b, int parity) { std::this_thread::yield(); return (int)(__atomic_load_n(b, __ATOMIC_ACQUIRE) & 1) != parity; }
void prof(int) {}
}}}  // namespace soglu::lub::hw

using namespace soglu::lub;

template <bool WITH_INV>
static void run(const std::vector<double>& A, std::vector<double>& S, std::vector<double>& W) {
    S.assign(64 * LD, 0.0);
    W.assign(64 * LD, 7.7e300);          // poison: the kernel has to clear it
    std::vector<double>& scr = *g_scr_ptr;    // persists across calls like the CTA's shared memory (barrier phases!)
    for (int i = 0; i < 64; i++) for (int j = 0; j < 64; j++) S[i * LD + j] = A[i * 64 + j];
    std::vector<std::thread> th;
    for (int ct = 0; ct < 256; ct++)
        th.emplace_back([&, ct]() { tl_ct = ct; tl_rng += 977u * ct; lu_blocked<WITH_INV, false>(S.data(), WITH_INV ? W.data() : nullptr, scr.data(), ct); });
    for (auto& t : th) t.join();
}

// Cholesky variant: the sweep with lltdcmpSimple's clamp, no U^-1
static void run_llt(const std::vector<double>& A, std::vector<double>& S, std::vector<double>& W) {
    S.assign(64 * LD, 0.0);
    W.assign(64 * LD, 7.7e300);
    std::vector<double>& scr = *g_scr_ptr;
    for (int i = 0; i < 64; i++) for (int j = 0; j < 64; j++) S[i * LD + j] = A[i * 64 + j];
    std::vector<std::thread> th;
    for (int ct = 0; ct < 256; ct++)
        th.emplace_back([&, ct]() { tl_ct = ct; lu_blocked<true, true, false>(S.data(), W.data(), scr.data(), ct); });
    for (auto& t : th) t.join();
}

static double clampLU(double p) { return (p < 1e-9 && p > -1e-9) ? ((p < 0) ? -1e-9 : 1e-9) : p; }
static void ref_lu(std::vector<double> a, std::vector<double>& out) {
    for (int k = 0; k < 64; k++) {
        const double p = clampLU(a[k * 64 + k]);
        a[k * 64 + k] = p;
        for (int i = k + 1; i < 64; i++) {
            const double l = a[i * 64 + k] / p;
            a[i * 64 + k] = l;
            for (int j = k + 1; j < 64; j++) a[i * 64 + j] -= l * a[k * 64 + j];
        }
    }
    out = a;
}
static double maxabs_prod_minus_eye(const std::vector<double>& X, const std::vector<double>& Y) {
    double m = 0;
    for (int i = 0; i < 64; i++) for (int j = 0; j < 64; j++) {
        double s = 0;
        for (int k = 0; k < 64; k++) s += X[i * 64 + k] * Y[k * 64 + j];
        m = std::fmax(m, std::fabs(s - (i == j ? 1.0 : 0.0)));
    }
    return m;
}

int main() {
    pthread_barrier_init(&g_cta, nullptr, 256);
    for (auto& b : g_warp) pthread_barrier_init(&b, nullptr, 32);
    std::vector<double> scr_store(SCRATCH_DOUBLES, 3.3e300);
    g_scr_ptr = &scr_store;
    {
        std::vector<std::thread> th;
        for (int ct = 0; ct < 256; ct++) th.emplace_back([&, ct]() { tl_ct = ct; lu_setup(scr_store.data(), ct); });
        for (auto& t : th) t.join();
    }
    unsigned long long st = 4242;
    auto rnd = [&]() { st = st * 6364136223846793005ull + 1442695040888963407ull; return ((st >> 11) * (1.0 / 9007199254740992.0)) * 2 - 1; };
    double worst_lu = 0, worst_li = 0, worst_ui = 0, worst_noinv = 0, worst_clamped = 0;
    for (int trial = 0; trial < 6; trial++) {
        std::vector<double> A(4096);
        for (auto& v : A) v = rnd();
        for (int i = 0; i < 64; i++) A[i * 64 + i] += (trial == 3 ? 6.0 : 40.0);       // trial 3: barely dominant
        if (trial == 4) { A[0] = 0.0; A[21 * 64 + 21] = 1e-12; }                        // clamp at pivot 0 and inside a panel
        if (trial == 5) for (int j = 0; j < 64; j++) A[33 * 64 + j] = A[32 * 64 + j];   // exact zero pivot after elimination
        std::vector<double> S, W, ref;
        run<true>(A, S, W);
        ref_lu(A, ref);
        double e = 0, scale = 0;
        for (int i = 0; i < 64; i++) for (int j = 0; j < 64; j++) {
            e = std::fmax(e, std::fabs(S[i * LD + j] - ref[i * 64 + j]));
            scale = std::fmax(scale, std::fabs(ref[i * 64 + j]));
        }
        if (trial < 4) { worst_lu = std::fmax(worst_lu, e / scale); std::printf("trial %d: factors vs plain LU %.3e\n", trial, e / scale); }
        std::vector<double> L(4096, 0), U(4096, 0), Li(4096, 0), Ui(4096, 0);
        for (int i = 0; i < 64; i++) for (int j = 0; j < 64; j++) {
            L[i * 64 + j] = j < i ? S[i * LD + j] : (i == j ? 1.0 : 0.0);
            U[i * 64 + j] = j >= i ? S[i * LD + j] : 0.0;
            Li[i * 64 + j] = j < i ? W[i * LD + j] : (i == j ? 1.0 : 0.0);
            Ui[i * 64 + j] = j >= i ? W[i * LD + j] : 0.0;
        }
        if (trial < 4) {       // (the clamped cases are numerically singular: L U = A is checked instead)
            worst_li = std::fmax(worst_li, maxabs_prod_minus_eye(Li, L));
            worst_ui = std::fmax(worst_ui, maxabs_prod_minus_eye(U, Ui));
        } else {
            // L U must reproduce A except in the clamped pivots' positions; compare with the reference's product
            double d = 0, sc = 0;
            for (int i = 0; i < 64; i++) for (int j = 0; j < 64; j++) {
                double s = 0, s2 = 0;
                for (int k = 0; k < 64; k++) {
                    s += L[i * 64 + k] * U[k * 64 + j];
                    const double rl = k < i ? ref[i * 64 + k] : (k == i ? 1.0 : 0.0), ru = j >= k ? ref[k * 64 + j] : 0.0;
                    s2 += rl * ru;
                }
                d = std::fmax(d, std::fabs(s - s2));
                sc = std::fmax(sc, std::fabs(s2));
            }
            // a clamped pivot of 1e-9 amplifies the rounding differences of the two schedules by ~1e9
            worst_clamped = std::fmax(worst_clamped, d / sc);
            std::printf("trial %d (clamped): product diff %.3e of %.3e\n", trial, d, sc);
        }
        std::vector<double> S2, W2;
        run<false>(A, S2, W2);
        for (int i = 0; i < 64; i++) for (int j = 0; j < 64; j++) worst_noinv = std::fmax(worst_noinv, std::fabs(S2[i * LD + j] - S[i * LD + j]));
    }
    // symmetric positive definite block: chol = L1 sqrt(D), chol^-1 = D^-1/2 L1^-1 as llt_task_blocked writes them
    double worst_chol = 0, worst_cinv = 0;
    for (int trial = 0; trial < 2; trial++) {
        std::vector<double> A(4096), Sm(4096), S, W;
        for (auto& v : A) v = rnd();
        for (int i = 0; i < 64; i++) for (int j = 0; j < 64; j++) { double s = 0; for (int k = 0; k < 64; k++) s += A[i * 64 + k] * A[j * 64 + k]; Sm[i * 64 + j] = s / 64 + (i == j); }
        run_llt(Sm, S, W);
        std::vector<double> C(4096, 0), Ci(4096, 0);
        for (int i = 0; i < 64; i++) for (int j = 0; j <= i; j++) {
            C[i * 64 + j] = (j < i ? S[i * LD + j] : 1.0) * std::sqrt(S[j * LD + j]);
            Ci[i * 64 + j] = (j < i ? W[i * LD + j] : 1.0) / std::sqrt(S[i * LD + i]);
        }
        for (int i = 0; i < 64; i++) for (int j = i + 1; j < 64; j++) if (W[i * LD + j] != 0.0) worst_cinv = 1.0;   // no U^-1 was asked for
        double r = 0;
        for (int i = 0; i < 64; i++) for (int j = 0; j < 64; j++) { double s = 0; for (int k = 0; k < 64; k++) s += C[i * 64 + k] * C[j * 64 + k]; r = std::fmax(r, std::fabs(s - Sm[i * 64 + j])); }
        worst_chol = std::fmax(worst_chol, r);
        worst_cinv = std::fmax(worst_cinv, maxabs_prod_minus_eye(Ci, C));
    }
    std::printf("chol %.3e  chol_inv %.3e\n", worst_chol, worst_cinv);
    std::printf("lu_vs_ref %.3e  clamped %.3e  Linv %.3e  Uinv %.3e  noinv_diff %.3e\n", worst_lu, worst_clamped, worst_li, worst_ui, worst_noinv);
    const bool ok = worst_lu < 1e-13 && worst_clamped < 1e-7 && worst_li < 1e-13 && worst_ui < 1e-13 && worst_noinv == 0.0 && worst_chol < 1e-13 && worst_cinv < 1e-13;
    std::printf(ok ? "OK\n" : "FAIL\n");
    return ok ? 0 : 1;
}
