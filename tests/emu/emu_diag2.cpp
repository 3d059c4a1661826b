// Host emulation of the two-pivots-per-barrier diagonal-block kernel (csrc/device/diag2.cuh): the SAME per-thread
// code is run for all 256 (ty, tx) "threads", one barrier interval at a time, and checked against a plain
// no-pivoting LU / Cholesky / triangular inverse with the reference's pivot clamps.  Built and run by
// tests/test_diag2_emulation.py (CPU only; the GPU tests then check the real kernel against the oracle).
#define SOGLU_DIAG2_HOST 1
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../sparse-operator-graph-lu_b200/csrc/device/diag2.cuh"

using namespace soglu::diag2;

struct Regs { double a[4][4], wl[4][4], wu[4][4]; };

template <bool WITH_INV, bool WU, bool LLT>
static void run(const std::vector<double>& A, std::vector<double>& Aout, std::vector<double>& WL, std::vector<double>& WUt, std::vector<double>& ip) {
    std::vector<double> As(64 * LD, 0.0), xbuf(SCRATCH_DOUBLES, 1e300);   // poison: a read of an unpublished slot shows
    for (int i = 0; i < 64; i++) for (int j = 0; j < 64; j++) As[i * LD + j] = A[i * 64 + j];
    std::vector<Regs> R(256);
    for (int t = 0; t < 256; t++) init<WITH_INV, WU, LLT>(As.data(), xbuf.data(), R[t].a, R[t].wl, R[t].wu, t >> 4, t & 15);
    for (int kr = 0; kr < 4; kr++)
        for (int ko = 0; ko < 16; ko += 2)   // --- barrier ---
            for (int t = 0; t < 256; t++) eliminate2<WITH_INV, WU, LLT>(kr, ko, (ko >> 1) & 1, xbuf.data(), R[t].a, R[t].wl, R[t].wu, t >> 4, t & 15);
    Aout.assign(4096, 0); WL.assign(4096, 0); WUt.assign(4096, 0); ip.assign(64, 0);
    for (int t = 0; t < 256; t++)
        for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) {
            const int i = (t >> 4) + 16 * r, j = (t & 15) + 16 * c;
            Aout[i * 64 + j] = R[t].a[r][c];
            if (WITH_INV) { WL[i * 64 + j] = R[t].wl[r][c]; if (WU) WUt[i * 64 + j] = R[t].wu[r][c]; }
        }
    if (WITH_INV) for (int k = 0; k < 64; k++) ip[k] = xbuf[IPBUF + k];
}

static double clampLU(double p) { return (p < 1e-9 && p > -1e-9) ? ((p < 0) ? -1e-9 : 1e-9) : p; }

// plain elimination with the reference's clamp; a holds L (strict lower, unit diagonal implied) and U
static void ref_lu(std::vector<double> a, bool llt, std::vector<double>& out) {
    for (int k = 0; k < 64; k++) {
        double p = a[k * 64 + k];
        p = llt ? (p < 1e-20 ? 1e-20 : p) : clampLU(p);
        a[k * 64 + k] = p;
        for (int i = k + 1; i < 64; i++) {
            const double l = a[i * 64 + k] / p;
            a[i * 64 + k] = l;
            for (int j = k + 1; j < 64; j++) a[i * 64 + j] -= l * a[k * 64 + j];
        }
    }
    out = a;
}

static double maxabs_prod_minus_eye(const std::vector<double>& X, const std::vector<double>& Y) {   // max |X*Y - I|
    double m = 0;
    for (int i = 0; i < 64; i++) for (int j = 0; j < 64; j++) {
        double s = 0;
        for (int k = 0; k < 64; k++) s += X[i * 64 + k] * Y[k * 64 + j];
        m = std::fmax(m, std::fabs(s - (i == j ? 1.0 : 0.0)));
    }
    return m;
}

int main() {
    unsigned long long st = 12345;
    auto rnd = [&]() { st = st * 6364136223846793005ull + 1442695040888963407ull; return ((st >> 11) * (1.0 / 9007199254740992.0)) * 2 - 1; };
    double worst_lu = 0, worst_li = 0, worst_ui = 0, worst_ch = 0, worst_ci = 0, worst_noinv = 0;
    for (int trial = 0; trial < 6; trial++) {
        std::vector<double> A(4096);
        for (auto& v : A) v = rnd();
        for (int i = 0; i < 64; i++) A[i * 64 + i] += 40.0;
        if (trial == 4) { A[0] = 0.0; A[5 * 64 + 5] = 1e-12; }               // exercises the +-1e-9 clamp (pivot 0 and an inner one)
        if (trial == 5) for (int j = 0; j < 64; j++) { A[17 * 64 + j] = A[16 * 64 + j]; }   // exact zero pivot after elimination
        std::vector<double> out, WL, WUt, ip, ref;
        run<true, true, false>(A, out, WL, WUt, ip);
        ref_lu(A, false, ref);
        double e = 0, scale = 0;
        for (int q = 0; q < 4096; q++) { e = std::fmax(e, std::fabs(out[q] - ref[q])); scale = std::fmax(scale, std::fabs(ref[q])); }
        worst_lu = std::fmax(worst_lu, e / scale);
        std::vector<double> L(4096, 0), U(4096, 0), Ui(4096, 0);
        for (int i = 0; i < 64; i++) for (int j = 0; j < 64; j++) {
            L[i * 64 + j] = j < i ? out[i * 64 + j] : (i == j ? 1.0 : 0.0);
            U[i * 64 + j] = j >= i ? out[i * 64 + j] : 0.0;
            Ui[j * 64 + i] = WUt[i * 64 + j] * ip[i];       // scaled + transposed, as lu_task writes it
        }
        if (trial < 4) {      // (the clamped cases are singular: only the factors are compared)
            worst_li = std::fmax(worst_li, maxabs_prod_minus_eye(WL, L));
            worst_ui = std::fmax(worst_ui, maxabs_prod_minus_eye(U, Ui));
        }
        // without the inverses the factors must be bitwise the same
        std::vector<double> out2, d1, d2, d3;
        run<false, false, false>(A, out2, d1, d2, d3);
        for (int q = 0; q < 4096; q++) worst_noinv = std::fmax(worst_noinv, std::fabs(out2[q] - out[q]));
        // Cholesky of A*A^T + I through the LLT variant
        std::vector<double> S(4096, 0);
        for (int i = 0; i < 64; i++) for (int j = 0; j < 64; j++) { double s = 0; for (int k = 0; k < 64; k++) s += A[i * 64 + k] * A[j * 64 + k]; S[i * 64 + j] = s / 64 + (i == j); }
        if (trial == 5) for (int j = 0; j < 64; j++) { S[9 * 64 + j] = 0; S[j * 64 + 9] = 0; }     // zero pivot -> 1e-20 clamp
        run<true, false, true>(S, out, WL, WUt, ip);
        ref_lu(S, true, ref);
        e = 0; scale = 0;
        for (int q = 0; q < 4096; q++) { e = std::fmax(e, std::fabs(out[q] - ref[q])); scale = std::fmax(scale, std::fabs(ref[q])); }
        worst_ch = std::fmax(worst_ch, e / scale);
        if (trial < 5) {
            std::vector<double> C(4096, 0), Ci(4096, 0);     // chol = L1 * sqrt(D), inverse = D^-1/2 * L1^-1, as llt_task writes them
            for (int i = 0; i < 64; i++) for (int j = 0; j <= i; j++) {
                C[i * 64 + j] = (j < i ? out[i * 64 + j] : 1.0) * std::sqrt(out[j * 64 + j]);
                Ci[i * 64 + j] = WL[i * 64 + j] / std::sqrt(out[i * 64 + i]);
            }
            worst_ci = std::fmax(worst_ci, maxabs_prod_minus_eye(Ci, C));
            double r = 0;
            for (int i = 0; i < 64; i++) for (int j = 0; j < 64; j++) { double s = 0; for (int k = 0; k < 64; k++) s += C[i * 64 + k] * C[j * 64 + k]; r = std::fmax(r, std::fabs(s - S[i * 64 + j])); }
            worst_ci = std::fmax(worst_ci, r);
        }
    }
    std::printf("lu_vs_ref %.3e  Linv %.3e  Uinv %.3e  chol_vs_ref %.3e  chol_inv %.3e  noinv_diff %.3e\n", worst_lu, worst_li, worst_ui, worst_ch, worst_ci, worst_noinv);
    const bool ok = worst_lu < 1e-12 && worst_li < 1e-12 && worst_ui < 1e-12 && worst_ch < 1e-12 && worst_ci < 1e-12 && worst_noinv == 0.0;
    std::printf(ok ? "OK\n" : "FAIL\n");
    return ok ? 0 : 1;
}
