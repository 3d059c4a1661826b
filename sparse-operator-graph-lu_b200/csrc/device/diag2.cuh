// Register-resident 64x64 diagonal-block factorisation, TWO pivots per barrier.
//
// 256 threads as a 16x16 grid; thread (ty, tx) owns elements (ty + 16 r, tx + 16 c), r, c = 0..3, of A and
// of the two inverse accumulators W_L, W_U.  Per interval (pivots k, k+1; k even) the owners publish rows
// k, k+1 and columns k, k+1 of A (and rows k, k+1 of W_L, W_U) in the state BEFORE pivot k; after one barrier
// every thread derives row k+1 / column k+1 AFTER pivot k redundantly,
//     u1[j] = a[k+1][j] - l[k+1][k] a[k][j],        c1[i] = a[i][k+1] - l[i][k] a[k][k+1],
// and applies both rank-1 updates at once:  a[i][j] -= l[i][k] a[k][j] + l[i][k+1] u1[j].
// 32 barriers instead of 64; the per-pivot cost was dominated by the barrier and the pivot-reciprocal
// chain (tools/diag_bench.py), not by the FMAs.  Same arithmetic as one-pivot elimination up to the
// order of the two subtractions.
//
// Semantics (reference): LU without pivoting, unit L, |u_kk| < 1e-9 clamped sign-preserving (ludcmpSimple,
// MatrixStdDouble.cpp:2711-2784); LLT = true uses lltdcmpSimple's clamp (pivot < 1e-20 -> 1e-20, 2629-2668).
// WITH_INV also forms W_L = L^-1 (forward elimination of the identity with the multipliers l[i][k]) and, with
// WU, W_U = unscaled (U^T)^-1 (multipliers u[k][i] / u[k][k] from the pivot rows; scaled by 1/u_ii and
// transposed by the caller), i.e. inv_lower / inv_upper (2787-2866) in the same sweep.
//
// The per-thread steps are plain functions of (ty, tx) on caller-owned arrays so that tests/emu_diag2.cpp
// can run the very same code for all 256 "threads" on the host (each barrier interval for every thread in
// turn); SOGLU_DIAG2_HOST selects that build.
#pragma once

#ifdef SOGLU_DIAG2_HOST
#define SOGLU_HD inline
#define SOGLU_UNROLL
namespace soglu { namespace diag2 { inline double rcp(double x) { return 1.0 / x; } } }
#else
#define SOGLU_HD __device__ __forceinline__
#define SOGLU_UNROLL _Pragma("unroll")
namespace soglu { namespace diag2 { __device__ __forceinline__ double rcp(double x) { return ptx::fast_rcp(x); } } }
#endif

namespace soglu {
namespace diag2 {

constexpr int LD = 68;             // leading dimension of a block (BLK_LD)
// exchange buffers (doubles): per interval parity 8 vectors of 64, then 1/u_kk and the next pivot's reciprocal
constexpr int ROW0 = 0, ROW1 = 64, COL0 = 128, COL1 = 192, WL0 = 256, WL1 = 320, WU0 = 384, WU1 = 448, PAR = 512;
constexpr int IPBUF = 2 * PAR, IPNEXT = IPBUF + 64, SCRATCH_DOUBLES = IPNEXT + 2;

template <bool LLT>
SOGLU_HD double clamp_pivot(double p) {
    if (LLT) return (p < 1e-20) ? 1e-20 : p;
    return (p < 1e-9 && p > -1e-9) ? ((p < 0) ? -1e-9 : 1e-9) : p;
}

// owners of rows / columns k, k+1 (k = 16 kr + ko, ko even) write them to the buffers of parity `par`
template <bool WITH_INV, bool WU>
SOGLU_HD void publish(int kr, int ko, int par, double* xbuf, const double (&a)[4][4], const double (&wl)[4][4], const double (&wu)[4][4],
                      int ty, int tx) {
    double* B = xbuf + par * PAR;
    if (ty == ko || ty == ko + 1) {
        const int o = (ty == ko) ? 0 : 64;
        SOGLU_UNROLL
        for (int c = 0; c < 4; c++) {
            if (c >= kr) B[ROW0 + o + tx + 16 * c] = a[kr][c];
            if (WITH_INV && c <= kr) {
                B[WL0 + o + tx + 16 * c] = wl[kr][c];
                if (WU) B[WU0 + o + tx + 16 * c] = wu[kr][c];
            }
        }
    }
    if (tx == ko || tx == ko + 1) {
        const int o = (tx == ko) ? 0 : 64;
        SOGLU_UNROLL
        for (int r = 0; r < 4; r++)
            if (r >= kr) B[COL0 + o + ty + 16 * r] = a[r][kr];
    }
}

// load the block, start W at the identity, clamp pivot 0 and publish its reciprocal and rows/columns 0, 1
template <bool WITH_INV, bool WU, bool LLT>
SOGLU_HD void init(const double* As, double* xbuf, double (&a)[4][4], double (&wl)[4][4], double (&wu)[4][4], int ty, int tx) {
    SOGLU_UNROLL
    for (int r = 0; r < 4; r++) {
        SOGLU_UNROLL
        for (int c = 0; c < 4; c++) {
            a[r][c] = As[(ty + 16 * r) * LD + tx + 16 * c];
            if (WITH_INV) {
                wl[r][c] = (r == c && ty == tx) ? 1.0 : 0.0;
                if (WU) wu[r][c] = wl[r][c];
            }
        }
    }
    if (ty == 0 && tx == 0) {
        const double p = clamp_pivot<LLT>(a[0][0]);
        a[0][0] = p;
        xbuf[IPNEXT + 0] = rcp(p);
    }
    publish<WITH_INV, WU>(0, 0, 0, xbuf, a, wl, wu, ty, tx);
}

// pivots k = 16 kr + ko and k + 1, after the barrier that follows their publication
template <bool WITH_INV, bool WU, bool LLT>
SOGLU_HD void eliminate2(int kr, int ko, int par, double* xbuf, double (&a)[4][4], double (&wl)[4][4], double (&wu)[4][4], int ty, int tx) {
    const double* B = xbuf + par * PAR;
    const int k = 16 * kr + ko, k1 = k + 1, k2 = k + 2;
    const double ipk = xbuf[IPNEXT + par];
    const double rb0k1 = B[ROW0 + k1];                 // a[k][k+1]
    const double lk1k = B[COL0 + k1] * ipk;            // l[k+1][k]
    const double p1 = clamp_pivot<LLT>(fma(-lk1k, rb0k1, B[ROW1 + k1]));
    const double ipk1 = rcp(p1);
    if (WITH_INV && ty == 0 && tx == 0) { xbuf[IPBUF + k] = ipk; xbuf[IPBUF + k1] = ipk1; }

    // The warps holding element (k+2, k+2) finish that one element first, clamp it and publish its reciprocal for
    // the next interval, so that division overlaps everybody else's trailing update (warp-uniform branch).
    bool own_next = false;
    double p_next = 0.0;
    {
        const int ko2 = k2 & 15;
        if (k2 < 64 && (ty >> 1) == (ko2 >> 1)) {
            own_next = (ty == ko2) && (tx == ko2);
            const double a_sel = (ko != 14) ? a[kr][kr] : a[kr < 3 ? kr + 1 : 3][kr < 3 ? kr + 1 : 3];
            const double l0 = B[COL0 + k2] * ipk;
            const double l1 = fma(-l0, rb0k1, B[COL1 + k2]) * ipk1;
            const double u1 = fma(-lk1k, B[ROW0 + k2], B[ROW1 + k2]);
            p_next = clamp_pivot<LLT>(fma(-l1, u1, fma(-l0, B[ROW0 + k2], a_sel)));
            const double ipn = rcp(p_next);
            if (own_next) xbuf[IPNEXT + (par ^ 1)] = ipn;
        }
    }

    // (the three updates run one after the other so that only one set of multipliers is live at a time:
    //  a, W_L and W_U already take 96 of the 168 registers a thread can have)
    double l0[4], l1[4];
    SOGLU_UNROLL
    for (int r = 0; r < 4; r++) {
        l0[r] = l1[r] = 0.0;
        if (r >= kr) {
            const int i = ty + 16 * r;
            const bool act0 = (r > kr) || (ty > ko), act1 = (r > kr) || (ty > ko + 1);
            const double f = B[COL0 + i] * ipk;
            const double c1 = fma(-f, rb0k1, B[COL1 + i]);        // column k+1 after pivot k
            l0[r] = act0 ? f : 0.0;
            l1[r] = act1 ? c1 * ipk1 : 0.0;
        }
    }
    {
        double rb0m[4], u1m[4];
        SOGLU_UNROLL
        for (int c = 0; c < 4; c++) {
            rb0m[c] = u1m[c] = 0.0;
            if (c >= kr) {
                const int j = tx + 16 * c;
                const bool act0 = (c > kr) || (tx > ko), act1 = (c > kr) || (tx > ko + 1);
                const double rb0 = B[ROW0 + j];
                const double u1 = fma(-lk1k, rb0, B[ROW1 + j]);       // row k+1 after pivot k
                rb0m[c] = act0 ? rb0 : 0.0;
                u1m[c] = act1 ? u1 : 0.0;
            }
        }
        SOGLU_UNROLL
        for (int r = 0; r < 4; r++) {
            if (r < kr) continue;
            SOGLU_UNROLL
            for (int c = 0; c < 4; c++)
                if (c >= kr) a[r][c] = fma(-l1[r], u1m[c], fma(-l0[r], rb0m[c], a[r][c]));
            // the multipliers replace the eliminated entries of columns k and k+1
            const bool act0 = (r > kr) || (ty > ko), act1 = (r > kr) || (ty > ko + 1);
            if (tx == ko && act0) a[r][kr] = l0[r];
            if (tx == ko + 1 && act1) a[r][kr] = l1[r];
        }
    }
    if (ty == ko + 1 && tx == ko + 1) a[kr][kr] = p1;          // keep the clamped pivot k+1
    if (own_next) {                                            // and the clamped pivot k+2
        if (ko != 14) a[kr][kr] = p_next; else a[kr < 3 ? kr + 1 : 3][kr < 3 ? kr + 1 : 3] = p_next;
    }
    if (WITH_INV) {
        SOGLU_UNROLL
        for (int c = 0; c < 4; c++) {
            if (c > kr) continue;
            const int j = tx + 16 * c;
            const double w0 = B[WL0 + j];
            const double w1 = fma(-lk1k, w0, B[WL1 + j]);         // row k+1 of W_L after pivot k
            SOGLU_UNROLL
            for (int r = 0; r < 4; r++)
                if (r >= kr) wl[r][c] = fma(-l1[r], w1, fma(-l0[r], w0, wl[r][c]));
        }
    }
    if (WITH_INV && WU) {
        const double mk1k = rb0k1 * ipk;                       // u[k][k+1] / u[k][k]
        double m0[4], m1[4];                                   // multipliers of (U^T)^-1: from the pivot ROWS
        SOGLU_UNROLL
        for (int r = 0; r < 4; r++) {
            m0[r] = m1[r] = 0.0;
            if (r >= kr) {
                const int i = ty + 16 * r;
                const bool act0 = (r > kr) || (ty > ko), act1 = (r > kr) || (ty > ko + 1);
                const double g = B[ROW0 + i] * ipk;               // u[k][i] / u[k][k]
                const double u1i = fma(-lk1k, B[ROW0 + i], B[ROW1 + i]);
                m0[r] = act0 ? g : 0.0;
                m1[r] = act1 ? u1i * ipk1 : 0.0;
            }
        }
        SOGLU_UNROLL
        for (int c = 0; c < 4; c++) {
            if (c > kr) continue;
            const int j = tx + 16 * c;
            const double v0 = B[WU0 + j];
            const double v1 = fma(-mk1k, v0, B[WU1 + j]);
            SOGLU_UNROLL
            for (int r = 0; r < 4; r++)
                if (r >= kr) wu[r][c] = fma(-m1[r], v1, fma(-m0[r], v0, wu[r][c]));
        }
    }
    // rows / columns k+2, k+3 for the next interval
    if (ko != 14) publish<WITH_INV, WU>(kr, ko + 2, par ^ 1, xbuf, a, wl, wu, ty, tx);
    else if (kr < 3) publish<WITH_INV, WU>(kr + 1, 0, par ^ 1, xbuf, a, wl, wu, ty, tx);
}

}  // namespace diag2
}  // namespace soglu
