// Blocked LU of one 64x64 diagonal block with both triangular inverses, for the 256 math threads of the executor
// (option lu_mode = 1; the register-resident one-pivot-per-barrier sweep lu3_reg stays the default until this
// kernel has been measured on a B200).
//
// Same mathematics as ludcmpSimple + inv_lower + inv_upper (MatrixStdDouble.cpp:2711-2784, 2787-2802, 2829-2866):
// no pivoting, unit-diagonal L, a pivot with |u_kk| < 1e-9 is replaced by +-1e-9.  Different schedule: the sweep
// over 64 pivots with one CTA barrier each (about 580 cycles per pivot, 37 k cycles) becomes a right-looking
// factorisation in 16-column panels whose O(n^3) part runs on the FP64 tensor cores:
//
//   for kb = 0..3                                            S = the block (packed L\U in place), W = packed inverses
//     A  warp 0 factors the 16x16 diagonal block D_kb in registers (two lanes per row, the pivot row travels through
//        shared memory, ONE __syncwarp per pivot) and forms D's L^-1 and U^-1 in the same sweep        -> LDI, UDI
//     B  panel strips, one warp each, m8n8k4 DMMA:   U(kb, J>kb) = LDI * S(kb, J)        L(I>kb, kb) = S(I, kb) * UDI
//                                                    W_L(kb, J<kb) = LDI * W_L(kb, J)    W_U(I<kb, kb) = W_U(I, kb) * UDI
//     C  trailing 8x8 tiles, DMMA:                   S(I, J)   -= L(I, kb) * U(kb, J)            I, J > kb
//                                                    W_L(I, J) -= L(I, kb) * W_L(kb, J)          I > kb, J <= kb
//                                                    W_U(I, J) -= W_U(I, kb) * U(kb, J)          I <= kb, J > kb
//        warp 0 takes the four tiles of S(kb+1, kb+1) first and goes straight on to step A of kb+1, so the serial
//        part overlaps the other warps' tiles (LDI / UDI are double-buffered).
//
// W_L / W_U are the forward eliminations of the identity: [S | I] row operations give L^-1, [S ; I] column operations
// give U^-1 (blockwise the same recurrences as inv_lower / inv_upper).  L^-1 has a unit diagonal, so both inverses
// share one 64x64 array: strictly lower part = L^-1, upper part with diagonal = U^-1.  Two CTA barriers per panel
// (8 in total) instead of 64.  Estimated 10-15 k cycles (the one-warp 16x16 sweep dominates: ~100 SASS instructions per
// pivot, structural zeros of W_L / W_U and the finished columns of D skipped per unrolled pivot);
// tests/emu/emu_lub.cpp runs THIS code on the host with
// one thread per CUDA thread (pthread barriers, emulated DMMA fragments and shuffles) against a plain LU.
#pragma once

#if defined(SOGLU_LUB_HOST)
#define LUB_FN static inline
#define LUB_NOINLINE static
namespace soglu {
namespace lub {
namespace hw {   // provided by the host harness
void sync_warp();
void sync_math();
double shfl(double v, int src_lane);
void dmma(double& c0, double& c1, double a, double b);
double rcp(double x);
}  // namespace hw
}  // namespace lub
}  // namespace soglu
#else
#include "ptx.cuh"
#define LUB_FN __device__ __forceinline__
#define LUB_NOINLINE __device__ __noinline__     // one copy of the unrolled 16-pivot sweep per variant (20 KB of code each)
namespace soglu {
namespace lub {
namespace hw {
LUB_FN void sync_warp() { __syncwarp(); }
LUB_FN void sync_math() { ptx::named_bar_sync(1, 256); }
LUB_FN double shfl(double v, int src_lane) { return __shfl_sync(0xffffffffu, v, src_lane); }
LUB_FN void dmma(double& c0, double& c1, double a, double b) { ptx::dmma884(c0, c1, a, b); }
LUB_FN double rcp(double x) { return ptx::fast_rcp(x); }
}  // namespace hw
}  // namespace lub
}  // namespace soglu
#endif

namespace soglu {
namespace lub {

constexpr int LD = 68;        // leading dimension of S and W (tasks.h BLK_LD)
constexpr int DLD = 20;       // leading dimension of LDI / UDI: like 68, 8 banks per row -> conflict-free DMMA fragments
// scratch layout in doubles
constexpr int SCR_BUF = 2 * 16 * DLD;        // one LDI + UDI pair
constexpr int SCR_LDI = 0, SCR_UDI = 16 * DLD;
constexpr int SCR_ROWA = 2 * SCR_BUF;        // [2][16] pivot row of D
constexpr int SCR_ROWL = SCR_ROWA + 32;      // [2][16] row k of the 16x16 W_L
constexpr int SCR_ROWU = SCR_ROWL + 32;      // [2][16] row k of the 16x16 W_U
constexpr int SCR_IP = SCR_ROWU + 32;        // [16] 1 / u_kk
constexpr int SCR_IPK = SCR_IP + 16;         // [2]  1 / u_kk of the pivot in use
constexpr int SCRATCH_DOUBLES = SCR_IPK + 2; // 1394

struct alignas(16) D2 { double x, y; };     // 16-byte shared-memory accesses (every offset below is even)

template <bool LLT>
LUB_FN double clamp_pivot(double p) {
    if (LLT) return (p < 1e-20) ? 1e-20 : p;
    return (p < 1e-9 && p > -1e-9) ? ((p < 0) ? -1e-9 : 1e-9) : p;
}

// ---- step A: one warp, 16x16 ---------------------------------------------------------------------------------------
// Lane (r, h) = (lane >> 1, lane & 1) owns columns 8h..8h+7 of row r of D, of W_L and of W_U (both start as I).
// Pivot k: row k goes to shared memory, every lane subtracts m_r = d_rk / d_kk times it from its part of [D | W_L]
// and m'_r = d_kr / d_kk times row k of W_U from W_U (the transposed elimination that inverts U, as in lu3_reg).
// Results: D overwritten with packed L\U, ldi = L_D^-1 (full 16x16, unit diagonal), udi = U_D^-1 (full 16x16).
template <bool LLT>
LUB_NOINLINE void diag16(double* D, double* ldi, double* udi, double* scr, int lane) {
    const int r = lane >> 1, h = lane & 1;
    double* rowA = scr + SCR_ROWA;
    double* rowL = scr + SCR_ROWL;
    double* rowU = scr + SCR_ROWU;
    double* ip16 = scr + SCR_IP;
    double* ipk = scr + SCR_IPK;
    double a[8], wl[8], wu[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {
        a[j] = D[r * LD + 8 * h + j];
        wl[j] = wu[j] = (8 * h + j == r) ? 1.0 : 0.0;
    }
    if (lane == 0) {
        const double p = clamp_pivot<LLT>(a[0]);
        a[0] = p;
        const double ip0 = hw::rcp(p);
        ipk[0] = ip0;
        ip16[0] = ip0;
    }
#pragma unroll
    for (int k = 0; k < 16; k++) {
        const int hk = k >> 3, jk = k & 7, pb = (k & 1) * 16;
        if (r == k) {
#pragma unroll
            for (int j = 0; j < 8; j += 2) {
                *reinterpret_cast<D2*>(rowA + pb + 8 * h + j) = D2{a[j], a[j + 1]};
                *reinterpret_cast<D2*>(rowL + pb + 8 * h + j) = D2{wl[j], wl[j + 1]};
                *reinterpret_cast<D2*>(rowU + pb + 8 * h + j) = D2{wu[j], wu[j + 1]};
            }
        }
        hw::sync_warp();     // row k and 1 / d_kk are visible; the buffers of pivot k - 1 may be overwritten at k + 1
        const double ip = ipk[k & 1];
        // Structural zeros, known per unrolled pivot: row k of W_L / W_U is zero beyond column k, so for k < 8 only the
        // columns j <= k can change (in either half); and for k >= 8 the columns j <= k - 8 of D are finished in both halves.
        const int jw = (k < 8) ? k + 1 : 8;        // W_L / W_U columns j < jw are updated
        const int ja = (k >= 8) ? k - 7 : 0;       // D columns j >= ja are updated
        double ra[8], rl[8], ru[8];
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
            if (j + 1 >= ja) {
                const D2 va = *reinterpret_cast<const D2*>(rowA + pb + 8 * h + j);
                ra[j] = va.x; ra[j + 1] = va.y;
            }
            if (j < jw) {
                const D2 vl = *reinterpret_cast<const D2*>(rowL + pb + 8 * h + j);
                const D2 vu = *reinterpret_cast<const D2*>(rowU + pb + 8 * h + j);
                rl[j] = vl.x; rl[j + 1] = vl.y; ru[j] = vu.x; ru[j + 1] = vu.y;
            }
        }
        const double mc = a[jk] * ip;                             // d_rk / d_kk, meaningful in the half that holds column k
        double m = hw::shfl(mc, (lane & ~1) | hk);
        const bool act = r > k;
        m = act ? m : 0.0;
        const double m2 = act ? rowA[pb + r] * ip : 0.0;          // d_kr / d_kk
#pragma unroll
        for (int j = 0; j < 8; j++) {
            if (j >= ja) {
                const double rv = (8 * h + j > k) ? ra[j] : 0.0;  // columns <= k of D are finished
                a[j] = fma(-m, rv, a[j]);
            }
            if (j < jw) {
                wl[j] = fma(-m, rl[j], wl[j]);
                wu[j] = fma(-m2, ru[j], wu[j]);
            }
        }
        if (act && h == hk) a[jk] = m;                            // the multiplier is the entry of L
        if (k < 15) {
            // next pivot: clamp + reciprocal.  Every lane runs the arithmetic on its own element (straight-line code: the
            // reciprocal's latency overlaps the updates above instead of a one-lane divergent branch); only the lane that
            // owns d(k+1, k+1) keeps and publishes the result (visible after the next sync).
            const int k1 = k + 1, hk1 = k1 >> 3, jk1 = k1 & 7;
            const double p = clamp_pivot<LLT>(a[jk1]);
            const double ipn = hw::rcp(p);
            if (r == k1 && h == hk1) {
                a[jk1] = p;
                ipk[k1 & 1] = ipn;
                ip16[k1] = ipn;
            }
        }
    }
    hw::sync_warp();         // ip16 complete
    const double ipr = ip16[r];
#pragma unroll
    for (int j = 0; j < 8; j++) {
        D[r * LD + 8 * h + j] = a[j];
        ldi[r * DLD + 8 * h + j] = wl[j];
        udi[(8 * h + j) * DLD + r] = wu[j] * ipr;                 // U^-1 = W_U^T * diag(1 / u_ii)
    }
}

// ---- step B: panel strips (one warp each) ----------------------------------------------------------------------------
// X(16x8) = LDI(16x16) * X
LUB_FN void row_strip(double* X, const double* ldi, int lane) {
    const int g = lane >> 2, t = lane & 3;
    double b[4];
#pragma unroll
    for (int ks = 0; ks < 4; ks++) b[ks] = X[(4 * ks + t) * LD + g];
    double c[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
#pragma unroll
    for (int mi = 0; mi < 2; mi++)
#pragma unroll
        for (int ks = 0; ks < 4; ks++) {
            if (mi == 0 && ks >= 2) continue;                     // LDI is lower triangular
            hw::dmma(c[mi][0], c[mi][1], ldi[(8 * mi + g) * DLD + 4 * ks + t], b[ks]);
        }
    hw::sync_warp();         // every lane has read its part of X
#pragma unroll
    for (int mi = 0; mi < 2; mi++) { X[(8 * mi + g) * LD + 2 * t] = c[mi][0]; X[(8 * mi + g) * LD + 2 * t + 1] = c[mi][1]; }
}
// X(8x16) = X * UDI(16x16)
LUB_FN void col_strip(double* X, const double* udi, int lane) {
    const int g = lane >> 2, t = lane & 3;
    double a[4];
#pragma unroll
    for (int ks = 0; ks < 4; ks++) a[ks] = X[g * LD + 4 * ks + t];
    double c[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
#pragma unroll
    for (int ni = 0; ni < 2; ni++)
#pragma unroll
        for (int ks = 0; ks < 4; ks++) {
            if (ni == 0 && ks >= 2) continue;                     // UDI is upper triangular
            hw::dmma(c[ni][0], c[ni][1], a[ks], udi[(4 * ks + t) * DLD + 8 * ni + g]);
        }
    hw::sync_warp();
#pragma unroll
    for (int ni = 0; ni < 2; ni++) { X[g * LD + 8 * ni + 2 * t] = c[ni][0]; X[g * LD + 8 * ni + 2 * t + 1] = c[ni][1]; }
}

template <bool WITH_INV, bool WU>
LUB_FN void panel_strips(double* S, double* W, const double* ldi, const double* udi, int kb, int warp, int lane) {
    const int d0 = 16 * kb;
    // items 0..5: row strips (S to the right of the diagonal block, then W_L to its left), 6..11: column strips
    // (S below, then W_U above), 12: the diagonal block of W
    for (int item = warp; item < 13; item += 8) {
        if (item < 6) {
            const int n_s = 2 * (3 - kb);
            if (item < n_s) row_strip(S + d0 * LD + d0 + 16 + 8 * item, ldi, lane);
            else if (WITH_INV) row_strip(W + d0 * LD + 8 * (item - n_s), ldi, lane);
        } else if (item < 12) {
            const int q = item - 6, n_s = 2 * (3 - kb);
            if (q < n_s) col_strip(S + (d0 + 16 + 8 * q) * LD + d0, udi, lane);
            else if (WITH_INV && WU) col_strip(W + (8 * (q - n_s)) * LD + d0, udi, lane);
        } else if (WITH_INV) {
            for (int e = lane; e < 256; e += 32) {
                const int i = e >> 4, j = e & 15;
                W[(d0 + i) * LD + d0 + j] = (j < i) ? ldi[i * DLD + j] : (WU ? udi[i * DLD + j] : 0.0);
            }
        }
    }
}

// ---- step C: C(8x8) -= A(8x16) * B(16x8) ---------------------------------------------------------------------------------
LUB_FN void tile_update(double* C, const double* A, int lda, const double* B, int ldb, int lane) {
    const int g = lane >> 2, t = lane & 3;
    double p0 = 0.0, p1 = 0.0;
#pragma unroll
    for (int ks = 0; ks < 4; ks++) hw::dmma(p0, p1, A[g * lda + 4 * ks + t], B[(4 * ks + t) * ldb + g]);
    C[g * LD + 2 * t] -= p0;
    C[g * LD + 2 * t + 1] -= p1;
}

// tile `idx` of the trailing update of panel kb: [0, R*R) S tiles, then R*Q W_L tiles, then Q*R W_U tiles
// (R = 8-row tiles behind the panel, Q = 8-column tiles up to and including it)
LUB_FN void trailing_tile(double* S, double* W, const double* ldi, const double* udi, int kb, int idx, int lane) {
    const int d0 = 16 * kb, t0 = d0 + 16, R = 2 * (3 - kb), Q = 2 * (kb + 1);
    if (idx < R * R) {
        const int r0 = t0 + 8 * (idx / R), c0 = t0 + 8 * (idx % R);
        tile_update(S + r0 * LD + c0, S + r0 * LD + d0, LD, S + d0 * LD + c0, LD, lane);
    } else if (idx < R * R + R * Q) {
        const int q = idx - R * R, r0 = t0 + 8 * (q / Q), c0 = 8 * (q % Q);
        if (c0 < d0) tile_update(W + r0 * LD + c0, S + r0 * LD + d0, LD, W + d0 * LD + c0, LD, lane);
        else tile_update(W + r0 * LD + c0, S + r0 * LD + d0, LD, ldi + (c0 - d0), DLD, lane);
    } else {
        const int q = idx - R * R - R * Q, r0 = 8 * (q / R), c0 = t0 + 8 * (q % R);
        if (r0 < d0) tile_update(W + r0 * LD + c0, W + r0 * LD + d0, LD, S + d0 * LD + c0, LD, lane);
        else tile_update(W + r0 * LD + c0, udi + (r0 - d0) * DLD, DLD, S + d0 * LD + c0, LD, lane);
    }
}

// ---- the task: S (64x64, ld 68) holds A on entry and packed L\U on exit; W (64x64, ld 68) receives the packed
// inverses (WITH_INV); scr = SCRATCH_DOUBLES doubles; ct = 0..255.  Ends with a barrier: S and W are complete.
// LLT selects lltdcmpSimple's pivot clamp (the Cholesky factor is L * sqrt(diag U), formed by the caller);
// WU = false skips U^-1 (the symmetric path only needs L^-1).
template <bool WITH_INV, bool LLT, bool WU = true>
LUB_FN void lu_blocked(double* S, double* W, double* scr, int ct) {
    const int warp = ct >> 5, lane = ct & 31;
    if (WITH_INV) {
        for (int e = ct; e < 64 * LD; e += 256) W[e] = 0.0;
    }
    if (warp == 0) diag16<LLT>(S, scr + SCR_LDI, scr + SCR_UDI, scr, lane);
    hw::sync_math();
    for (int kb = 0; kb < 4; kb++) {
        const double* ldi = scr + (kb & 1) * SCR_BUF + SCR_LDI;
        const double* udi = scr + (kb & 1) * SCR_BUF + SCR_UDI;
        panel_strips<WITH_INV, WU>(S, W, ldi, udi, kb, warp, lane);
        hw::sync_math();
        if (kb == 3) break;
        const int R = 2 * (3 - kb), Q = 2 * (kb + 1);
        const int total = R * R + (WITH_INV ? (WU ? 2 : 1) * R * Q : 0);      // the W_U tiles come last
        if (warp == 0) {
            // the next diagonal block first, then its factorisation while the other warps finish the update
            trailing_tile(S, W, ldi, udi, kb, 0, lane);
            trailing_tile(S, W, ldi, udi, kb, 1, lane);
            trailing_tile(S, W, ldi, udi, kb, R, lane);
            trailing_tile(S, W, ldi, udi, kb, R + 1, lane);
            hw::sync_warp();
            double* nb = scr + ((kb + 1) & 1) * SCR_BUF;
            diag16<LLT>(S + (16 * (kb + 1)) * (LD + 1), nb + SCR_LDI, nb + SCR_UDI, scr, lane);
        } else {
            int mine = warp - 1;            // position among the tiles left to warps 1..7
            for (int idx = 0; idx < total; idx++) {
                if (idx == 0 || idx == 1 || idx == R || idx == R + 1) continue;
                if (mine == 0) trailing_tile(S, W, ldi, udi, kb, idx, lane);
                mine = (mine == 0) ? 6 : mine - 1;
            }
        }
        hw::sync_math();
    }
}

}  // namespace lub
}  // namespace soglu
