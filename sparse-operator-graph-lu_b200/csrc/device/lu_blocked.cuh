// Blocked LU of one 64x64 diagonal block with both triangular inverses, for the 256 math threads of the executor.
//
// Same mathematics as ludcmpSimple + inv_lower + inv_upper (MatrixStdDouble.cpp:2711-2784, 2787-2802, 2829-2866):
// no pivoting, unit-diagonal L, a pivot with |u_kk| < 1e-9 is replaced by +-1e-9 (lltdcmpSimple's clamp, 2640, for the
// Cholesky variant).  The schedule is built around the one thing that cannot be parallelised, the chain
// pivot -> reciprocal -> multiplier -> next pivot (64 links):
//
//   for kb = 0..3 (16-column panels; S = the block, packed L\U in place; W = packed inverses, starts as 0)
//     SWEEP   warp 0 factors the 16x16 diagonal block D_kb, one lane per row, registers only: per pivot ONE shuffle
//             (the pivot), clamp, reciprocal, multiplier, <= 15 FMAs; the pivot row goes through shared memory with
//             one __syncwarp.  It publishes row k of U, column k of L and 1/u_kk and bumps a shared-memory counter.
//     STRIPS  warps 1-4 stream BEHIND that counter, no barrier: 64 lanes own the columns of the 16-row band
//             [W_L | S] (row operations: x_i -= l_ik x_k) and 64 lanes own the rows of the 16-column band [S ; W_U]
//             (column operations: x_k *= 1/u_kk, x_j -= x_k u_kj), 16 values in registers each.  When the sweep
//             ends they are one pivot behind: U(kb, J), L(I, kb), the rows of L^-1 and the columns of U^-1 that
//             belong to this panel are final, and D_kb's own inverses LDI / UDI fall out of the identity columns.
//     barrier X (all 256)
//     TRAIL   8x8 tiles on the FP64 tensor cores (m8n8k4 DMMA):  S(I,J)   -= L(I,kb) U(kb,J)        I, J > kb
//                                                                 W_L(I,J) -= L(I,kb) W_L(kb,J)      I > kb, J <= kb
//                                                                 W_U(I,J) -= W_U(I,kb) U(kb,J)      I <= kb, J > kb
//             warp 0 takes the four tiles of D_kb+1 and goes straight on to the next SWEEP (it only ARRIVES at
//             barrier Y); warps 1-7 share the other tiles, meet at barrier Y and the strip warps catch up with
//             the sweep that is already running.
//
// W_L / W_U are the eliminations of the identity: [S | I] row operations give L^-1, [S ; I] column operations give
// U^-1 (blockwise the recurrences of inv_lower / inv_upper).  L^-1 has a unit diagonal, so both inverses share one
// 64x64 array: strictly lower part = L^-1, upper part with diagonal = U^-1.
//
// tests/emu/emu_lub.cpp runs THIS code on the host with one thread per CUDA thread (pthread barriers, emulated DMMA
// fragments / shuffles / flags) against a plain LU; tools/lu_lab.cu times it on the GPU phase by phase.
#pragma once

#if defined(SOGLU_LUB_HOST)
#define LUB_FN static inline
#define LUB_NOINLINE static
namespace soglu {
namespace lub {
namespace hw {   // provided by the host harness
void sync_warp();
void sync_math();                 // barrier X: all 256 math threads
void arrive_y();                  // barrier Y: warp 0 arrives without waiting ...
void sync_y();                    // ... warps 1-7 wait for all 256
double shfl(double v, int src_lane);
void dmma(double& c0, double& c1, double a, double b);
double rcp(double x);
double neg_rcp(double x);         // -1 / x
void pivot_init(unsigned long long* bar);          // "pivot published" barrier (device: an mbarrier, one arrival per phase)
void pivot_init_done();
void pivot_signal(unsigned long long* bar);        // release
bool pivot_ready(const unsigned long long* bar, int parity);   // acquire: has the phase with this parity completed
void prof(int slot);
}  // namespace hw
}  // namespace lub
}  // namespace soglu
#else
#include "ptx.cuh"
#define LUB_FN __device__ __forceinline__
#define LUB_NOINLINE __device__ __noinline__
namespace soglu {
namespace lub {
namespace hw {
LUB_FN void sync_warp() { __syncwarp(); }
LUB_FN void sync_math() { ptx::named_bar_sync(1, 256); }
LUB_FN void arrive_y() { asm volatile("bar.arrive 2, 256;" ::: "memory"); }
LUB_FN void sync_y() { ptx::named_bar_sync(2, 256); }
LUB_FN double shfl(double v, int src_lane) { return __shfl_sync(0xffffffffu, v, src_lane); }
LUB_FN void dmma(double& c0, double& c1, double a, double b) { ptx::dmma884(c0, c1, a, b); }
LUB_FN double rcp(double x) { return ptx::fast_rcp(x); }
LUB_FN double neg_rcp(double x) { return ptx::fast_neg_rcp(x); }
// "pivot k published": one mbarrier per pivot, initialised ONCE per kernel (lu_setup; re-initialising them per task
// from different warps lost arrivals on the B200), completed by one arrive (release.cta) per task and observed with
// try_wait on the task's phase parity (acquire.cta) -- no MEMBAR on the sweep's chain, and waiting strip warps sleep
// in the barrier unit instead of taking issue slots from the sweep warp
LUB_FN void pivot_init(unsigned long long* bar) { ptx::mbar_init(reinterpret_cast<uint64_t*>(bar), 1); }
LUB_FN void pivot_init_done() { ptx::fence_mbar_init(); }
LUB_FN void pivot_signal(unsigned long long* bar) { ptx::mbar_arrive(reinterpret_cast<uint64_t*>(bar)); }
LUB_FN bool pivot_ready(const unsigned long long* bar, int parity) {
    return ptx::mbar_try_wait(const_cast<uint64_t*>(reinterpret_cast<const uint64_t*>(bar)), (uint32_t)parity);
}
#if defined(SOGLU_LUB_DEBUG)
__device__ int g_lub_iter;
#endif
#if defined(SOGLU_LUB_PROF)
__device__ long long* g_lub_prof;
LUB_FN void prof(int slot) { if ((threadIdx.x & 31) == 0) g_lub_prof[slot * 8 + ((threadIdx.x >> 5) & 7)] = clock64(); }
#else
LUB_FN void prof(int) {}
#endif
}  // namespace hw
}  // namespace lub
}  // namespace soglu
#endif

namespace soglu {
namespace lub {

constexpr int LD = 68;        // leading dimension of S and W (tasks.h BLK_LD)
constexpr int DLD = 20;       // leading dimension of LDI / UDI: like 68, 8 banks per row -> conflict-free DMMA fragments
// scratch layout in doubles
constexpr int SCR_LDI = 0;                   // [16][DLD] inverse of the unit-lower factor of D_kb (full, zeros above)
constexpr int SCR_UDI = 16 * DLD;            // [16][DLD] inverse of its upper factor (full, zeros below)
constexpr int SCR_LC = 2 * 16 * DLD;         // [16][16]  column k of MINUS the L factor of D_kb at [k][i] (entries i > k)
constexpr int SCR_IP = SCR_LC + 256;         // [64]      -1 / u_kk
constexpr int SCR_BAR = SCR_IP + 64;         // [64] u64  "pivot k published" barriers (one phase per task)
constexpr int SCR_PAR = SCR_BAR + 64;        // int: phase parity of the barriers = diagonal tasks this CTA has run, mod 2
constexpr int SCR_TRASH = SCR_PAR + 2;       // [32][2] where the sweep's unselected lanes store, one 16-byte slot per lane (see st_sel)
constexpr int SCRATCH_DOUBLES = SCR_TRASH + 64;  // 1090

struct alignas(16) D2 { double x, y; };     // 16-byte shared-memory accesses (every offset below is even)

// Store by ONE selected lane without a divergent branch: every lane stores, the others into their own 16-byte trash
// slot (distinct banks: 31 lanes storing to ONE address are serialised, measured +60 cycles per pivot).  ptxas turns
// predicated shared-memory stores back into branches, and two divergent branches per pivot cost more than the
// pivot's whole dependency chain.
LUB_FN void st_sel(double* real, double* trash, double v, bool pred) { *(pred ? real : trash) = v; }
// -x by flipping the sign bit (an integer op on the high word instead of an FP64-pipe DADD)
LUB_FN double flip_sign(double x) {
    unsigned long long b;
    static_assert(sizeof b == sizeof x, "");
    __builtin_memcpy(&b, &x, 8);
    b ^= 0x8000000000000000ull;
    __builtin_memcpy(&x, &b, 8);
    return x;
}

template <bool LLT>
LUB_FN double clamp_pivot(double p) {
    if (LLT) return (p < 1e-20) ? 1e-20 : p;
    return (p < 1e-9 && p > -1e-9) ? ((p < 0) ? -1e-9 : 1e-9) : p;
}

// ---- SWEEP: warp 0, the 16x16 diagonal block at (d0, d0) -----------------------------------------------------------
// Lane r (0..15; lanes 16..31 shadow them so that the shuffles stay full-warp and store nothing) owns row r.
// Pivot k:  p = clamp(row k's a[k]) by shuffle, ip = 1/p on every lane; lane k stores its row (the final row k of U)
// and every lane reads it back after one __syncwarp; m = a[k] * ip is the entry of L and goes to LC[k][r].
// Pivot k is published one iteration late (after the NEXT __syncwarp, which also orders the stores of pivot k), so
// the loop has a single warp barrier per pivot.  kb is a run-time value: ONE copy of the unrolled 16-pivot code.
// ABL (lab only, tools/lu_lab.cu): ablations that break the result but show what a pivot costs --
// 1: no publication, 2: no row broadcast through shared memory, 4: no reciprocal, 8: no pivot shuffle
template <bool LLT, int ABL = 0>
LUB_FN void diag_sweep(double* S, double* scr, int kb, int lane) {
    const int d0 = 16 * kb, r = lane & 15;
    const bool store = lane < 16;
    double* D = S + d0 * LD + d0;
    double* nlc = scr + SCR_LC;
    double* nipb = scr + SCR_IP + d0;
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(scr + SCR_BAR) + d0;
    double* trash = scr + SCR_TRASH + 2 * lane;
    double a[16];
#pragma unroll
    for (int j = 0; j < 16; j += 2) {
        const D2 v = *reinterpret_cast<const D2*>(D + r * LD + j);
        a[j] = v.x; a[j + 1] = v.y;
    }
#pragma unroll
    for (int k = 0; k < 16; k++) {
        // Program order = issue order: the pivot shuffle first, then the row broadcast through shared memory (store,
        // warp barrier, loads -- all independent of the shuffle), and only then the clamp / reciprocal chain that
        // waits for the shuffle, so the broadcast's latency hides behind it.  One lane stores by address selection
        // (st_sel), not under a branch.
        const bool is_k = lane == k;
        const double pk = (ABL & 8) ? a[k] : hw::shfl(a[k], k);
        double u[16];
        if (!(ABL & 2)) {
#pragma unroll
            for (int j = (k + 1) & ~1; j < 16; j += 2) *reinterpret_cast<D2*>(is_k ? D + k * LD + j : trash) = D2{a[j], a[j + 1]};   // k even: a[k] unclamped, fixed below
            hw::sync_warp();
        }
        if (!(ABL & 1) && k > 0 && lane == 0) hw::pivot_signal(bar + k - 1);    // pivot k - 1: row of U, column of L, 1/u are visible
        if (!(ABL & 2)) {
#pragma unroll
            for (int j = (k + 1) & ~1; j < 16; j += 2) {
                const D2 v = *reinterpret_cast<const D2*>(D + k * LD + j);
                u[j] = v.x; u[j + 1] = v.y;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 16; j++) u[j] = nipb[j];
        }
        const double p = clamp_pivot<LLT>(pk);
        const double nip = (ABL & 4) ? p : hw::neg_rcp(p);        // -1 / p: the sign rides along in the Newton steps
        st_sel(D + k * LD + k, trash, p, is_k);
        st_sel(nipb + k, trash, nip, is_k);
        // No predication of the arithmetic: a lane whose row is finished (r <= k) only touches registers that are dead
        // (its U entries went to shared memory at pivot r; only a[j < r], the L entries, are read again at the end).
        const double nm = a[k] * nip;                             // -l_rk
#pragma unroll
        for (int j = k + 1; j < 16; j++) a[j] = fma(nm, u[j], a[j]);
        // The L entry goes to shared memory right away (for the strip warps and in place) and its register dies here.
        // Measured: keeping it (a[k] = nm, stored after the loop) costs 60 cycles per pivot -- ptxas then schedules the
        // reciprocal chain behind the stores.
        const bool below = store && r > k;
        st_sel(nlc + k * 16 + r, trash, nm, below);
        st_sel(D + r * LD + k, trash, flip_sign(nm), below);
    }
    hw::sync_warp();
    if (!(ABL & 1) && lane == 0) hw::pivot_signal(bar + 15);
}

LUB_FN void wait_pivot(const unsigned long long* bar, int par) {
#if defined(SOGLU_LUB_DEBUG)
    long long spins = 0;
    while (!hw::pivot_ready(bar, par)) {
        if (++spins > 2000000) { if ((threadIdx.x & 31) == 0) printf("stuck: call %d warp %d waits for pivot barrier at smem %u\n", hw::g_lub_iter, (int)threadIdx.x / 32, ptx::smem_u32(bar)); return; }
    }
#else
    while (!hw::pivot_ready(bar, par)) {}
#endif
}

// ---- STRIPS: warps 1-4 ---------------------------------------------------------------------------------------------
// q = 0..63: column q of the band [W_L | S] (rows d0..d0+15): columns >= d0+16 are S (become U(kb, J)), columns
// < d0+16 are W_L (columns d0..d0+15 start as the identity and become LDI).  x_i -= l_ik x_k, i > k.
template <bool WITH_INV>
LUB_FN void row_strip_lane(double* S, double* W, double* scr, int kb, int q, int par) {
    const int d0 = 16 * kb;
    const bool is_s = q >= d0 + 16;
    if (!is_s && !WITH_INV) return;
    const bool ident = !is_s && q >= d0;
    double* base = (is_s ? S : W) + d0 * LD + q;
    const double* nlc = scr + SCR_LC;
    const unsigned long long* bar = reinterpret_cast<const unsigned long long*>(scr + SCR_BAR) + d0;
    double x[16];
#pragma unroll
    for (int i = 0; i < 16; i++) x[i] = ident ? ((i == q - d0) ? 1.0 : 0.0) : base[i * LD];
#pragma unroll
    for (int k = 0; k < 15; k++) {
        wait_pivot(bar + k, par);
#pragma unroll
        for (int i = (k + 1) & ~1; i < 16; i += 2) {
            const D2 nl = *reinterpret_cast<const D2*>(nlc + k * 16 + i);
            if (i > k) x[i] = fma(nl.x, x[k], x[i]);
            x[i + 1] = fma(nl.y, x[k], x[i + 1]);
        }
    }
    if (ident) {
        // LDI in full (the trailing tiles multiply with it) and its strictly lower part packed into W
        double* ldi = scr + SCR_LDI;
        const int c = q - d0;
#pragma unroll
        for (int i = 0; i < 16; i++) {
            ldi[i * DLD + c] = x[i];
            if (i > c) base[i * LD] = x[i];
        }
    } else {
#pragma unroll
        for (int i = 0; i < 16; i++) base[i * LD] = x[i];
    }
}

// q = 0..63: row q of the band [S ; W_U] (columns d0..d0+15): rows >= d0+16 are S (become L(I, kb)), rows < d0+16
// are W_U (rows d0..d0+15 start as the identity and become UDI).  x_k *= 1/u_kk;  x_j -= x_k u_kj, j > k.
template <bool WITH_INV, bool WU>
LUB_FN void col_strip_lane(double* S, double* W, double* scr, int kb, int q, int par) {
    const int d0 = 16 * kb;
    const bool is_s = q >= d0 + 16;
    if (!is_s && !(WITH_INV && WU)) return;
    const bool ident = !is_s && q >= d0;
    double* base = (is_s ? S : W) + q * LD + d0;
    const double* D = S + d0 * LD + d0;
    const double* nipb = scr + SCR_IP + d0;
    const unsigned long long* bar = reinterpret_cast<const unsigned long long*>(scr + SCR_BAR) + d0;
    double x[16];
#pragma unroll
    for (int j = 0; j < 16; j += 2) {
        if (ident) { x[j] = (j == q - d0) ? 1.0 : 0.0; x[j + 1] = (j + 1 == q - d0) ? 1.0 : 0.0; }
        else { const D2 v = *reinterpret_cast<const D2*>(base + j); x[j] = v.x; x[j + 1] = v.y; }
    }
#pragma unroll
    for (int k = 0; k < 16; k++) {
        wait_pivot(bar + k, par);
        const double nm = x[k] * nipb[k];
        x[k] = flip_sign(nm);
#pragma unroll
        for (int j = (k + 1) & ~1; j < 16; j += 2) {
            const D2 u = *reinterpret_cast<const D2*>(D + k * LD + j);
            if (j > k) x[j] = fma(nm, u.x, x[j]);
            x[j + 1] = fma(nm, u.y, x[j + 1]);
        }
    }
    if (ident) {
        double* udi = scr + SCR_UDI;
        const int rr = q - d0;
#pragma unroll
        for (int j = 0; j < 16; j++) {
            udi[rr * DLD + j] = x[j];
            if (j >= rr) base[j] = x[j];
        }
    } else {
#pragma unroll
        for (int j = 0; j < 16; j += 2) *reinterpret_cast<D2*>(base + j) = D2{x[j], x[j + 1]};
    }
}

// ---- TRAIL: C(16x16) -= A(16x16) * B(16x16), one warp: 2 x 2 DMMA tiles, all 16 fragments loaded first, four independent
// accumulator chains (a dependent DMMA has 26-32 cycles of latency: tile after tile was 345 cycles per 8x8 tile) -------
LUB_FN void macro_update(double* C, const double* A, int lda, const double* B, int ldb, int lane) {
    const int g = lane >> 2, t = lane & 3;
    double af[2][4], bf[2][4];
#pragma unroll
    for (int ks = 0; ks < 4; ks++) {
        af[0][ks] = A[g * lda + 4 * ks + t];
        af[1][ks] = A[(8 + g) * lda + 4 * ks + t];
        bf[0][ks] = B[(4 * ks + t) * ldb + g];
        bf[1][ks] = B[(4 * ks + t) * ldb + 8 + g];
    }
    double p[2][2][2] = {};
#pragma unroll
    for (int ks = 0; ks < 4; ks++)
#pragma unroll
        for (int mi = 0; mi < 2; mi++)
#pragma unroll
            for (int ni = 0; ni < 2; ni++) hw::dmma(p[mi][ni][0], p[mi][ni][1], af[mi][ks], bf[ni][ks]);
#pragma unroll
    for (int mi = 0; mi < 2; mi++)
#pragma unroll
        for (int ni = 0; ni < 2; ni++) {
            D2* c = reinterpret_cast<D2*>(C + (8 * mi + g) * LD + 8 * ni + 2 * t);
            D2 v = *c;
            v.x -= p[mi][ni][0]; v.y -= p[mi][ni][1];
            *c = v;
        }
}

// macro tile `idx` of the trailing update of panel KB (16x16 blocks): [0, R*R) S, then R*Q W_L, then Q*R W_U
// (R = 16-row blocks behind the panel, Q = 16-column blocks up to and including it; compile-time per panel).
// idx 0 is the next diagonal block.
template <int KB>
LUB_FN void trailing_macro(double* S, double* W, const double* scr, int idx, int lane) {
    constexpr int d0 = 16 * KB, t0 = d0 + 16, R = 3 - KB, Q = KB + 1;
    const double* ldi = scr + SCR_LDI;
    const double* udi = scr + SCR_UDI;
    double* C;
    const double *A, *B;
    int lda = LD, ldb = LD;
    if (idx < R * R) {
        const int r0 = t0 + 16 * (idx / R), c0 = t0 + 16 * (idx % R);
        C = S + r0 * LD + c0; A = S + r0 * LD + d0; B = S + d0 * LD + c0;
    } else if (idx < R * R + R * Q) {
        const int q = idx - R * R, r0 = t0 + 16 * (q / Q), c0 = 16 * (q % Q);
        C = W + r0 * LD + c0; A = S + r0 * LD + d0;
        if (c0 < d0) B = W + d0 * LD + c0;
        else { B = ldi; ldb = DLD; }
    } else {
        const int q = idx - R * R - R * Q, r0 = 16 * (q / R), c0 = t0 + 16 * (q % R);
        C = W + r0 * LD + c0; B = S + d0 * LD + c0;
        if (r0 < d0) A = W + r0 * LD + d0;
        else { A = udi; lda = DLD; }
    }
    macro_update(C, A, lda, B, ldb, lane);
}

// warp 0 takes the next diagonal block and only ARRIVES at barrier Y (on to the next sweep); warps 1-7 share the
// other macro tiles round-robin (ONE copy of the tile code: the loop must not be unrolled -- 56 tiles inlined were
// 130 KB of code that every warp ran through once) and wait at Y.
template <int KB, bool WITH_INV, bool WU>
LUB_FN void trail(double* S, double* W, const double* scr, int warp, int lane) {
    constexpr int R = 3 - KB, Q = KB + 1;
    constexpr int total = R * R + (WITH_INV ? (WU ? 2 : 1) * R * Q : 0);      // the W_U tiles come last
    if (warp == 0) {
        trailing_macro<KB>(S, W, scr, 0, lane);
        hw::sync_warp();
        hw::arrive_y();
    } else {
#pragma unroll 1
        for (int idx = warp; idx < total; idx += 7) trailing_macro<KB>(S, W, scr, idx, lane);
        hw::sync_y();                                  // Y: the strips of the next panel may read S and W
    }
}

// once per kernel, by the 256 math threads, before the first task (ends with a barrier)
LUB_FN void lu_setup(double* scr, int ct) {
    if (ct < 64) hw::pivot_init(reinterpret_cast<unsigned long long*>(scr + SCR_BAR) + ct);
    if (ct == 0) *reinterpret_cast<int*>(scr + SCR_PAR) = 0;
    hw::pivot_init_done();
    hw::sync_math();
}

// ---- the task: S (64x64, ld 68) holds A on entry and packed L\U on exit; W (64x64, ld 68) receives the packed
// inverses (WITH_INV); scr = SCRATCH_DOUBLES doubles; ct = 0..255.  Ends with a barrier: S and W are complete.
// LLT selects lltdcmpSimple's pivot clamp (the Cholesky factor is L * sqrt(diag U), formed by the caller);
// WU = false skips U^-1 (the symmetric path only needs L^-1).
template <bool WITH_INV, bool LLT, bool WU = true>
LUB_FN void lu_blocked(double* S, double* W, double* scr, int ct) {
    const int warp = ct >> 5, lane = ct & 31;
    if (WITH_INV) {
        for (int e = ct; e < 64 * LD / 2; e += 256) reinterpret_cast<D2*>(W)[e] = D2{0.0, 0.0};
    }
    hw::prof(0);
    hw::sync_math();
    int* parp = reinterpret_cast<int*>(scr + SCR_PAR);
    const int par = *parp;
#pragma unroll 1
    for (int kb = 0; kb < 4; kb++) {
        // Math warp m is warp m + 1 of the CTA (warp 0 is the executor's producer), i.e. scheduler partition (m + 1) % 4:
        // the strips run on m = 1, 2, 5, 6 (partitions 2, 3, 2, 3) so that the sweep warp (m = 0, partition 1) does not
        // share its issue slots with a polling warp (m = 4 would).
        if (warp == 0) diag_sweep<LLT>(S, scr, kb, lane);
        else if (warp == 1 || warp == 5) row_strip_lane<WITH_INV>(S, W, scr, kb, (warp == 1 ? 0 : 32) + lane, par);
        else if (warp == 2 || warp == 6) col_strip_lane<WITH_INV, WU>(S, W, scr, kb, (warp == 2 ? 0 : 32) + lane, par);
        hw::prof(3 * kb + 1);
        hw::sync_math();                               // X: panel kb of L, U and the band of W are final
        hw::prof(3 * kb + 2);
        if (kb == 3) {
            if (ct == 0) *parp = par ^ 1;          // every pivot barrier has completed one more phase (read again after the next task's first barrier)
            break;
        }
        if (kb == 0) trail<0, WITH_INV, WU>(S, W, scr, warp, lane);
        else if (kb == 1) trail<1, WITH_INV, WU>(S, W, scr, warp, lane);
        else trail<2, WITH_INV, WU>(S, W, scr, warp, lane);
        hw::prof(3 * kb + 3);
    }
}

}  // namespace lub
}  // namespace soglu
