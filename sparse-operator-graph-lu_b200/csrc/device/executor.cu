// Persistent, dependency-counted DAG executor for the factorisation (replaces the OpenMP
// stage loop of BlockPlanner::calculate, BlockPlanner.cpp:376-651) and the 64x64 FP64
// block kernels it dispatches to (MatrixStdDouble.cpp, see per-function notes).
//
// One CTA per SM stays resident for a whole segment of the factorisation (one segment unless the block
// pool is smaller than the number of blocks and slots are recycled between launches):
//   warp 0      scheduler + TMA producer: claims the next task of the segment IN TASK ORDER (atomicAdd on a
//               counter; the task compiler has sorted the tasks most-urgent-first, Compiler::static_order), reads
//               its 64-byte record, spins with ld.acquire on the task group's dependency counter until it is zero
//               and streams the operand blocks into a 3-stage shared-memory ring with cp.async.bulk (34 816 B per
//               block, mbarrier complete_tx).  It runs ahead of the math warps, so the loads of task N+1 overlap
//               the MMAs and the write-back of task N -- and in the latency-bound phases a CTA already holds the
//               record of the task it is waiting for when the last predecessor finishes.
//   warps 1..8  math: m8n8k4 FP64 tensor-core MMAs (DMMA) for the Schur updates
//               C = init +/- sum_p A_p*B_p, the whole accumulation chain of one target block (or of
//               a 16- / 32-row slice of it in narrow levels) kept in registers and written once;
//               the blocked diagonal-block kernel for the rest (lu_blocked.cuh: LU / Cholesky that also form
//               L^-1 and U^-1, one warp on the pivot chain, four streaming behind it, DMMA trailing updates),
//               standalone triangular inverses, subtract.
//   warp 9      signal: the math warps hand a finished task over (mbarrier, 8 arrivals) and go straight on to
//               the next task's MMAs; this warp makes the result visible (one fence) and walks the successor
//               list: one fire-and-forget red.add(-1) on the successor group's dependency counter per lane (a task,
//               or the 2 / 4 row slices of a split GEMM task, which share their leader's counter).  Nobody waits for
//               an atomic's return value and there is no queue to append to.
//
// Why a static order cannot deadlock: the order is topological, so the lowest unfinished task of a run has all its
// predecessors finished; every lower task of its GPU is finished too, so it has been claimed (claims go in order) and
// its scheduler sees the counter reach zero.  On several GPUs every GPU's order is a filter of ONE global order and
// the argument holds for the globally lowest unfinished task.  (Cooperative launch: all CTAs are co-resident.)
//
// Memory-ordering protocol: writer CTA: st.global data -> mbarrier hand-over to the signal warp -> fence.acq_rel.gpu ->
// red.relaxed.gpu.add(dep, -1); reader CTA: ld.acquire.gpu(dep) == 0 (the reds of all predecessors form one RMW chain on
// the counter, so the acquire synchronises with every predecessor's fence) -> fence.proxy.async -> cp.async.bulk of the
// data.  (acq_rel fences, not __threadfence(): that one is MEMBAR.SC + an L1 invalidate.)  Successors on the same GPU
// use gpu scope; successors on a peer GPU (multi-GPU run: counters and block pools of the peers are mapped through
// CUDA IPC or peer access) use a system-scope fence and red over NVLink, and the readers poll with ld.acquire.sys.
//
// Watchdog: a claim-then-wait executor under a cooperative launch hangs for good if a signal is lost (a peer that died,
// a graph with a missing edge).  Every scheduler lane therefore checks, once per 1024 polls, the abort word of its
// GPU and %globaltimer against the launch's deadline; on a timeout it raises the abort word (on every GPU of the
// run) with the task position it was stuck at, all CTAs drain and soglu_factor returns SOGLU_ERR_CUDA.
#include "executor.cuh"
#include "ptx.cuh"
#include "lu_blocked.cuh"

namespace soglu {

namespace {

constexpr int N_STAGES = 3;
constexpr int N_MATH_WARPS = 8;
constexpr int N_MATH = N_MATH_WARPS * 32;          // 256
constexpr int N_THREADS = N_MATH + 64;             // + producer warp (warp 0) + signal warp (warp 9)
constexpr int SIG_RING = 4;                        // finished tasks in flight between the math warps and the signal warp
constexpr int STAGE_BYTES = 2 * BLK_BYTES;         // A and B
constexpr int BAR_MATH = 1;                        // named barrier id for the math warps

__device__ __forceinline__ unsigned long long gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ unsigned smid() {
    unsigned r;
    asm volatile("mov.u32 %0, %smid;" : "=r"(r));
    return r;
}

// block reference -> address (owner in the top 3 bits; owner 0 / world 1 on a single GPU)
__device__ __forceinline__ double* blk_ptr(const ExecParams& P, int32_t ref) {
    return P.pools[(uint32_t)ref >> REF_SHIFT] + (size_t)(ref & REF_MASK) * BLK_ELEMS;
}

struct StageDesc {
    int32_t type, flags, task, out, out2, init, out4, first_last;   // first_last: bit0 first, bit1 last
};

struct __align__(16) SmemCtl {
    uint64_t full[N_STAGES];
    uint64_t empty[N_STAGES];
    StageDesc desc[N_STAGES];
    uint64_t sig_full[SIG_RING];     // math warps -> signal warp: task sig_task[q] has issued all its stores (8 arrivals)
    uint64_t sig_empty[SIG_RING];    // signal warp -> math warps: entry q may be reused
    int32_t sig_task[SIG_RING];      // task id, -1 = leave
    double scratch[192];   // row exchange buffer + reciprocals of the standalone triangular inverses
};

// the blocked diagonal kernel keeps its scratch (pivot barriers, panel inverses) behind SmemCtl
constexpr size_t SMEM_BYTES = (size_t)N_STAGES * STAGE_BYTES + sizeof(SmemCtl) + (size_t)lub::SCRATCH_DOUBLES * sizeof(double);
static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB of shared memory a CTA can opt in to");

// ---- small block kernels on a block resident in shared memory (256 math threads) -------
__device__ __forceinline__ void math_sync() { ptx::named_bar_sync(BAR_MATH, N_MATH); }


// ---- diagonal-block tasks: lu_blocked.cuh factors the block in place in the stage's A half and leaves the packed
// inverses in its (unused) B half; the result blocks are then written out with 16-byte stores.
// (L, U) = LU(A) without pivoting, unit-diagonal L, |u_kk| < 1e-9 clamped sign-preserving (ludcmpSimple,
// MatrixStdDouble.cpp:2711-2784; plain FP64 instead of x87 long double) and, fused, both triangular inverses
// (inv_lower / inv_upper, MatrixStdDouble.cpp:2787-2802, 2829-2866).
__device__ __forceinline__ void lu_task_blocked(double* As, double* Ws, double* scr, double* gL, double* gU, double* gLi, double* gUi, bool tri, int32_t* warn, int ct) {
    const bool inv = gLi != nullptr || gUi != nullptr;
    if (inv) lub::lu_blocked<true, false>(As, Ws, scr, ct);
    else lub::lu_blocked<false, false>(As, nullptr, scr, ct);
    // tri (TF_TRI_OUT): the identically-zero half of every result is already zero in its slot -- only the pairs that touch the
    // triangle are stored (the write-out is bound by the SM's store bandwidth: 139 KB are 2.9 us, half of it 1.5 us)
    for (int e = ct; e < BLK * BLK / 2; e += N_MATH) {
        const int i = e >> 5, j = (e & 31) * 2, o = i * BLK_LD + j;
        const bool lo = !tri || j <= i, up = !tri || j + 1 >= i;
        const double2 v = *reinterpret_cast<const double2*>(As + o);
        if (lo) *reinterpret_cast<double2*>(gL + o) = make_double2(j < i ? v.x : (j == i ? 1.0 : 0.0), j + 1 < i ? v.y : (j + 1 == i ? 1.0 : 0.0));
        if (up) *reinterpret_cast<double2*>(gU + o) = make_double2(j >= i ? v.x : 0.0, j + 1 >= i ? v.y : 0.0);
        if (inv) {
            const double2 w = *reinterpret_cast<const double2*>(Ws + o);
            if (gLi && lo) *reinterpret_cast<double2*>(gLi + o) = make_double2(j < i ? w.x : (j == i ? 1.0 : 0.0), j + 1 < i ? w.y : (j + 1 == i ? 1.0 : 0.0));
            if (gUi && up) *reinterpret_cast<double2*>(gUi + o) = make_double2(j >= i ? w.x : 0.0, j + 1 >= i ? w.y : 0.0);
            // The reference's only numeric sanity signal, inv_check_diag after upperInv (MatrixStdDouble.cpp:2871-2937,
            // BlockPlanner.cpp:575-577): diag(U U^-1) within 1 +- 1e-3.  For triangular factors that diagonal is u_ii times
            // the inverse's ii entry (its off-diagonal probe is identically 0), so in effect it flags NaN / Inf pivots.
            if (gUi && warn && (j == i || j + 1 == i)) {
                const double p = (j == i) ? v.x * w.x : v.y * w.y;
                if (!(p <= 1.0 + 1e-3 && p >= 1.0 - 1e-3)) atomicAdd(warn, 1);
            }
        }
    }
    // the stage was written through the generic proxy; the next bulk copy into it comes through the async proxy
    ptx::fence_proxy_async();
}

// L = chol(A), pivot < 1e-20 clamped (lltdcmpSimple, MatrixStdDouble.cpp:2629-2668), optionally with the fused lowerInv of
// that L (inv_lower uses the stored diagonal, 2787-2802).  The blocked sweep gives A = L1 D L1^T with unit L1, so
// chol(A) = L1 sqrt(D) and chol(A)^-1 = D^-1/2 L1^-1.  Only the symmetric (LL^T) path of the planner emits this task
// (BlockPlanner.cpp:941-989).
__device__ __forceinline__ void llt_task_blocked(double* As, double* Ws, double* scr, double* gL, double* gLi, bool tri, int ct) {
    if (gLi) lub::lu_blocked<true, true, false>(As, Ws, scr, ct);
    else lub::lu_blocked<false, true, false>(As, nullptr, scr, ct);
    double* sq = scr + lub::SCR_LDI;   // [64] sqrt(d_k), [64] 1 / sqrt(d_k)   (the panel-inverse scratch is free again)
    if (ct < 64) {
        const double s = sqrt(As[ct * BLK_LD + ct]);
        sq[ct] = s;
        sq[64 + ct] = 1.0 / s;
    }
    math_sync();
    for (int e = ct; e < BLK * BLK; e += N_MATH) {
        const int i = e >> 6, j = e & 63, o = i * BLK_LD + j;
        if (tri && j > i) continue;
        gL[o] = (j < i) ? As[o] * sq[j] : (j == i ? sq[i] : 0.0);
        if (gLi) gLi[o] = ((j < i) ? Ws[o] : (j == i ? 1.0 : 0.0)) * sq[64 + i];
    }
    math_sync();                       // sq is read by everyone before the next task's kernel reuses the scratch
    ptx::fence_proxy_async();
}

// Y = T^-1 for a triangular block T in shared memory, general diagonal (standalone lowerInv /
// upperInv tasks).  Forward elimination on W = unscaled rows of the inverse; TRANS inverts an
// upper-triangular T through its transpose (multipliers read from row k, result written back
// transposed).  Multipliers depend only on T, so they are fetched one pivot ahead.
template <bool TRANS>
__device__ __forceinline__ void tri_inv_task(const double* __restrict__ Tm, double* __restrict__ xbuf, double* __restrict__ gout, int ct) {
    const int ty = ct >> 4, tx = ct & 15;
    double* wbuf = xbuf;          // [2][64]
    double* dinv = xbuf + 128;    // [64]
    double w[4][4];
    if (ct < BLK) dinv[ct] = 1.0 / Tm[ct * BLK_LD + ct];
#pragma unroll
    for (int r = 0; r < 4; r++)
#pragma unroll
        for (int c = 0; c < 4; c++) w[r][c] = (r == c && ty == tx) ? 1.0 : 0.0;
    math_sync();
    double f[4];
#pragma unroll
    for (int r = 0; r < 4; r++) f[r] = (TRANS ? Tm[ty + 16 * r] : Tm[(ty + 16 * r) * BLK_LD]) * dinv[0];
#pragma unroll
    for (int kr = 0; kr < 4; kr++) {
#pragma unroll 1
        for (int ko = 0; ko < 16; ko++) {
            const int k = 16 * kr + ko;
            const int pb = (k & 1) * 64;
            if (ty == ko) {
#pragma unroll
                for (int c = 0; c <= kr; c++) wbuf[pb + tx + 16 * c] = w[kr][c];
            }
            math_sync();
            double wv[4], fn[4];
#pragma unroll
            for (int c = 0; c <= kr; c++) wv[c] = wbuf[pb + tx + 16 * c];
            const int kn = (k + 1) & 63;
            const double dn = dinv[kn];
#pragma unroll
            for (int r = 0; r < 4; r++) fn[r] = (TRANS ? Tm[kn * BLK_LD + ty + 16 * r] : Tm[(ty + 16 * r) * BLK_LD + kn]) * dn;
            wv[kr] = (tx <= ko) ? wv[kr] : 0.0;
#pragma unroll
            for (int r = kr; r < 4; r++) {
                const double fr = ((r > kr) || (ty > ko)) ? f[r] : 0.0;
#pragma unroll
                for (int c = 0; c <= kr; c++) w[r][c] = fma(-fr, wv[c], w[r][c]);
            }
#pragma unroll
            for (int r = 0; r < 4; r++) f[r] = fn[r];
        }
    }
#pragma unroll
    for (int r = 0; r < 4; r++) {
        const double di = dinv[ty + 16 * r];
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const double v = w[r][c] * di;
            if (TRANS) gout[(tx + 16 * c) * BLK_LD + ty + 16 * r] = v;
            else gout[(ty + 16 * r) * BLK_LD + tx + 16 * c] = v;
        }
    }
}

// ---- Schur update: out = init -/+ sum_p A_p(rows) * B_p(64x64) from shared memory, FP64 tensor cores ------------
// A warp owns MT x NT DMMA tiles (8x8 each) starting at (row_base, col_base); 16 k-steps of 4 per operand stage.
// Whole block (64 rows): 8 warps x (4 x 2 tiles); half (32 rows): 8 x (2 x 2); quarter: 8 x (2 x 1).
// The warp runs ALL operand stages of the task in one software pipeline: the fragments of k-step k+1 are loaded while
// the MMAs of step k issue, and on the last step of a stage the next stage's full barrier is awaited and its first
// fragments are fetched under the last 8 MMAs -- the tensor pipe does not drain at stage boundaries (all 8 warps reach
// them together: ~450 cycles per pair were lost there).  On the task's last stage the warp's tile of the initial value
// (fused sub) is requested from global memory before the stage's MMAs, so its L2 latency runs under them; then
// out = init -/+ acc with 16-byte stores straight from the accumulators.  `it` = ring position of the task's first
// stage on entry, of its last stage on return.
template <bool TRANSB, int MT, int NT>
__device__ __forceinline__ void gemm_task(SmemCtl* ctl, double* stage_base, uint32_t& it, int32_t first_last, int32_t flags, int row_base, int col_base,
                                          int lane, double* __restrict__ out, const double* __restrict__ ini) {
    const int g = lane >> 2, t = lane & 3;
    double acc[MT][NT][2];
#pragma unroll
    for (int mi = 0; mi < MT; mi++)
#pragma unroll
        for (int ni = 0; ni < NT; ni++) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
    const int a_off = (row_base + g) * BLK_LD + t;
    const int b_off = TRANSB ? BLK_ELEMS + (col_base + g) * BLK_LD + t : BLK_ELEMS + t * BLK_LD + col_base + g;
    auto frag = [&](const double* S, int k0, double (&a)[MT], double (&b)[NT]) {
#pragma unroll
        for (int mi = 0; mi < MT; mi++) a[mi] = S[a_off + mi * 8 * BLK_LD + k0];
#pragma unroll
        for (int ni = 0; ni < NT; ni++) b[ni] = TRANSB ? S[b_off + ni * 8 * BLK_LD + k0] : S[b_off + k0 * BLK_LD + ni * 8];
    };
    int s = it % N_STAGES;
    const double* S = stage_base + (size_t)s * (STAGE_BYTES / 8);
    double a[MT], b[NT];
    frag(S, 0, a, b);
    bool last = first_last & 2;
    double2 c0[MT][NT];
#pragma unroll
    for (int mi = 0; mi < MT; mi++)
#pragma unroll
        for (int ni = 0; ni < NT; ni++) c0[mi][ni] = make_double2(0.0, 0.0);
    while (true) {
        if (last && (flags & TF_INIT)) {
#pragma unroll
            for (int mi = 0; mi < MT; mi++)
#pragma unroll
                for (int ni = 0; ni < NT; ni++) c0[mi][ni] = ptx::ld_cg_f64x2(ini + (row_base + 8 * mi + g) * BLK_LD + col_base + 8 * ni + 2 * t);
        }
        const int s_cur = s;
        const bool last_cur = last;
#pragma unroll
        for (int k0 = 0; k0 < BLK; k0 += 4) {
            double ac[MT], bc[NT];
#pragma unroll
            for (int mi = 0; mi < MT; mi++) ac[mi] = a[mi];
#pragma unroll
            for (int ni = 0; ni < NT; ni++) bc[ni] = b[ni];
            if (k0 + 4 < BLK) {
                frag(S, k0 + 4, a, b);
            } else if (!last_cur) {
                it++;
                s = it % N_STAGES;
                ptx::mbar_wait(&ctl->full[s], (it / N_STAGES) & 1);
                last = ctl->desc[s].first_last & 2;
                S = stage_base + (size_t)s * (STAGE_BYTES / 8);
                frag(S, 0, a, b);
            }
#pragma unroll
            for (int mi = 0; mi < MT; mi++)
#pragma unroll
                for (int ni = 0; ni < NT; ni++) ptx::dmma884(acc[mi][ni][0], acc[mi][ni][1], ac[mi], bc[ni]);
        }
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&ctl->empty[s_cur]);
        if (last_cur) break;
    }
    const bool neg = flags & TF_NEGATE;
#pragma unroll
    for (int mi = 0; mi < MT; mi++)
#pragma unroll
        for (int ni = 0; ni < NT; ni++) {
            const int off = (row_base + 8 * mi + g) * BLK_LD + col_base + 8 * ni + 2 * t;
            double2 v = make_double2(acc[mi][ni][0], acc[mi][ni][1]);
            if (neg) { v.x = -v.x; v.y = -v.y; }
            v.x += c0[mi][ni].x; v.y += c0[mi][ni].y;
            *reinterpret_cast<double2*>(out + off) = v;
        }
}

// math warp -> signal warp: this warp has issued all its stores of `task` (task < 0: tell the signal warp to leave)
__device__ __forceinline__ void hand_over(SmemCtl* ctl, uint32_t sig_it, int task, int mw, int lane) {
    const int q = sig_it % SIG_RING;
    if (lane == 0) {
        ptx::mbar_wait(&ctl->sig_empty[q], ((sig_it / SIG_RING) & 1) ^ 1);
        if (mw == 0) ctl->sig_task[q] = task;
    }
    __syncwarp();
    if (lane == 0) ptx::mbar_arrive(&ctl->sig_full[q]);
}

// scheduler lane, once per 1024 polls: has somebody aborted the run, or has the deadline passed?  On a timeout the
// abort word {1, task position, CTA, rank} is raised on every GPU of the run.
__device__ __noinline__ bool watchdog_expired(const ExecParams& P, int slot, unsigned long long deadline) {
    if (*reinterpret_cast<volatile int32_t*>(P.abort) != 0) return true;
    if (deadline == 0 || gtime() < deadline) return false;
    if (atomicCAS(P.abort, 0, 1) == 0) {
        P.abort[1] = slot; P.abort[2] = (int32_t)blockIdx.x; P.abort[3] = P.rank;
    }
    for (int g = 0; g < P.world; g++)
        if (g != P.rank && P.aborts[g]) atomicCAS_system(P.aborts[g], 0, 2);     // 2 = raised by a peer
    __threadfence_system();
    return true;
}

__global__ void __launch_bounds__(N_THREADS, 1) executor_kernel(const __grid_constant__ ExecParams P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* stage_base = reinterpret_cast<double*>(smem_raw);
    SmemCtl* ctl = reinterpret_cast<SmemCtl*>(smem_raw + (size_t)N_STAGES * STAGE_BYTES);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < N_STAGES; s++) {
            ptx::mbar_init(&ctl->full[s], 1);
            ptx::mbar_init(&ctl->empty[s], N_MATH_WARPS);
        }
        for (int q = 0; q < SIG_RING; q++) {
            ptx::mbar_init(&ctl->sig_full[q], N_MATH_WARPS);
            ptx::mbar_init(&ctl->sig_empty[q], 1);
        }
        ptx::fence_mbar_init();
    }
    __syncthreads();

    if (warp == 0) {
        // ================= scheduler + TMA producer =====================================
        if (lane == 0) {
            uint32_t it = 0;
            // stream the operand blocks of one ready task into the stage ring
            auto issue = [&](int t, const Task& T) {
                if (P.trace) { P.trace[6 * (size_t)t + 1] = gtime(); P.trace[6 * (size_t)t + 5] = smid(); }
                ptx::fence_proxy_async();
                const int nst = (T.type == T_GEMM) ? T.n_pairs : 1;
                const bool two = (T.type == T_GEMM || T.type == T_SUB);
                for (int p = 0; p < nst; p++, it++) {
                    const int s = it % N_STAGES;
                    ptx::mbar_wait(&ctl->empty[s], ((it / N_STAGES) & 1) ^ 1);
                    StageDesc d;
                    d.type = T.type; d.flags = T.flags; d.task = t; d.out = T.out; d.out2 = T.out2; d.init = T.init; d.out4 = T.out4;
                    d.first_last = (p == 0 ? 1 : 0) | (p == nst - 1 ? 2 : 0);
                    ctl->desc[s] = d;
                    const Pair pr = (p == 0) ? T.first[0] : ((p == 1) ? T.first[1] : P.pairs[T.pair_begin + p]);
                    double* As = stage_base + (size_t)s * (STAGE_BYTES / 8);
                    // GEMM row slice: only rows [16*row0, 16*(row0+nrows)) of A are needed (same place in smem)
                    const int a_off = (T.type == T_GEMM) ? ((T.flags >> TF_ROW0_SHIFT) & 3) * 16 * BLK_LD : 0;
                    const uint32_t a_bytes = (T.type == T_GEMM) ? (uint32_t)((T.flags >> TF_NROWS_SHIFT) & 7) * 16 * BLK_LD * 8 : (uint32_t)BLK_BYTES;
                    ptx::mbar_arrive_expect_tx(&ctl->full[s], two ? a_bytes + BLK_BYTES : a_bytes);
                    ptx::bulk_g2s(As + a_off, blk_ptr(P, pr.a) + a_off, a_bytes, &ctl->full[s]);
                    if (two) ptx::bulk_g2s(As + BLK_ELEMS, blk_ptr(P, pr.b), BLK_BYTES, &ctl->full[s]);
                }
            };
            // tell the math warps to leave once the segment's tasks are all claimed
            auto quit = [&]() {
                const int s = it % N_STAGES;
                ptx::mbar_wait(&ctl->empty[s], ((it / N_STAGES) & 1) ^ 1);
                ctl->desc[s].type = T_EXIT;
                ptx::mbar_arrive(&ctl->full[s]);
            };
            // Claim the next position and wait for its task.  Product path: position s IS the s-th task of the segment;
            // the scheduler fetches its record and spins on the group's dependency counter, which the finishing
            // predecessors count down with fire-and-forget reductions.  Debug executor (one launch per dependency
            // level, nothing to wait for): position s of the launch's task list.
            const unsigned long long deadline = P.watchdog_ns ? gtime() + P.watchdog_ns : 0;
            bool aborted = false;
            while (!aborted) {
                const int slot = atomicAdd(P.head, 1);
                if (slot >= P.n_tasks) break;
                int t;
                Task T;
                uint32_t polls = 0;
                if (P.signal) {
                    t = P.task0 + slot;
                    T = P.tasks[t];          // the record is fetched while the task is still waiting for its operands
                    const int32_t fl = T.flags;
                    const bool gemm = T.type == T_GEMM;
                    const int rows16 = (fl >> TF_NROWS_SHIFT) & 7;
                    // row slices wait on their leader's counter (slice index = first row / rows per slice)
                    const int32_t* cnt = P.dep + t - ((gemm && rows16 > 0 && rows16 < 4) ? ((fl >> TF_ROW0_SHIFT) & 3) / rows16 : 0);
                    while (true) {
                        const int32_t d = (P.world > 1) ? ptx::ld_acquire_sys(cnt) : ptx::ld_acquire(cnt);
                        if (d <= 0) break;
                        if ((++polls & 1023u) == 0 && watchdog_expired(P, slot, deadline)) { aborted = true; break; }
                    }
                    if (P.trace) P.trace[6 * (size_t)t + 0] = gtime();
                } else {
                    t = P.ready[slot];
                    T = P.tasks[t];
                }
                if (!aborted) issue(t, T);
            }
            quit();
        }
        return;
    }

    if (warp == N_MATH_WARPS + 1) {
        // ================= signal warp: release the successors of finished tasks ====================
        for (uint32_t it = 0;; it++) {
            const int q = it % SIG_RING;
            ptx::mbar_wait(&ctl->sig_full[q], (it / SIG_RING) & 1);
            const int t = ctl->sig_task[q];
            if (t < 0) break;
            if (P.signal && t != P.debug_drop) {
                const Task* T = P.tasks + t;
                const int sb = T->succ_begin, se = T->succ_end;
                // The math threads' stores happen before this fence (their mbarrier arrivals were observed above), so it
                // releases them at gpu scope.  Successors on this GPU are released at gpu scope (cheap); only successors on
                // peer GPUs pay for system-scope fences and atomics over NVLink.  A counter may be decremented from both
                // scopes: the atomics themselves are performed at the owning GPU's L2 either way.
                ptx::fence_acq_rel_gpu();
                bool remote = false;
                for (int e = sb + lane; e < se; e += 32) {
                    const int32_t ref = P.succ[e];
                    const int o = (uint32_t)ref >> REF_SHIFT, nx = ref & TASK_LOCAL_MASK;
                    if (o != P.rank) { remote = true; continue; }
                    ptx::red_add(P.dep + nx, -1);
                }
                if (__any_sync(0xffffffffu, remote)) {
                    ptx::fence_acq_rel_sys();
                    for (int e = sb + lane; e < se; e += 32) {
                        const int32_t ref = P.succ[e];
                        const int o = (uint32_t)ref >> REF_SHIFT, nx = ref & TASK_LOCAL_MASK;
                        if (o != P.rank) ptx::red_add_sys(P.deps[o] + nx, -1);
                    }
                }
                if (P.trace && lane == 0) P.trace[6 * (size_t)t + 4] = gtime();
            }
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&ctl->sig_empty[q]);
        }
        return;
    }

    // ======================= math warps ====================================================
    const int mw = warp - 1;                 // 0..7
    const int ct = threadIdx.x - 32;         // 0..255
    double* lub_scr = reinterpret_cast<double*>(ctl + 1);
    lub::lu_setup(lub_scr, ct);              // pivot barriers of the diagonal-block kernel, once per launch
    uint32_t sig_it = 0;
    for (uint32_t it = 0;; it++) {
        const int s = it % N_STAGES;
        ptx::mbar_wait(&ctl->full[s], (it / N_STAGES) & 1);
        const StageDesc d = ctl->desc[s];
        if (d.type == T_EXIT) { hand_over(ctl, sig_it++, -1, mw, lane); break; }
        double* As = stage_base + (size_t)s * (STAGE_BYTES / 8);
        double* Bs = As + BLK_ELEMS;

        if (P.trace && ct == 0) P.trace[6 * (size_t)d.task + 2] = gtime();
        if (d.type == T_GEMM) {
            // all operand stages of the task in one go (`it` moves on to the task's last stage)
            const int nrows16 = (d.flags >> TF_NROWS_SHIFT) & 7, row0 = ((d.flags >> TF_ROW0_SHIFT) & 3) * 16;
            double* out = blk_ptr(P, d.out);
            const double* ini = blk_ptr(P, d.init);
            if (d.flags & TF_TRANSB) {
                if (nrows16 == 4) gemm_task<true, 4, 2>(ctl, stage_base, it, d.first_last, d.flags, 32 * (mw >> 2), 16 * (mw & 3), lane, out, ini);
                else if (nrows16 == 2) gemm_task<true, 2, 2>(ctl, stage_base, it, d.first_last, d.flags, row0 + 16 * (mw >> 2), 16 * (mw & 3), lane, out, ini);
                else gemm_task<true, 2, 1>(ctl, stage_base, it, d.first_last, d.flags, row0, 8 * mw, lane, out, ini);
            } else {
                if (nrows16 == 4) gemm_task<false, 4, 2>(ctl, stage_base, it, d.first_last, d.flags, 32 * (mw >> 2), 16 * (mw & 3), lane, out, ini);
                else if (nrows16 == 2) gemm_task<false, 2, 2>(ctl, stage_base, it, d.first_last, d.flags, row0 + 16 * (mw >> 2), 16 * (mw & 3), lane, out, ini);
                else gemm_task<false, 2, 1>(ctl, stage_base, it, d.first_last, d.flags, row0, 8 * mw, lane, out, ini);
            }
            // no CTA barrier: each warp hands its part over and goes on to the next task
            if (P.trace && ct == 0) P.trace[6 * (size_t)d.task + 3] = gtime();
            hand_over(ctl, sig_it++, d.task, mw, lane);
            continue;
        } else {
            double* out = blk_ptr(P, d.out);
            switch (d.type) {
                case T_SUB: {
                    // R = S2 - S1 (mat_sub / mat_copy / mat_neg, MatrixStdDouble.cpp:2948-3121);
                    // a missing source was loaded from the all-zero slot
                    const double2* a2 = reinterpret_cast<const double2*>(As);
                    const double2* b2 = reinterpret_cast<const double2*>(Bs);
                    double2* o2 = reinterpret_cast<double2*>(out);
                    for (int i = ct; i < BLK_ELEMS / 2; i += N_MATH) {
                        const double2 x = a2[i], y = b2[i];
                        o2[i] = make_double2(x.x - y.x, x.y - y.y);
                    }
                    break;
                }
                case T_LU:
                    lu_task_blocked(As, Bs, lub_scr, out, blk_ptr(P, d.out2), (d.flags & TF_LINV) ? blk_ptr(P, d.init) : nullptr,
                                    (d.flags & TF_UINV) ? blk_ptr(P, d.out4) : nullptr, (d.flags & TF_TRI_OUT) != 0, P.abort + 8, ct);
                    break;
                case T_LLT:
                    llt_task_blocked(As, Bs, lub_scr, out, (d.flags & TF_LINV) ? blk_ptr(P, d.init) : nullptr, (d.flags & TF_TRI_OUT) != 0, ct);
                    break;
                case T_LOWERINV:
                    tri_inv_task<false>(As, ctl->scratch, out, ct);
                    break;
                case T_UPPERINV:
                    tri_inv_task<true>(As, ctl->scratch, out, ct);
                    break;
                default: break;
            }
        }
        // ---- diagonal / copy task complete: the stage (work space of these kernels) is free again ---------
        math_sync();
        if (lane == 0) ptx::mbar_arrive(&ctl->empty[s]);
        if (P.trace && ct == 0) P.trace[6 * (size_t)d.task + 3] = gtime();
        hand_over(ctl, sig_it++, d.task, mw, lane);
    }
}

// debug micro-benchmark: cycle counts of the diagonal-block kernels in isolation (1 CTA); pool slot 1 = the block,
// slots 2..5 receive L, U, L^-1, U^-1, slots 6 / 7 the standalone inverses of L and U
__global__ void __launch_bounds__(N_THREADS, 1) diag_bench_kernel(double* pool, int iters, long long* cycles) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* As = reinterpret_cast<double*>(smem_raw);
    double* Ws = As + BLK_ELEMS;
    SmemCtl* ctl = reinterpret_cast<SmemCtl*>(smem_raw + (size_t)N_STAGES * STAGE_BYTES);
    double* scr = reinterpret_cast<double*>(ctl + 1);
    if (threadIdx.x < 32 || threadIdx.x >= 32 + N_MATH) return;
    const int ct = threadIdx.x - 32;
    lub::lu_setup(scr, ct);
    auto slot = [&](int sl) { return pool + (size_t)sl * BLK_ELEMS; };
    long long t_lu3 = 0, t_lu = 0, t_invl = 0, t_invu = 0;
    for (int it = 0; it < iters; it++) {
        for (int i = ct; i < BLK_ELEMS; i += N_MATH) As[i] = slot(1)[i];
        math_sync();
        long long c0 = clock64();
        lu_task_blocked(As, Ws, scr, slot(2), slot(3), slot(4), slot(5), true, nullptr, ct);
        math_sync();
        long long c1 = clock64();
        for (int i = ct; i < BLK_ELEMS; i += N_MATH) As[i] = slot(1)[i];
        math_sync();
        long long c2 = clock64();
        lu_task_blocked(As, Ws, scr, slot(2), slot(3), nullptr, nullptr, true, nullptr, ct);
        math_sync();
        long long c3 = clock64();
        for (int i = ct; i < BLK_ELEMS; i += N_MATH) As[i] = slot(2)[i];
        math_sync();
        long long c4 = clock64();
        tri_inv_task<false>(As, ctl->scratch, slot(6), ct);
        math_sync();
        long long c5 = clock64();
        for (int i = ct; i < BLK_ELEMS; i += N_MATH) As[i] = slot(3)[i];
        math_sync();
        long long c6 = clock64();
        tri_inv_task<true>(As, ctl->scratch, slot(7), ct);
        math_sync();
        long long c7 = clock64();
        t_lu3 += c1 - c0; t_lu += c3 - c2; t_invl += c5 - c4; t_invu += c7 - c6;
    }
    if (ct == 0) { cycles[0] = t_lu3 / iters; cycles[1] = t_lu / iters; cycles[2] = t_invl / iters; cycles[3] = t_invu / iters; }
}

__global__ void pack_blocks_kernel(double* __restrict__ pool, const double* __restrict__ dense, const int32_t* __restrict__ slots, int64_t n) {
    for (int64_t b = blockIdx.x; b < n; b += gridDim.x) {
        if (slots[b] < 0) continue;   // block owned by another GPU
        double* dst = pool + (size_t)slots[b] * BLK_ELEMS;
        const double* src = dense + (size_t)b * BLK * BLK;
        for (int i = threadIdx.x; i < BLK_ELEMS; i += blockDim.x) {
            const int r = i / BLK_LD, c = i - r * BLK_LD;
            dst[i] = (c < BLK) ? src[r * BLK + c] : 0.0;
        }
    }
}
// sparse input upload: clear the input blocks this GPU owns, then scatter the entry list into them
__global__ void zero_blocks_kernel(double* __restrict__ pool, const int32_t* __restrict__ slots, int64_t n) {
    for (int64_t b = blockIdx.x; b < n; b += gridDim.x) {
        if (slots[b] < 0) continue;
        double2* dst = reinterpret_cast<double2*>(pool + (size_t)slots[b] * BLK_ELEMS);
        for (int i = threadIdx.x; i < BLK_ELEMS / 2; i += blockDim.x) dst[i] = make_double2(0.0, 0.0);
    }
}
__global__ void scatter_entries_kernel(double* __restrict__ pool, const int32_t* __restrict__ slots, const int32_t* __restrict__ entry_input,
                                       const int32_t* __restrict__ entry_pos, const double* __restrict__ vals, int64_t n) {
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
        const int32_t sl = slots[entry_input[k]];
        if (sl < 0) continue;         // block owned by another GPU
        const int pos = entry_pos[k];
        pool[(size_t)sl * BLK_ELEMS + (pos >> 6) * BLK_LD + (pos & 63)] = vals[k];
    }
}
__global__ void unpack_block_kernel(const double* __restrict__ pool, int32_t slot, double* __restrict__ dense) {
    const double* src = pool + (size_t)slot * BLK_ELEMS;
    for (int i = threadIdx.x; i < BLK * BLK; i += blockDim.x) dense[i] = src[(i >> 6) * BLK_LD + (i & 63)];
}

}  // namespace

size_t executor_smem_bytes() { return SMEM_BYTES; }

int executor_max_grid(int device) {
    cudaFuncSetAttribute(executor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
    int per_sm = 0, sms = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, executor_kernel, N_THREADS, SMEM_BYTES) != cudaSuccess) return 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    return per_sm * sms;
}

cudaError_t launch_executor(const ExecParams& p, int grid, cudaStream_t stream) {
    const void* kernel = (const void*)executor_kernel;
    const size_t smem = SMEM_BYTES;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    ExecParams pp = p;
    void* args[] = {&pp};
    // cooperative launch: the runtime guarantees that all CTAs are co-resident, which the
    // claim-then-wait executor relies on
    return cudaLaunchCooperativeKernel(kernel, dim3(grid), dim3(N_THREADS), args, smem, stream);
}

cudaError_t launch_diag_bench(double* pool, int iters, long long* cycles, cudaStream_t stream) {
    cudaError_t e = cudaFuncSetAttribute(diag_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
    if (e != cudaSuccess) return e;
    diag_bench_kernel<<<1, N_THREADS, SMEM_BYTES, stream>>>(pool, iters, cycles);
    return cudaGetLastError();
}

cudaError_t launch_pack_blocks(double* pool, const double* dense, const int32_t* slots, int64_t n, cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    int grid = (int)(n < 1184 ? n : 1184);
    pack_blocks_kernel<<<grid, 256, 0, stream>>>(pool, dense, slots, n);
    return cudaGetLastError();
}
cudaError_t launch_scatter_entries(double* pool, const int32_t* slots, int64_t n_blocks, const int32_t* entry_input, const int32_t* entry_pos,
                                   const double* vals, int64_t n_entries, cudaStream_t stream) {
    if (n_blocks <= 0) return cudaSuccess;
    zero_blocks_kernel<<<(int)(n_blocks < 1184 ? n_blocks : 1184), 256, 0, stream>>>(pool, slots, n_blocks);
    if (n_entries > 0) {
        const int64_t want = (n_entries + 255) / 256;
        scatter_entries_kernel<<<(int)(want < 4736 ? want : 4736), 256, 0, stream>>>(pool, slots, entry_input, entry_pos, vals, n_entries);
    }
    return cudaGetLastError();
}
cudaError_t launch_unpack_block(const double* pool, int32_t slot, double* dense, cudaStream_t stream) {
    unpack_block_kernel<<<1, 256, 0, stream>>>(pool, slot, dense);
    return cudaGetLastError();
}

}  // namespace soglu
