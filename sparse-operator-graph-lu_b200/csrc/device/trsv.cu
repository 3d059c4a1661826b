// Block-sparse forward/back substitution on the factor blocks (replaces the sequential
// quadtree recursion of BlockPlanner::solve / lowerSolver / upperSolver / updateVector,
// BlockPlanner.cpp:669-862).
//
// After GPS ordering the factors are banded, so at block granularity the solve is a chain:
// block row i needs block row i-1.  One persistent kernel walks that chain
// as a software pipeline: block row i belongs to CTA (i mod grid).  Per row the CTA
//   * starts two TMA bulk copies at once: the NEAREST off-diagonal block (the one that
//     multiplies the segment produced last, i.e. the chain-critical one) and the diagonal
//     operator -- the explicit 64x64 inverse of the diagonal block when the factorisation
//     produced one (fused lu task), else the diagonal block itself;
//   * folds in all other off-diagonal blocks in ascending distance from the critical one,
//     each a 64x64 GEMV streamed from HBM/L2 with the next block's loads already in flight;
//   * waits for the last segment, applies the staged block and the diagonal operator from
//     shared memory, and publishes its 64 values.
// Segments are SELF-VALIDATING: y and x are pre-filled with a NaN sentinel, each consumer
// thread spins on exactly the 8-byte word it needs, so the chain has no flag, no fence and a
// single L2 round trip per hop.  Forward (L) and backward (U or L^T) sweeps share one launch.
//
// Per factor block 32 768 B are read once per sweep (HBM-bound per block); the chain of
// 2 x n_block_rows dependent hops bounds the total (reported with the number in bench.py).
#include "executor.cuh"
#include "ptx.cuh"

namespace soglu {
namespace {

constexpr int TR_THREADS = 256;
constexpr unsigned long long SENTINEL = 0xFFF8DEADFFF8DEADull;   // a NaN no arithmetic produces

__device__ __forceinline__ unsigned long long ld_volatile_u64(const double* p) {
    unsigned long long v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// Watchdog (see executor.cu): a consumer that is still polling after watchdog_ns raises the abort word and goes on
// with the sentinel (a NaN); every other poll loop sees the word at its next check and does the same, the kernel
// drains and soglu_solve reports SOGLU_ERR_CUDA with the block row that never arrived.
__device__ __noinline__ bool trsv_watchdog(const TrsvParams& P, const double* p, unsigned long long t0) {
    if (*reinterpret_cast<volatile int32_t*>(P.abort) != 0) return true;
    unsigned long long now;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(now));
    if (P.watchdog_ns == 0 || now - t0 < P.watchdog_ns) return false;
    if (atomicCAS(P.abort, 0, 1) == 0) { P.abort[1] = (int32_t)((p - ((p >= P.x && p < P.x + (size_t)P.n_rows * BLK) ? P.x : P.y)) / BLK); P.abort[2] = (int32_t)blockIdx.x; }
    return true;
}
__device__ __forceinline__ double poll_value(const TrsvParams& P, const double* p) {
    unsigned long long v = ld_volatile_u64(p);
    if (v == SENTINEL) {
        unsigned long long t0;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
        uint32_t polls = 0;
        while ((v = ld_volatile_u64(p)) == SENTINEL)
            if ((++polls & 4095u) == 0 && trsv_watchdog(P, p, t0)) break;
    }
    return __longlong_as_double((long long)v);
}
__device__ __forceinline__ void store_value(double* p, double x) {
    unsigned long long v = (unsigned long long)__double_as_longlong(x);
    if (v == SENTINEL) v = 0x7FF8000000000000ull;   // never publish the sentinel itself
    asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// a finished value goes into every GPU's copy of the vector (8-byte stores are single-copy atomic over NVLink too);
// the local copy last, so that a local consumer never runs ahead of what the peers can see by more than the link latency
__device__ __forceinline__ void publish_value(const TrsvParams& P, double* const* all, size_t idx, double x) {
    for (int g = 0; g < P.world; g++)
        if (g != P.rank) store_value(all[g] + idx, x);
    store_value(all[P.rank] + idx, x);
}

__device__ __forceinline__ const double* blk_ptr(const TrsvParams& P, int32_t ref) {
    return P.pools[(uint32_t)ref >> REF_SHIFT] + (size_t)(ref & REF_MASK) * BLK_ELEMS;
}

constexpr int RING = 4;       // off-diagonal blocks in flight per CTA (TMA bulk copies; hides NVLink latency too)

struct __align__(16) TrsvSmem {
    double last[BLK_ELEMS];          // nearest off-diagonal block of the row
    double diag[BLK_ELEMS];          // diagonal block or its explicit inverse
    double ring[RING][BLK_ELEMS];    // the other off-diagonal blocks, streamed through shared memory
    double r[BLK];                   // running right-hand side segment
    double v[BLK];                   // source segment of the block being applied
    double t[BLK];                   // r after the off-diagonal part (input of the diagonal operator)
    uint64_t bar;
    uint64_t ring_bar[RING];
};

// 4 threads per row: racc[row] -= sum_c M[row][c] * v[c]  (or M^T), M in global memory
template <bool TRANS>
__device__ __forceinline__ double gemv_part_global(const double* __restrict__ M, const double* __restrict__ v, int row, int part) {
    double s = 0.0;
    if (!TRANS) {
        const double* m = M + row * BLK_LD;
#pragma unroll
        for (int c = 0; c < 8; c++) {
            const int col = (c * 4 + part) * 2;
            const double2 a = ptx::ld_cg_f64x2(m + col);
            s = fma(a.x, v[col], s);
            s = fma(a.y, v[col + 1], s);
        }
    } else {
#pragma unroll 4
        for (int k = part; k < BLK; k += 4) s = fma(ptx::ld_cg_f64(M + k * BLK_LD + row), v[k], s);
    }
    return s;
}
template <bool TRANS>
__device__ __forceinline__ double gemv_part_smem(const double* __restrict__ M, const double* __restrict__ v, int row, int part) {
    double s = 0.0;
    if (!TRANS) {
        const double* m = M + row * BLK_LD;
#pragma unroll
        for (int c = 0; c < 8; c++) {
            const int col = (c * 4 + part) * 2;
            const double2 a = *reinterpret_cast<const double2*>(m + col);
            s = fma(a.x, v[col], s);
            s = fma(a.y, v[col + 1], s);
        }
    } else {
#pragma unroll 4
        for (int k = part; k < BLK; k += 4) s = fma(M[k * BLK_LD + row], v[k], s);
    }
    return s;
}
__device__ __forceinline__ double quad_sum(double s) {
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    return s;
}

// One block row of one sweep.  UPPER: backward sweep (columns > row); TRANS: apply blocks transposed.
template <bool UPPER, bool TRANS>
__device__ void solve_row(const TrsvParams& P, int row, const int64_t* __restrict__ ptr, const int32_t* __restrict__ col,
                          const int32_t* __restrict__ slot, const int32_t* __restrict__ diag, const int32_t* __restrict__ dinv,
                          const double* __restrict__ rhs, bool rhs_is_computed, const double* __restrict__ sol, double* const* sol_all, TrsvSmem& S,
                          uint32_t& phase, uint32_t& ring_phase, int tid) {
    const int64_t b = ptr[row], e = ptr[row + 1];
    const int nb = (int)(e - b);
    // chain-critical block: forward = largest column (< row), backward = smallest column (> row)
    const int64_t kcrit = UPPER ? b : e - 1;
    const int inv_slot = dinv[row];
    if (tid == 0) {
        const uint32_t bytes = (nb > 0 ? 2u : 1u) * BLK_BYTES;
        ptx::mbar_arrive_expect_tx(&S.bar, bytes);
        ptx::bulk_g2s(S.diag, blk_ptr(P, inv_slot != 0 ? inv_slot : diag[row]), BLK_BYTES, &S.bar);
        if (nb > 0) ptx::bulk_g2s(S.last, blk_ptr(P, slot[kcrit]), BLK_BYTES, &S.bar);
    }
    if (tid < BLK) S.r[tid] = rhs_is_computed ? poll_value(P, rhs + (size_t)row * BLK + tid) : rhs[(size_t)row * BLK + tid];
    const int rrow = tid >> 2, part = tid & 3;
    // non-critical blocks, far to near, RING bulk copies in flight
    auto blk_of = [&](int q) -> int64_t { return UPPER ? (e - 1 - q) : (b + q); };
    if (tid == 0) {
        for (int q = 0; q < RING && q < nb - 1; q++) {
            ptx::mbar_arrive_expect_tx(&S.ring_bar[q], BLK_BYTES);
            ptx::bulk_g2s(S.ring[q], blk_ptr(P, slot[blk_of(q)]), BLK_BYTES, &S.ring_bar[q]);
        }
    }
    for (int q = 0; q < nb - 1; q++) {
        const int64_t k = blk_of(q);
        const int rs = q % RING;
        __syncthreads();
        if (tid < BLK) S.v[tid] = poll_value(P, sol + (size_t)col[k] * BLK + tid);
        ptx::mbar_wait(&S.ring_bar[rs], (ring_phase >> rs) & 1);
        ring_phase ^= 1u << rs;
        __syncthreads();
        const double s = quad_sum(gemv_part_smem<TRANS>(S.ring[rs], S.v, rrow, part));
        if (part == 0) S.r[rrow] -= s;
        __syncthreads();                       // everyone is done with ring[rs]
        if (tid == 0 && q + RING < nb - 1) {
            ptx::mbar_arrive_expect_tx(&S.ring_bar[rs], BLK_BYTES);
            ptx::bulk_g2s(S.ring[rs], blk_ptr(P, slot[blk_of(q + RING)]), BLK_BYTES, &S.ring_bar[rs]);
        }
    }
    __syncthreads();
    // critical block from shared memory
    ptx::mbar_wait(&S.bar, phase);
    phase ^= 1;
    if (nb > 0) {
        if (tid < BLK) S.v[tid] = poll_value(P, sol + (size_t)col[kcrit] * BLK + tid);
        __syncthreads();
        const double s = quad_sum(gemv_part_smem<TRANS>(S.last, S.v, rrow, part));
        if (part == 0) S.t[rrow] = S.r[rrow] - s;
    } else if (tid < BLK) {
        S.t[tid] = S.r[tid];
    }
    __syncthreads();
    if (inv_slot != 0) {
        // x = D^-1 t with the explicit inverse (a GEMV instead of a 64-step substitution)
        const double s = quad_sum(gemv_part_smem<TRANS>(S.diag, S.t, rrow, part));
        if (part == 0) publish_value(P, sol_all, (size_t)row * BLK + rrow, s);
    } else if (tid < 32) {
        // substitution with the stored diagonal (lowerSolver / upperSolver, BlockPlanner.cpp:757, 821)
        const int lane = tid;
        const double* D = S.diag;
        double r0 = S.t[lane], r1 = S.t[lane + 32];
        if (!UPPER) {
            for (int k = 0; k < BLK; k++) {
                double xk = (k < 32 ? r0 : r1) / D[k * BLK_LD + k];
                xk = __shfl_sync(0xffffffffu, xk, k & 31);
                if (lane == (k & 31)) { if (k < 32) r0 = xk; else r1 = xk; }
                if (lane > k) r0 -= (TRANS ? D[k * BLK_LD + lane] : D[lane * BLK_LD + k]) * xk;
                if (lane + 32 > k) r1 -= (TRANS ? D[k * BLK_LD + lane + 32] : D[(lane + 32) * BLK_LD + k]) * xk;
            }
        } else {
            for (int k = BLK - 1; k >= 0; k--) {
                double xk = (k < 32 ? r0 : r1) / D[k * BLK_LD + k];
                xk = __shfl_sync(0xffffffffu, xk, k & 31);
                if (lane == (k & 31)) { if (k < 32) r0 = xk; else r1 = xk; }
                if (lane < k) r0 -= (TRANS ? D[k * BLK_LD + lane] : D[lane * BLK_LD + k]) * xk;
                if (lane + 32 < k) r1 -= (TRANS ? D[k * BLK_LD + lane + 32] : D[(lane + 32) * BLK_LD + k]) * xk;
            }
        }
        publish_value(P, sol_all, (size_t)row * BLK + lane, r0);
        publish_value(P, sol_all, (size_t)row * BLK + lane + 32, r1);
    }
    __syncthreads();
}

__global__ void __launch_bounds__(TR_THREADS) trsv_kernel(const __grid_constant__ TrsvParams P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    TrsvSmem& S = *reinterpret_cast<TrsvSmem*>(smem_raw);
    const int tid = threadIdx.x;
    const int G = gridDim.x;
    if (tid == 0) {
        ptx::mbar_init(&S.bar, 1);
        for (int q = 0; q < RING; q++) ptx::mbar_init(&S.ring_bar[q], 1);
        ptx::fence_mbar_init();
    }
    __syncthreads();
    uint32_t phase = 0, ring_phase = 0;
    // forward sweep: L y = b, over the block rows this GPU owns
    for (int k = blockIdx.x; k < P.n_my_rows; k += G)
        solve_row<false, false>(P, P.my_rows[k], P.l_ptr, P.l_col, P.l_slot, P.l_diag, P.l_dinv, P.b, P.b_polled != 0, P.y, P.y_all, S, phase, ring_phase, tid);
    // backward sweep: U x = y (or L^T x = y); its right-hand side is the forward result
    for (int k = blockIdx.x; k < P.n_my_rows; k += G) {
        const int row = P.my_rows[P.n_my_rows - 1 - k];
        if (P.symmetric)
            solve_row<true, true>(P, row, P.u_ptr, P.u_col, P.u_slot, P.u_diag, P.u_dinv, P.y, true, P.x, P.x_all, S, phase, ring_phase, tid);
        else
            solve_row<true, false>(P, row, P.u_ptr, P.u_col, P.u_slot, P.u_diag, P.u_dinv, P.y, true, P.x, P.x_all, S, phase, ring_phase, tid);
    }
}

__global__ void fill_sentinel_kernel(double* __restrict__ a, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) reinterpret_cast<unsigned long long*>(a)[i] = SENTINEL;
}

// Barrier across the GPUs of a sharded solve, on the device: every rank stores `epoch` into its slot of every peer's
// flag array and waits until all slots of its own array carry it.  It runs between the sentinel fill and the solve
// kernel: no GPU may publish a segment into a peer's vectors before that peer has filled them (the fill would wipe it
// and the peer would wait for ever -- which is how the watchdog found this).
struct FlagsAll { int32_t* p[MAX_GPUS]; };
__global__ void peer_epoch_kernel(FlagsAll f, int world, int rank, int32_t epoch, int32_t* abort, unsigned long long watchdog_ns) {
    const int g = threadIdx.x;
    if (g >= world) return;
    __threadfence_system();
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(f.p[g] + rank), "r"(epoch) : "memory");
    unsigned long long t0, now;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
    int32_t v;
    uint32_t polls = 0;
    while (true) {
        asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(f.p[rank] + g) : "memory");
        if (v >= epoch) break;
        if ((++polls & 1023u) == 0) {
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(now));
            if (*reinterpret_cast<volatile int32_t*>(abort) != 0) break;
            if (watchdog_ns && now - t0 > watchdog_ns) { if (atomicCAS(abort, 0, 1) == 0) { abort[1] = -1 - g; abort[2] = 0; } break; }
        }
    }
}

// ---- iterative refinement helpers: r = b - A x on the permuted, padded system (CSR), x += d ------
struct RAll { double* p[MAX_GPUS]; };
__global__ void residual_kernel(const int64_t* __restrict__ rp, const int32_t* __restrict__ ci, const double* __restrict__ v,
                                const double* __restrict__ b, const double* __restrict__ x, RAll r_all, int world, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s = b[i];
    for (int64_t k = rp[i]; k < rp[i + 1]; k++) s = fma(-v[k], x[ci[k]], s);
    for (int g = 0; g < world; g++) store_value(r_all.p[g] + i, s);     // the peers' solve polls its copy (sentinel protocol)
}
__global__ void axpy_kernel(double* __restrict__ x, const double* __restrict__ d, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] += d[i];
}

}  // namespace

cudaError_t launch_residual(const int64_t* rp, const int32_t* ci, const double* v, const double* b, const double* x, double* const* r_all, int world,
                            int64_t n, cudaStream_t stream) {
    RAll ra = {};
    for (int g = 0; g < world; g++) ra.p[g] = r_all[g];
    residual_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(rp, ci, v, b, x, ra, world, n);
    return cudaGetLastError();
}
cudaError_t launch_peer_epoch(int32_t* const* flags_all, int world, int rank, int32_t epoch, int32_t* abort, unsigned long long watchdog_ns, cudaStream_t stream) {
    FlagsAll f = {};
    for (int g = 0; g < world; g++) f.p[g] = flags_all[g];
    peer_epoch_kernel<<<1, 32, 0, stream>>>(f, world, rank, epoch, abort, watchdog_ns);
    return cudaGetLastError();
}
cudaError_t launch_fill_sentinel(double* a, int64_t n, cudaStream_t stream) {
    fill_sentinel_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(a, n);
    return cudaGetLastError();
}
cudaError_t launch_axpy(double* x, const double* d, int64_t n, cudaStream_t stream) {
    axpy_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(x, d, n);
    return cudaGetLastError();
}

int trsv_max_grid(int device) {
    cudaFuncSetAttribute(trsv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TrsvSmem));
    int per_sm = 0, sms = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, trsv_kernel, TR_THREADS, sizeof(TrsvSmem)) != cudaSuccess) return 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    return per_sm * sms;
}

cudaError_t launch_trsv(const TrsvParams& p, int grid, cudaStream_t stream) {
    cudaError_t e = cudaFuncSetAttribute(trsv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TrsvSmem));
    if (e != cudaSuccess) return e;
    TrsvParams pp = p;
    void* args[] = {&pp};
    return cudaLaunchCooperativeKernel((const void*)trsv_kernel, dim3(grid), dim3(TR_THREADS), args, sizeof(TrsvSmem), stream);
}

}  // namespace soglu
