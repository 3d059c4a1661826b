// Persistent, dependency-counted DAG executor for the factorisation (replaces the OpenMP
// stage loop of BlockPlanner::calculate, BlockPlanner.cpp:376-651) and the 64x64 FP64
// block kernels it dispatches to (MatrixStdDouble.cpp, see per-function notes).
//
// One CTA per SM stays resident for the whole factorisation:
//   warp 0      scheduler + TMA producer: claims the next slot of the global ready queue,
//               waits until a finished predecessor publishes a task there, then streams the
//               task's operand blocks into a 3-stage shared-memory ring with cp.async.bulk
//               (34 816 B per block, mbarrier complete_tx).  It runs ahead of the math warps,
//               so the loads of task N+1 overlap the MMAs and the write-back of task N.
//   warps 1..8  math: m8n8k4 FP64 tensor-core MMAs (DMMA) for the Schur updates
//               C = init +/- sum_p A_p*B_p with the whole accumulation chain of one target
//               block kept in registers and written once; shared-memory right-looking
//               LU / Cholesky / triangular inverses / subtract for the other task types.
//               After the write-back they decrement the dependency counters of the
//               successor tasks and publish the ones that reach zero.
//
// Memory-ordering protocol (gpu scope): writer CTA: st.global data -> bar.sync ->
// __threadfence -> atomicSub(dep) [-> __threadfence -> atomicAdd(tail) -> st.release(ready)];
// reader CTA: ld.acquire(ready) -> fence.proxy.async -> cp.async.bulk of the data.
#include "executor.cuh"
#include "ptx.cuh"

namespace soglu {

namespace {

constexpr int N_STAGES = 3;
constexpr int N_MATH_WARPS = 8;
constexpr int N_MATH = N_MATH_WARPS * 32;          // 256
constexpr int N_THREADS = N_MATH + 32;             // + producer warp
constexpr int STAGE_BYTES = 2 * BLK_BYTES;         // A and B
constexpr int BAR_MATH = 1;                        // named barrier id for the math warps

struct StageDesc {
    int32_t type, flags, task, out, out2, init, first, last;
};

struct __align__(16) SmemCtl {
    uint64_t full[N_STAGES];
    uint64_t empty[N_STAGES];
    StageDesc desc[N_STAGES];
};

constexpr size_t SMEM_BYTES = (size_t)N_STAGES * STAGE_BYTES + sizeof(SmemCtl);

// ---- small block kernels on a block resident in shared memory (256 math threads) -------
__device__ __forceinline__ void math_sync() { ptx::named_bar_sync(BAR_MATH, N_MATH); }

__device__ __forceinline__ void store_block(double* __restrict__ g, const double* __restrict__ s, int ct) {
    // 4352 doubles = 2176 double2, coalesced 16-byte stores
    const double2* s2 = reinterpret_cast<const double2*>(s);
    double2* g2 = reinterpret_cast<double2*>(g);
#pragma unroll 3
    for (int i = ct; i < BLK_ELEMS / 2; i += N_MATH) g2[i] = s2[i];
}

// (L, U) = LU(A) without pivoting, unit-diagonal L, |u_kk| < 1e-9 clamped sign-preserving
// (ludcmpSimple, MatrixStdDouble.cpp:2711-2784; plain FP64 instead of x87 long double).
// Right-looking elimination, one barrier per column; A is overwritten with U, L goes to Lb.
__device__ void lu_block(double* __restrict__ A, double* __restrict__ Lb, int ct) {
    const int ty = ct >> 4, tx = ct & 15;
    for (int i = ct; i < BLK_ELEMS; i += N_MATH) Lb[i] = 0.0;
    math_sync();
    if (ct < BLK) Lb[ct * BLK_LD + ct] = 1.0;
    for (int k = 0; k < BLK; k++) {
        double p = A[k * BLK_LD + k];
        if (p < 1e-9 && p > -1e-9) p = (p < 0) ? -1e-9 : 1e-9;
        const double ip = 1.0 / p;
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const int i = ty + 16 * r;
            if (i <= k) continue;
            const double lik = A[i * BLK_LD + k] * ip;
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const int j = tx + 16 * c;
                if (j > k) A[i * BLK_LD + j] -= lik * A[k * BLK_LD + j];
            }
            if (tx == (k & 15)) Lb[i * BLK_LD + k] = lik;
        }
        if (ct == 0) A[k * BLK_LD + k] = p;
        math_sync();
    }
    // strict lower part of A still holds unscaled column values: U is upper triangular
    for (int idx = ct; idx < BLK * BLK; idx += N_MATH) {
        const int i = idx >> 6, j = idx & 63;
        if (j < i) A[i * BLK_LD + j] = 0.0;
    }
    math_sync();
}

// L = chol(A) reading the lower triangle, pivot < 1e-20 clamped (lltdcmpSimple,
// MatrixStdDouble.cpp:2629-2668).  A is used as workspace, L goes to Lb.
__device__ void llt_block(double* __restrict__ A, double* __restrict__ Lb, int ct) {
    const int ty = ct >> 4, tx = ct & 15;
    for (int i = ct; i < BLK_ELEMS; i += N_MATH) Lb[i] = 0.0;
    math_sync();
    for (int k = 0; k < BLK; k++) {
        double p = A[k * BLK_LD + k];
        if (p < 1e-20) p = 1e-20;
        const double lkk = sqrt(p);
        const double il = 1.0 / lkk;
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const int i = ty + 16 * r;
            if (i <= k) continue;
            const double lik = A[i * BLK_LD + k] * il;
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const int j = tx + 16 * c;
                if (j > k && j <= i) A[i * BLK_LD + j] -= lik * (A[j * BLK_LD + k] * il);
            }
            if (tx == (k & 15)) Lb[i * BLK_LD + k] = lik;
        }
        if (ct == 0) Lb[k * BLK_LD + k] = lkk;
        math_sync();
    }
}

// Y = L^-1 for lower-triangular L with general diagonal (inv_lower, MatrixStdDouble.cpp:
// 2787-2802).  Row-oriented elimination on W = unscaled rows: W_i -= L_ik/L_kk * W_k.
__device__ void inv_lower_block(const double* __restrict__ L, double* __restrict__ Y, int ct) {
    const int ty = ct >> 4, tx = ct & 15;
    for (int i = ct; i < BLK_ELEMS; i += N_MATH) Y[i] = 0.0;
    math_sync();
    if (ct < BLK) Y[ct * BLK_LD + ct] = 1.0;
    math_sync();
    for (int k = 0; k < BLK - 1; k++) {
        const double dk = 1.0 / L[k * BLK_LD + k];
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const int i = ty + 16 * r;
            if (i <= k) continue;
            const double f = L[i * BLK_LD + k] * dk;
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const int j = tx + 16 * c;
                if (j <= k) Y[i * BLK_LD + j] -= f * Y[k * BLK_LD + j];
            }
        }
        math_sync();
    }
    for (int idx = ct; idx < BLK * BLK; idx += N_MATH) {
        const int i = idx >> 6, j = idx & 63;
        if (j <= i) Y[i * BLK_LD + j] *= 1.0 / L[i * BLK_LD + i];
    }
    math_sync();
}

// Y = U^-1 for upper-triangular U (inv_upper, MatrixStdDouble.cpp:2829-2866).
__device__ void inv_upper_block(const double* __restrict__ U, double* __restrict__ Y, int ct) {
    const int ty = ct >> 4, tx = ct & 15;
    for (int i = ct; i < BLK_ELEMS; i += N_MATH) Y[i] = 0.0;
    math_sync();
    if (ct < BLK) Y[ct * BLK_LD + ct] = 1.0;
    math_sync();
    for (int k = BLK - 1; k > 0; k--) {
        const double dk = 1.0 / U[k * BLK_LD + k];
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const int i = ty + 16 * r;
            if (i >= k) continue;
            const double f = U[i * BLK_LD + k] * dk;
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const int j = tx + 16 * c;
                if (j >= k) Y[i * BLK_LD + j] -= f * Y[k * BLK_LD + j];
            }
        }
        math_sync();
    }
    for (int idx = ct; idx < BLK * BLK; idx += N_MATH) {
        const int i = idx >> 6, j = idx & 63;
        if (j >= i) Y[i * BLK_LD + j] *= 1.0 / U[i * BLK_LD + i];
    }
    math_sync();
}

// ---- Schur update: acc(32x16 per warp) += A(64x64) * B(64x64) from shared memory -------
// warp w: rows 32*(w>>2) .. +31, cols 16*(w&3) .. +15; 4 x 2 DMMA tiles, 16 k-steps of 4.
template <bool TRANSB>
__device__ __forceinline__ void mma_block(const double* __restrict__ As, const double* __restrict__ Bs, double (&acc)[4][2][2],
                                          int warp, int lane) {
    const int g = lane >> 2, t = lane & 3;
    const double* a0 = As + (32 * (warp >> 2) + g) * BLK_LD + t;
    const double* b0 = TRANSB ? Bs + (16 * (warp & 3) + g) * BLK_LD + t : Bs + t * BLK_LD + 16 * (warp & 3) + g;
#pragma unroll
    for (int k0 = 0; k0 < BLK; k0 += 4) {
        double a[4], b[2];
#pragma unroll
        for (int mi = 0; mi < 4; mi++) a[mi] = a0[mi * 8 * BLK_LD + k0];
#pragma unroll
        for (int ni = 0; ni < 2; ni++) b[ni] = TRANSB ? b0[ni * 8 * BLK_LD + k0] : b0[k0 * BLK_LD + ni * 8];
#pragma unroll
        for (int mi = 0; mi < 4; mi++)
#pragma unroll
            for (int ni = 0; ni < 2; ni++) ptx::dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
    }
}

__global__ void __launch_bounds__(N_THREADS, 1) executor_kernel(ExecParams P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* stage_base = reinterpret_cast<double*>(smem_raw);
    SmemCtl* ctl = reinterpret_cast<SmemCtl*>(smem_raw + (size_t)N_STAGES * STAGE_BYTES);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < N_STAGES; s++) {
            ptx::mbar_init(&ctl->full[s], 1);
            ptx::mbar_init(&ctl->empty[s], N_MATH_WARPS);
        }
        ptx::fence_mbar_init();
    }
    __syncthreads();

    if (warp == 0) {
        // ================= scheduler + TMA producer =====================================
        if (lane == 0) {
            uint32_t it = 0;
            while (true) {
                const int slot = atomicAdd(P.head, 1);
                int t = -1;
                if (slot < P.n_tasks) {
                    while ((t = ptx::ld_acquire(P.ready + slot)) < 0) __nanosleep(64);
                }
                if (t < 0) {
                    const int s = it % N_STAGES;
                    ptx::mbar_wait(&ctl->empty[s], ((it / N_STAGES) & 1) ^ 1);
                    ctl->desc[s].type = T_EXIT;
                    ptx::mbar_arrive(&ctl->full[s]);
                    break;
                }
                const Task T = P.tasks[t];
                ptx::fence_proxy_async();
                const int nst = (T.type == T_GEMM) ? T.n_pairs : 1;
                const bool two = (T.type == T_GEMM || T.type == T_SUB);
                for (int p = 0; p < nst; p++, it++) {
                    const int s = it % N_STAGES;
                    ptx::mbar_wait(&ctl->empty[s], ((it / N_STAGES) & 1) ^ 1);
                    StageDesc d;
                    d.type = T.type; d.flags = T.flags; d.task = t; d.out = T.out; d.out2 = T.out2; d.init = T.init;
                    d.first = (p == 0); d.last = (p == nst - 1);
                    ctl->desc[s] = d;
                    const Pair pr = P.pairs[T.pair_begin + p];
                    double* As = stage_base + (size_t)s * (STAGE_BYTES / 8);
                    ptx::mbar_arrive_expect_tx(&ctl->full[s], two ? 2 * BLK_BYTES : BLK_BYTES);
                    ptx::bulk_g2s(As, P.pool + (size_t)pr.a * BLK_ELEMS, BLK_BYTES, &ctl->full[s]);
                    if (two) ptx::bulk_g2s(As + BLK_ELEMS, P.pool + (size_t)pr.b * BLK_ELEMS, BLK_BYTES, &ctl->full[s]);
                }
            }
        }
        return;
    }

    // ======================= math warps ====================================================
    const int mw = warp - 1;                 // 0..7
    const int ct = threadIdx.x - 32;         // 0..255
    double acc[4][2][2];
    for (uint32_t it = 0;; it++) {
        const int s = it % N_STAGES;
        ptx::mbar_wait(&ctl->full[s], (it / N_STAGES) & 1);
        const StageDesc d = ctl->desc[s];
        if (d.type == T_EXIT) break;
        double* As = stage_base + (size_t)s * (STAGE_BYTES / 8);
        double* Bs = As + BLK_ELEMS;

        if (d.type == T_GEMM) {
            if (d.first) {
#pragma unroll
                for (int mi = 0; mi < 4; mi++)
#pragma unroll
                    for (int ni = 0; ni < 2; ni++) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
            }
            if (d.flags & TF_TRANSB) mma_block<true>(As, Bs, acc, mw, lane);
            else mma_block<false>(As, Bs, acc, mw, lane);
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&ctl->empty[s]);
            if (!d.last) continue;
            // epilogue: out = init -/+ acc, 16-byte stores straight from the accumulators
            const int g = lane >> 2, t = lane & 3;
            double* out = P.pool + (size_t)d.out * BLK_ELEMS;
            const double* ini = P.pool + (size_t)d.init * BLK_ELEMS;
            const bool neg = d.flags & TF_NEGATE, has_init = d.flags & TF_INIT;
#pragma unroll
            for (int mi = 0; mi < 4; mi++)
#pragma unroll
                for (int ni = 0; ni < 2; ni++) {
                    const int off = (32 * (mw >> 2) + 8 * mi + g) * BLK_LD + 16 * (mw & 3) + 8 * ni + 2 * t;
                    double2 v = make_double2(acc[mi][ni][0], acc[mi][ni][1]);
                    if (neg) { v.x = -v.x; v.y = -v.y; }
                    if (has_init) {
                        const double2 c0 = ptx::ld_cg_f64x2(ini + off);
                        v.x += c0.x; v.y += c0.y;
                    }
                    *reinterpret_cast<double2*>(out + off) = v;
                }
        } else {
            double* out = P.pool + (size_t)d.out * BLK_ELEMS;
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
                    lu_block(As, Bs, ct);
                    store_block(out, Bs, ct);
                    store_block(P.pool + (size_t)d.out2 * BLK_ELEMS, As, ct);
                    break;
                case T_LLT:
                    llt_block(As, Bs, ct);
                    store_block(out, Bs, ct);
                    break;
                case T_LOWERINV:
                    inv_lower_block(As, Bs, ct);
                    store_block(out, Bs, ct);
                    break;
                case T_UPPERINV:
                    inv_upper_block(As, Bs, ct);
                    store_block(out, Bs, ct);
                    break;
                default: break;
            }
        }
        // ---- task complete: make the result visible, then release the successors ---------
        math_sync();
        if (d.type != T_GEMM) {
            if (lane == 0) ptx::mbar_arrive(&ctl->empty[s]);
        }
        if (P.signal) {
            const Task* T = P.tasks + d.task;
            const int sb = T->succ_begin, se = T->succ_end;
            if (sb + ct < se) {
                __threadfence();
                for (int e = sb + ct; e < se; e += N_MATH) {
                    const int nx = P.succ[e];
                    if (atomicSub(P.dep + nx, 1) == 1) {
                        __threadfence();
                        const int pos = atomicAdd(P.tail, 1);
                        ptx::st_release(P.ready + pos, nx);
                    }
                }
            }
        }
    }
}

__global__ void pack_blocks_kernel(double* __restrict__ pool, const double* __restrict__ dense, const int32_t* __restrict__ slots, int64_t n) {
    for (int64_t b = blockIdx.x; b < n; b += gridDim.x) {
        double* dst = pool + (size_t)slots[b] * BLK_ELEMS;
        const double* src = dense + (size_t)b * BLK * BLK;
        for (int i = threadIdx.x; i < BLK_ELEMS; i += blockDim.x) {
            const int r = i / BLK_LD, c = i - r * BLK_LD;
            dst[i] = (c < BLK) ? src[r * BLK + c] : 0.0;
        }
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
    cudaError_t e = cudaFuncSetAttribute(executor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
    if (e != cudaSuccess) return e;
    ExecParams pp = p;
    void* args[] = {&pp};
    // cooperative launch: the runtime guarantees that all CTAs are co-resident, which the
    // claim-then-wait ready queue relies on
    return cudaLaunchCooperativeKernel((const void*)executor_kernel, dim3(grid), dim3(N_THREADS), args, SMEM_BYTES, stream);
}

cudaError_t launch_pack_blocks(double* pool, const double* dense, const int32_t* slots, int64_t n, cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    int grid = (int)(n < 1184 ? n : 1184);
    pack_blocks_kernel<<<grid, 256, 0, stream>>>(pool, dense, slots, n);
    return cudaGetLastError();
}
cudaError_t launch_unpack_block(const double* pool, int32_t slot, double* dense, cudaStream_t stream) {
    unpack_block_kernel<<<1, 256, 0, stream>>>(pool, slot, dense);
    return cudaGetLastError();
}

}  // namespace soglu
