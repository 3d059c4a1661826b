// Block-sparse forward/back substitution on the factor blocks (replaces the sequential
// quadtree recursion of BlockPlanner::solve / lowerSolver / upperSolver / updateVector,
// BlockPlanner.cpp:669-862).
//
// After GPS ordering the factors are banded, so at block granularity the solve is a chain:
// block row i needs block row i-1.  One persistent kernel walks that
// chain as a software pipeline: block row i belongs to CTA (i mod grid); the CTA consumes
// the off-diagonal blocks of its row in ascending column order, each as soon as the
// producing row has published its segment (per-row flag, release/acquire at gpu scope),
// so everything except the last dependency is already folded in when row i-1 finishes.
// The forward (L) and backward (U) sweeps run in the same launch; row i of the backward
// sweep additionally waits for y_i of the forward sweep.
//
// Work per off-diagonal block: a 64x64 GEMV straight from HBM/L2 (each factor block is read
// exactly once per sweep -> HBM-bound per block, latency-bound along the chain).
// The diagonal 64x64 triangular solve is done by one warp with the block staged in shared
// memory (divides by the stored diagonal like lowerSolver/upperSolver, 757 / 821).
#include "executor.cuh"
#include "ptx.cuh"

namespace soglu {
namespace {

constexpr int TR_THREADS = 256;

__device__ __forceinline__ void wait_flag(const int32_t* f) {
    while (ptx::ld_acquire(f) == 0) __nanosleep(32);
}

// r[0..63] -= M * v  (or M^T * v), M a pool block (ld 68), 256 threads: 4 threads per row
template <bool TRANS>
__device__ __forceinline__ void gemv_sub(const double* __restrict__ M, const double* __restrict__ v, double* __restrict__ racc, int tid) {
    // thread (row = tid>>2, part = tid&3) handles 16 interleaved double2 chunks of its row
    const int row = tid >> 2, part = tid & 3;
    double s = 0.0;
    if (!TRANS) {
        const double* m = M + row * BLK_LD;
#pragma unroll
        for (int c = 0; c < 8; c++) {
            const int col = (c * 4 + part) * 2;
            const double2 a = ptx::ld_cg_f64x2(m + col);
            s += a.x * v[col] + a.y * v[col + 1];
        }
    } else {
        // (M^T v)[row] = sum_k M[k][row] v[k]; part splits k
#pragma unroll 4
        for (int k = part; k < BLK; k += 4) s += ptx::ld_cg_f64(M + k * BLK_LD + row) * v[k];
    }
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    if (part == 0) racc[row] -= s;
}

// one sweep over one block row: rhs segment -> solution segment
template <bool UPPER, bool TRANS>
__device__ void solve_row(const double* __restrict__ pool, int row, const int64_t* __restrict__ ptr, const int32_t* __restrict__ col,
                          const int32_t* __restrict__ slot, const int32_t* __restrict__ diag, const double* __restrict__ rhs,
                          double* __restrict__ sol, int32_t* __restrict__ done, const int32_t* __restrict__ also_wait,
                          double* sD, double* sR, double* sV, int tid) {
    if (also_wait) {
        if (tid == 0) wait_flag(also_wait + row);
    }
    __syncthreads();
    if (tid < BLK) sR[tid] = ptx::ld_cg_f64(rhs + (size_t)row * BLK + tid);
    // stage the diagonal block while waiting for dependencies
    {
        const double* D = pool + (size_t)diag[row] * BLK_ELEMS;
        for (int i = tid; i < BLK_ELEMS / 2; i += TR_THREADS) reinterpret_cast<double2*>(sD)[i] = ptx::ld_cg_f64x2(D + 2 * i);
    }
    __syncthreads();
    const int64_t b = ptr[row], e = ptr[row + 1];
    // forward: ascending columns; backward: descending columns (nearest dependency last)
    for (int64_t q = 0; q < e - b; q++) {
        const int64_t k = UPPER ? (e - 1 - q) : (b + q);
        const int c = col[k];
        if (tid == 0) wait_flag(done + c);
        __syncthreads();
        if (tid < BLK) sV[tid] = ptx::ld_cg_f64(sol + (size_t)c * BLK + tid);
        __syncthreads();
        gemv_sub<TRANS>(pool + (size_t)slot[k] * BLK_ELEMS, sV, sR, tid);
        __syncthreads();
    }
    // diagonal solve by warp 0: lane owns rows lane and lane+32
    if (tid < 32) {
        const int lane = tid;
        double r0 = sR[lane], r1 = sR[lane + 32];
        if (!UPPER) {
            for (int k = 0; k < BLK; k++) {
                const double dkk = sD[k * BLK_LD + k];
                double xk = (k < 32 ? r0 : r1) / dkk;
                xk = __shfl_sync(0xffffffffu, xk, k & 31);
                if (lane == (k & 31)) { if (k < 32) r0 = xk; else r1 = xk; }
                // element (i,k) of the triangular matrix: D[i][k], or D[k][i] when transposed
                if (lane > k) r0 -= (TRANS ? sD[k * BLK_LD + lane] : sD[lane * BLK_LD + k]) * xk;
                if (lane + 32 > k) r1 -= (TRANS ? sD[k * BLK_LD + lane + 32] : sD[(lane + 32) * BLK_LD + k]) * xk;
            }
        } else {
            for (int k = BLK - 1; k >= 0; k--) {
                const double dkk = sD[k * BLK_LD + k];
                double xk = (k < 32 ? r0 : r1) / dkk;
                xk = __shfl_sync(0xffffffffu, xk, k & 31);
                if (lane == (k & 31)) { if (k < 32) r0 = xk; else r1 = xk; }
                if (lane < k) r0 -= (TRANS ? sD[k * BLK_LD + lane] : sD[lane * BLK_LD + k]) * xk;
                if (lane + 32 < k) r1 -= (TRANS ? sD[k * BLK_LD + lane + 32] : sD[(lane + 32) * BLK_LD + k]) * xk;
            }
        }
        sol[(size_t)row * BLK + lane] = r0;
        sol[(size_t)row * BLK + lane + 32] = r1;
        __syncwarp();
        if (lane == 0) {
            __threadfence();
            ptx::st_release(done + row, 1);
        }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(TR_THREADS) trsv_kernel(TrsvParams P) {
    __shared__ __align__(16) double sD[BLK_ELEMS];
    __shared__ double sR[BLK];
    __shared__ double sV[BLK];
    const int tid = threadIdx.x;
    const int G = gridDim.x;
    // forward sweep: L y = b
    for (int row = blockIdx.x; row < P.n_rows; row += G)
        solve_row<false, false>(P.pool, row, P.l_ptr, P.l_col, P.l_slot, P.l_diag, P.b, P.y, P.done_l, nullptr, sD, sR, sV, tid);
    // backward sweep: U x = y (or L^T x = y)
    for (int r = blockIdx.x; r < P.n_rows; r += G) {
        const int row = P.n_rows - 1 - r;
        if (P.symmetric)
            solve_row<true, true>(P.pool, row, P.u_ptr, P.u_col, P.u_slot, P.u_diag, P.y, P.x, P.done_u, P.done_l, sD, sR, sV, tid);
        else
            solve_row<true, false>(P.pool, row, P.u_ptr, P.u_col, P.u_slot, P.u_diag, P.y, P.x, P.done_u, P.done_l, sD, sR, sV, tid);
    }
}

}  // namespace

int trsv_max_grid(int device) {
    int per_sm = 0, sms = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, trsv_kernel, TR_THREADS, 0) != cudaSuccess) return 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    return per_sm * sms;
}

cudaError_t launch_trsv(const TrsvParams& p, int grid, cudaStream_t stream) {
    TrsvParams pp = p;
    void* args[] = {&pp};
    return cudaLaunchCooperativeKernel((const void*)trsv_kernel, dim3(grid), dim3(TR_THREADS), args, 0, stream);
}

}  // namespace soglu
