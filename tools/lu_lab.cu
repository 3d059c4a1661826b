// GPU lab for the diagonal-block kernel (csrc/device/lu_blocked.cuh): one CTA shaped like the executor's (32 idle
// producer threads + 256 math threads), the block resident in shared memory, cycles per call from clock64 and a
// per-warp timeline of the phases (SWEEP/STRIPS end, barrier X, TRAIL end per panel).  Checks the factors against a
// host LU and the inverses against L^-1 L = I, U U^-1 = I.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I sparse-operator-graph-lu_b200/csrc/device -o lu_lab tools/lu_lab.cu
// -DSOGLU_LUB_PROF adds the per-warp phase timeline (its clock reads + global stores cost a few hundred cycles per call)
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "lu_blocked.cuh"

using namespace soglu::lub;

template <bool WITH_INV>
__global__ void __launch_bounds__(288, 1) lab_kernel(const double* A, double* S_out, double* W_out, int iters, long long* cycles, long long* prof) {
    extern __shared__ __align__(128) unsigned char smem[];
    double* S = reinterpret_cast<double*>(smem);
    double* W = S + 64 * LD;
    double* scr = W + 64 * LD;
#if defined(SOGLU_LUB_PROF)
    if (threadIdx.x == 0) hw::g_lub_prof = prof;
#endif
    __syncthreads();
    if (threadIdx.x < 32) return;
    const int ct = threadIdx.x - 32;
    lu_setup(scr, ct);
    long long total = 0, best = 1ll << 60;
    for (int it = 0; it < iters; it++) {
        for (int e = ct; e < 64 * 64; e += 256) S[(e >> 6) * LD + (e & 63)] = A[e];
#if defined(SOGLU_LUB_DEBUG)
        if (ct == 0) { hw::g_lub_iter = it; printf("call %d: barriers at smem %u\n", it, soglu::ptx::smem_u32(scr + SCR_BAR)); }
#endif
        hw::sync_math();
        const long long c0 = clock64();
        lu_blocked<WITH_INV, false, true>(S, W, scr, ct);
        hw::sync_math();
        const long long c1 = clock64();
        total += c1 - c0;
        if (c1 - c0 < best) best = c1 - c0;
        if (ct == 0) prof[120] = c0;
    }
    if (ct == 0) { cycles[0] = total / iters; cycles[1] = best; }
    for (int e = ct; e < 64 * 64; e += 256) { S_out[e] = S[(e >> 6) * LD + (e & 63)]; if (WITH_INV) W_out[e] = W[(e >> 6) * LD + (e & 63)]; }
}

// the 16-pivot sweep of warp 0 alone (nobody waits on its barriers), with ablations
template <int ABL>
__global__ void __launch_bounds__(288, 1) sweep_kernel(const double* A, long long* cycles) {
    extern __shared__ __align__(128) unsigned char smem[];
    double* S = reinterpret_cast<double*>(smem);
    double* scr = S + 2 * 64 * LD;
    if (threadIdx.x < 32) return;
    const int ct = threadIdx.x - 32;
    lu_setup(scr, ct);
    for (int e = ct; e < 64 * 64; e += 256) S[(e >> 6) * LD + (e & 63)] = A[e];
    hw::sync_math();
    if (ct >= 32) return;
    long long best = 1ll << 60;
    for (int it = 0; it < 20; it++) {
        const long long c0 = clock64();
        diag_sweep<false, ABL>(S, scr, it & 3, ct);
        const long long c1 = clock64();
        if (c1 - c0 < best) best = c1 - c0;
    }
    if (ct == 0) cycles[0] = best;
}

static double clampLU(double p) { return (p < 1e-9 && p > -1e-9) ? ((p < 0) ? -1e-9 : 1e-9) : p; }

int main() {
    std::vector<double> A(4096);
    unsigned long long st = 4242;
    auto rnd = [&]() { st = st * 6364136223846793005ull + 1442695040888963407ull; return ((st >> 11) * (1.0 / 9007199254740992.0)) * 2 - 1; };
    for (auto& v : A) v = rnd();
    for (int i = 0; i < 64; i++) A[i * 64 + i] += 40.0;
    std::vector<double> ref = A;
    for (int k = 0; k < 64; k++) {
        const double p = clampLU(ref[k * 64 + k]);
        ref[k * 64 + k] = p;
        for (int i = k + 1; i < 64; i++) {
            const double l = ref[i * 64 + k] / p;
            ref[i * 64 + k] = l;
            for (int j = k + 1; j < 64; j++) ref[i * 64 + j] -= l * ref[k * 64 + j];
        }
    }
    double *dA, *dS, *dW; long long *dC, *dP;
    cudaMalloc(&dA, 32768); cudaMalloc(&dS, 32768); cudaMalloc(&dW, 32768); cudaMalloc(&dC, 64); cudaMalloc(&dP, 1024);
    cudaMemcpy(dA, A.data(), 32768, cudaMemcpyHostToDevice);
    const size_t smem = (2 * 64 * LD + SCRATCH_DOUBLES) * sizeof(double);
    cudaFuncSetAttribute(lab_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(lab_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    {
        long long c;
        auto run = [&](auto kern, const char* what) {
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            kern<<<1, 288, smem>>>(dA, dC);
            cudaError_t e = cudaDeviceSynchronize();
            cudaMemcpy(&c, dC, 8, cudaMemcpyDeviceToHost);
            std::printf("sweep of one 16x16 diagonal block, %-46s %6lld cycles (%5.1f per pivot) %s\n", what, c, c / 16.0, e == cudaSuccess ? "" : cudaGetErrorString(e));
        };
        run(sweep_kernel<0>, "full:");
        run(sweep_kernel<1>, "no publication:");
        run(sweep_kernel<3>, "no publication, no row broadcast:");
        run(sweep_kernel<7>, "no publication, no broadcast, no reciprocal:");
        run(sweep_kernel<15>, "none of the four (FMAs + stores only):");
        run(sweep_kernel<4>, "no reciprocal:");
        run(sweep_kernel<8>, "no pivot shuffle:");
        run(sweep_kernel<2>, "no row broadcast:");
    }
    for (int inv = 1; inv >= 0; inv--) {
        cudaMemset(dP, 0, 1024);
        if (inv) lab_kernel<true><<<1, 288, smem>>>(dA, dS, dW, 50, dC, dP);
        else lab_kernel<false><<<1, 288, smem>>>(dA, dS, dW, 50, dC, dP);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { std::printf("CUDA error: %s\n", cudaGetErrorString(e)); return 2; }
        long long cyc[2], prof[128];
        std::vector<double> S(4096), W(4096);
        cudaMemcpy(cyc, dC, 16, cudaMemcpyDeviceToHost);
        cudaMemcpy(prof, dP, 1024, cudaMemcpyDeviceToHost);
        cudaMemcpy(S.data(), dS, 32768, cudaMemcpyDeviceToHost);
        cudaMemcpy(W.data(), dW, 32768, cudaMemcpyDeviceToHost);
        double e_lu = 0, scale = 0;
        for (int i = 0; i < 4096; i++) { e_lu = std::fmax(e_lu, std::fabs(S[i] - ref[i])); scale = std::fmax(scale, std::fabs(ref[i])); }
        std::printf("%s: %lld cycles avg, %lld best   factors vs host LU %.2e", inv ? "lu + L^-1 + U^-1" : "lu only        ", cyc[0], cyc[1], e_lu / scale);
        if (inv) {
            double eli = 0, eui = 0;
            for (int i = 0; i < 64; i++) for (int j = 0; j < 64; j++) {
                double s = 0, u = 0;
                for (int k = 0; k < 64; k++) {
                    const double li = k < i ? W[i * 64 + k] : (k == i ? 1.0 : 0.0), l = j < k ? S[k * 64 + j] : (j == k ? 1.0 : 0.0);
                    s += li * l;
                    const double uu = k >= i ? S[i * 64 + k] : 0.0, ui = j >= k ? W[k * 64 + j] : 0.0;
                    u += uu * ui;
                }
                eli = std::fmax(eli, std::fabs(s - (i == j))); eui = std::fmax(eui, std::fabs(u - (i == j)));
            }
            std::printf("  |L^-1 L - I| %.2e  |U U^-1 - I| %.2e", eli, eui);
            if (eli > 1e-10) {
                // where: compare with the inverse of the host L (forward substitution per column)
                int shown = 0;
                for (int c = 0; c < 64 && shown < 12; c++) {
                    double x[64];
                    for (int i = 0; i < 64; i++) x[i] = (i == c);
                    for (int k = 0; k < 64; k++) for (int i = k + 1; i < 64; i++) x[i] -= ref[i * 64 + k] * x[k];
                    for (int i = c + 1; i < 64 && shown < 12; i++)
                        if (std::fabs(W[i * 64 + c] - x[i]) > 1e-10) { std::printf("\n    L^-1(%d,%d) = %.6e, expected %.6e", i, c, W[i * 64 + c], x[i]); shown++; }
                }
            }
        }
#if defined(SOGLU_LUB_PROF)
        std::printf("\n  timeline of the last call (cycles since start; rows = phase, columns = warps 0..7)\n");
        const char* names[12] = {"entry", "p0 sweep/strips done", "p0 after X", "p0 trail done", "p1 sweep/strips done", "p1 after X", "p1 trail done",
                                 "p2 sweep/strips done", "p2 after X", "p2 trail done", "p3 sweep/strips done", "p3 after X"};
        for (int s = 0; s < 12; s++) {
            std::printf("  %-22s", names[s]);
            for (int w = 0; w < 8; w++) std::printf(" %6lld", prof[s * 8 + w] ? prof[s * 8 + w] - prof[120] : -1);
            std::printf("\n");
        }
#else
        std::printf("\n");
#endif
    }
    return 0;
}
