// How long does it take a polling CTA to see a flag another CTA has just stored?  (the "queue wait" of one executor hop)
// CTA 0 (one thread) stores flag[i] = i at a timed instant; CTAs 1.. poll their flag with the given load flavour and record
// %globaltimer when they see it.  Reports the mean / max latency over many rounds, for 1 and 147 polling CTAs.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o poll_lat tools/poll_lat.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
__device__ __forceinline__ unsigned long long gt() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ int ld_acq(const int* p) { int v; asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ int ld_rlx(const int* p) { int v; asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ int ld_vol(const int* p) { int v; asm volatile("ld.volatile.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
template <int MODE>
__global__ void k(int* flags, unsigned long long* t_store, unsigned long long* t_seen, int rounds, int with_fence) {
    if (threadIdx.x != 0) return;
    const int b = blockIdx.x, nb = gridDim.x;
    if (b == 0) {
        for (int r = 1; r <= rounds; r++) {
            unsigned long long t0 = gt();
            while (gt() - t0 < 20000) {}                 // 20 us apart: every poller is spinning again
            for (int c = 1; c < nb; c++) {
                if (with_fence) asm volatile("fence.acq_rel.gpu;" ::: "memory");
                t_store[(size_t)r * nb + c] = gt();
                asm volatile("st.relaxed.gpu.global.s32 [%0], %1;" ::"l"(flags + 32 * c), "r"(r) : "memory");
            }
        }
    } else {
        for (int r = 1; r <= rounds; r++) {
            int v;
            do { v = MODE == 0 ? ld_acq(flags + 32 * b) : (MODE == 1 ? ld_rlx(flags + 32 * b) : ld_vol(flags + 32 * b)); } while (v < r);
            t_seen[(size_t)r * nb + b] = gt();
        }
    }
}
int main() {
    const int rounds = 200;
    for (int nb : {2, 148}) {
        int* flags; unsigned long long *ts, *tn;
        cudaMalloc(&flags, 148 * 128); cudaMalloc(&ts, sizeof(unsigned long long) * (rounds + 1) * nb); cudaMalloc(&tn, sizeof(unsigned long long) * (rounds + 1) * nb);
        for (int mode = 0; mode < 3; mode++)
            for (int fence = 0; fence < 2; fence++) {
                cudaMemset(flags, 0, 148 * 128);
                void* args[] = {&flags, &ts, &tn, (void*)&rounds, &fence};
                const void* fn = mode == 0 ? (const void*)k<0> : (mode == 1 ? (const void*)k<1> : (const void*)k<2>);
                cudaLaunchCooperativeKernel(fn, dim3(nb), dim3(32), args, 0, 0);
                if (cudaDeviceSynchronize() != cudaSuccess) { printf("error\n"); return 1; }
                unsigned long long* hs = new unsigned long long[(rounds + 1) * nb], *hn = new unsigned long long[(rounds + 1) * nb];
                cudaMemcpy(hs, ts, sizeof(unsigned long long) * (rounds + 1) * nb, cudaMemcpyDeviceToHost);
                cudaMemcpy(hn, tn, sizeof(unsigned long long) * (rounds + 1) * nb, cudaMemcpyDeviceToHost);
                double sum = 0, mx = 0; long cnt = 0;
                for (int r = 10; r <= rounds; r++) for (int c = 1; c < nb; c++) { double d = (double)hn[(size_t)r * nb + c] - (double)hs[(size_t)r * nb + c]; sum += d; if (d > mx) mx = d; cnt++; }
                printf("%3d polling CTAs, %-12s%s: store -> seen %.0f ns mean, %.0f ns max\n", nb - 1, mode == 0 ? "ld.acquire" : (mode == 1 ? "ld.relaxed" : "ld.volatile"), fence ? " (fence before each store)" : "", sum / cnt, mx);
                delete[] hs; delete[] hn;
            }
    }
    return 0;
}
