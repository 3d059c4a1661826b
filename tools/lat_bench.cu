// latency microbenchmarks on B200: dependent DFMA / DMUL chain, MUFU.RCP64H+Newton reciprocal,
// LDS round trip, named barrier with 128/256 threads, shuffle.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(long long* out, double x) {
    __shared__ double sm[512];
    int t = threadIdx.x;
    sm[t] = x + t; sm[t + 256] = 1.0;
    __syncthreads();
    double a = x;
    long long c0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 100; i++) {
#pragma unroll
        for (int j = 0; j < 10; j++) a = fma(a, 1.0000001, 1e-9);
    }
    long long c1 = clock64();
    double b = x;
#pragma unroll 1
    for (int i = 0; i < 100; i++) {
#pragma unroll
        for (int j = 0; j < 10; j++) b = 1.0 / (b + 1.5);
    }
    long long c2 = clock64();
    int idx = t;
#pragma unroll 1
    for (int i = 0; i < 1000; i++) idx = (int)sm[(idx & 255)] & 255;
    long long c3 = clock64();
#pragma unroll 1
    for (int i = 0; i < 1000; i++) asm volatile("bar.sync 1, 256;");
    long long c4 = clock64();
    if (t < 128) {
#pragma unroll 1
        for (int i = 0; i < 1000; i++) asm volatile("bar.sync 2, 128;");
    }
    long long c5 = clock64();
    double s = x;
#pragma unroll 1
    for (int i = 0; i < 1000; i++) s = __shfl_sync(0xffffffffu, s, (t + 1) & 31) + 1.0;
    long long c6 = clock64();
    // store -> barrier -> load -> dfma -> store chain (what a pivot step does)
    double v = x;
#pragma unroll 1
    for (int i = 0; i < 1000; i++) {
        if (t == (i & 255)) sm[i & 63] = v;
        asm volatile("bar.sync 1, 256;");
        v = fma(sm[i & 63], 1.0000001, v);
    }
    long long c7 = clock64();
    if (t == 0) {
        out[0] = (c1 - c0); out[1] = (c2 - c1); out[2] = (c3 - c2); out[3] = (c4 - c3); out[4] = (c5 - c4); out[5] = (c6 - c5); out[6] = c7 - c6;
    }
    if (a + b + idx + s + v == 12345.678) out[7] = 1;
}
int main() {
    long long* d; cudaMalloc(&d, 64);
    k<<<1, 256>>>(d, 0.5); k<<<1, 256>>>(d, 0.5);
    long long h[8]; cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
    printf("DFMA dependent latency: %.1f clk\n1/x (IEEE double) latency: %.1f clk\nLDS dependent (incl cvt/and): %.1f clk\nbar.sync 256 thr: %.1f clk\nbar.sync 128 thr: %.1f clk\nshfl+dadd: %.1f clk\nSTS->bar->LDS->DFMA step: %.1f clk\n",
           h[0] / 1000.0, h[1] / 1000.0, h[2] / 1000.0, h[3] / 1000.0, h[4] / 1000.0, h[5] / 1000.0, h[6] / 1000.0);
    return 0;
}
