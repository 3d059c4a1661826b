// Single-warp instruction latencies on B200 for the diagonal-block kernel's pivot chain (dependent chains, 1 warp active).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I sparse-operator-graph-lu_b200/csrc/device -o lat_bench2 tools/lat_bench2.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace soglu;
#define REP(n, body) _Pragma("unroll 1") for (int i_ = 0; i_ < 100; i_++) { _Pragma("unroll") for (int j_ = 0; j_ < n; j_++) { body; } }
__global__ void k(long long* out, double x, int nwarps) {
    __shared__ __align__(16) double sm[512];
    __shared__ uint64_t bars[8];
    const int t = threadIdx.x, lane = t & 31;
    uint64_t& bar = bars[t >> 5];
    sm[t & 511] = x + t;
    if (t < 8) { ptx::mbar_init(&bars[t], 1); ptx::fence_mbar_init(); }
    __syncthreads();
    if (t >= 32 * nwarps) return;
    long long c[16];
    double a = x, b = x + 1, s = x, d = x;
    int n = 0;
    c[n++] = clock64();
    REP(10, a = fma(a, 1.0000001, 1e-9));                       c[n++] = clock64();   // 0 DFMA
    REP(10, a = a * 1.0000001);                                  c[n++] = clock64();   // 1 DMUL
    REP(10, a = a + 1.0000001);                                  c[n++] = clock64();   // 2 DADD
    REP(10, asm volatile("rcp.approx.ftz.f64 %0, %0;" : "+d"(b)));  c[n++] = clock64();   // 3 MUFU.RCP64H (+ zeroing the low word)
    REP(10, b = ptx::fast_neg_rcp(b));                           c[n++] = clock64();   // 4 seed + 4 DFMA
    REP(10, s = __shfl_sync(0xffffffffu, s, (lane + 1) & 31));   c[n++] = clock64();   // 5 64-bit shuffle
    REP(10, d = (d < 1e-9 && d > -1e-9) ? ((d < 0) ? -1e-9 : 1e-9) : d; d = d * 1.0000001);  c[n++] = clock64();   // 6 clamp + DMUL
    { int idx = lane; REP(10, idx = (int)reinterpret_cast<int*>(sm)[idx & 255] & 255); a += idx; }  c[n++] = clock64();   // 7 LDS.32 dependent (+ LOP)
    REP(10, sm[lane] = d; __syncwarp(); d = sm[(lane + 1) & 31] + d;); c[n++] = clock64();   // 8 STS -> syncwarp -> LDS -> DADD
    { double p0 = x, p1 = x; REP(10, ptx::dmma884(p0, p1, a, b)); a += p0 + p1; }  c[n++] = clock64();   // 9 DMMA dependent
    { uint32_t ph = 0; REP(10, if (lane == 0) ptx::mbar_arrive(&bar); while (!ptx::mbar_try_wait(&bar, ph)) {} ph ^= 1;) }  c[n++] = clock64();   // 10 mbarrier arrive + try_wait (same warp)
    REP(10, asm volatile("bar.sync %0, 32;" ::"r"(3 + (t >> 5))));                    c[n++] = clock64();   // 11 named barrier, one warp
    if (t == 0) for (int i = 0; i + 1 < n; i++) out[i] = c[i + 1] - c[i];
    if (a + b + s + d == 12345.678) out[15] = 1;
}
int main() {
    setvbuf(stdout, nullptr, _IONBF, 0);
    long long* d; cudaMalloc(&d, 128);
    const char* names[] = {"DFMA", "DMUL", "DADD", "MUFU.RCP64H", "fast_neg_rcp (seed + 4 DFMA)", "SHFL 64-bit", "clamp + DMUL", "LDS.32 dependent (+cvt/and)",
                           "STS -> syncwarp -> LDS -> DADD", "DMMA.8x8x4 dependent", "mbarrier arrive + try_wait", "bar.sync (32 threads)"};
    for (int nw : {1, 8}) {
        k<<<1, 256>>>(d, 0.5, nw); k<<<1, 256>>>(d, 0.5, nw);
        long long h[16]; cudaMemcpy(h, d, 128, cudaMemcpyDeviceToHost);
        printf("%d warp(s) active, cycles per dependent step:\n", nw);
        for (int i = 0; i < 12; i++) printf("  %-34s %7.1f\n", names[i], h[i] / 1000.0);
    }
    return 0;
}
