// FP64 peak microbenchmark for B200 (sm_100a): register-resident DFMA versus the warp-level
// FP64 tensor-core MMA shapes (mma.sync ... f64).  tcgen05.mma has no f64 kind, so these are
// the only FP64 matrix paths.  Used to (a) fix the roofline denominator P64 for the
// Schur-update kernel and (b) decide DFMA vs DMMA for its inner product.
//
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_peak fp64_peak.cu && ./fp64_peak
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

template <int ACC>
__global__ void __launch_bounds__(256) k_dfma(double* out, int iters, double a, double b) {
    double acc[ACC];
#pragma unroll
    for (int i = 0; i < ACC; i++) acc[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ACC; i++) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ACC; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ void mma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void mma1684(double (&c)[4], const double (&a)[2], double b) {
    asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(b));
}
__device__ __forceinline__ void mma1688(double (&c)[4], const double (&a)[4], const double (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void mma16816(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

template <int ACC>
__global__ void __launch_bounds__(256) k_mma884(double* out, int iters, double a, double b) {
    double c[ACC][2];
#pragma unroll
    for (int i = 0; i < ACC; i++) { c[i][0] = i; c[i][1] = threadIdx.x; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ACC; i++) mma884(c[i][0], c[i][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ACC; i++) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ACC, int K>
__global__ void __launch_bounds__(256) k_mma168(double* out, int iters, double av, double bv) {
    double c[ACC][4];
    double a[8], b[4];
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = av + i;
#pragma unroll
    for (int i = 0; i < 4; i++) b[i] = bv + i;
#pragma unroll
    for (int i = 0; i < ACC; i++) { c[i][0] = i; c[i][1] = threadIdx.x; c[i][2] = 1; c[i][3] = 2; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ACC; i++) {
            if (K == 4) { double a2[2] = {a[0], a[1]}; mma1684(c[i], a2, b[0]); }
            else if (K == 8) { double a4[4] = {a[0], a[1], a[2], a[3]}; double b2[2] = {b[0], b[1]}; mma1688(c[i], a4, b2); }
            else mma16816(c[i], a, b);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ACC; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static double timeit(F launch, int reps) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    launch();
    launch();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; r++) {
        CK(cudaEventRecord(e0));
        launch();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    return best * 1e-3;
}

int main(int argc, char** argv) {
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    printf("device %s sms %d clock %d kHz\n", p.name, sms, p.clockRate);
    double* out;
    int ctas_per_sm = 8;
    int grid = sms * ctas_per_sm;
    CK(cudaMalloc(&out, sizeof(double) * grid * 256));
    int iters = 20000;
    long sustained = argc > 1 ? atol(argv[1]) : 0;   // seconds of sustained loop for the chosen best
    {
        double t = timeit([&] { k_dfma<16><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); }, 5);
        printf("DFMA  acc16  : %.2f TFLOP/s (%.3f ms)\n", 2.0 * 16 * iters * grid * 256 / t * 1e-12, t * 1e3);
        t = timeit([&] { k_dfma<32><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); }, 5);
        printf("DFMA  acc32  : %.2f TFLOP/s (%.3f ms)\n", 2.0 * 32 * iters * grid * 256 / t * 1e-12, t * 1e3);
    }
    {
        double t = timeit([&] { k_mma884<8><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); }, 5);
        printf("DMMA m8n8k4  acc8 : %.2f TFLOP/s (%.3f ms)\n", 2.0 * 256 * 8 * iters * (double)grid * 8 / t * 1e-12, t * 1e3);
        t = timeit([&] { k_mma884<16><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); }, 5);
        printf("DMMA m8n8k4  acc16: %.2f TFLOP/s (%.3f ms)\n", 2.0 * 256 * 16 * iters * (double)grid * 8 / t * 1e-12, t * 1e3);
    }
    {
        double t = timeit([&] { k_mma168<8, 4><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); }, 5);
        printf("DMMA m16n8k4 acc8 : %.2f TFLOP/s (%.3f ms)\n", 2.0 * 16 * 8 * 4 * 8 * iters * (double)grid * 8 / t * 1e-12, t * 1e3);
        t = timeit([&] { k_mma168<8, 8><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); }, 5);
        printf("DMMA m16n8k8 acc8 : %.2f TFLOP/s (%.3f ms)\n", 2.0 * 16 * 8 * 8 * 8 * iters * (double)grid * 8 / t * 1e-12, t * 1e3);
        t = timeit([&] { k_mma168<8, 16><<<grid, 256>>>(out, iters / 2, 1.0000001, 1e-9); }, 5);
        printf("DMMA m16n8k16 acc8: %.2f TFLOP/s (%.3f ms)\n", 2.0 * 16 * 8 * 16 * 8 * (iters / 2) * (double)grid * 8 / t * 1e-12, t * 1e3);
    }
    if (sustained > 0) {
        // sustained DFMA under the power cap: loop for `sustained` seconds, report the mean
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0));
        CK(cudaEventCreate(&e1));
        CK(cudaEventRecord(e0));
        int n = 0;
        double total = 0;
        while (total < (double)sustained) {
            for (int r = 0; r < 20; r++) k_dfma<16><<<grid, 256>>>(out, iters, 1.0000001, 1e-9);
            n += 20;
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            total = ms * 1e-3;
        }
        printf("DFMA sustained %.1f s: %.2f TFLOP/s\n", total, 2.0 * 16 * iters * (double)grid * 256 * n / total * 1e-12);
        CK(cudaEventRecord(e0));
        n = 0; total = 0;
        while (total < (double)sustained) {
            for (int r = 0; r < 20; r++) k_mma884<8><<<grid, 256>>>(out, iters, 1.0000001, 1e-9);
            n += 20;
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            total = ms * 1e-3;
        }
        printf("DMMA m8n8k4 sustained %.1f s: %.2f TFLOP/s\n", total, 2.0 * 256 * 8 * iters * (double)grid * 8 * n / total * 1e-12);
    }
    CK(cudaFree(out));
    return 0;
}
