// Operand-delivery microbenchmark for the Schur-update path (B200, sm_100a): how fast can 148 persistent CTAs pull
// 64x64 FP64 blocks (34 816 B each, the executor's layout) into shared memory with cp.async.bulk through a 3-stage
// ring -- from an L2-resident pool, from an HBM-sized pool, and while the FP64 tensor cores are busy.  The executor
// needs 2 x 34 816 B per 2.08 us per SM (= 5.0 TB/s over 148 SMs) to run its DMMA loop at the peak rate; it reaches
// 2.38 us per pair.  This tool says whether the L2 -> SM path can deliver that.
//
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o operand_bw operand_bw.cu && ./operand_bw
//
// Output per case: GB/s over all SMs and us per operand pair per SM, with and without a concurrent DMMA load.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int BLK_BYTES = 34816, STAGES = 3, THREADS = 288;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// warp 0: producer, streams `pairs` operand pairs (two blocks each) from pseudo-random slots of the pool; warps 1..8:
// consumers; dmma_per_pair > 0 makes each consumer warp issue that many register-resident DMMAs per pair (128 = what the
// executor's loop issues per warp and pair), dmma_per_pair = 0 measures pure delivery.
__global__ void __launch_bounds__(THREADS, 1) k_stream(const char* pool, uint32_t n_slots, int pairs, int dmma_per_pair, double* sink) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)STAGES * 2 * BLK_BYTES);
    uint64_t* empty = full + STAGES;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (warp == 0) {
        if (lane == 0) {
            uint32_t x = 0x9e3779b9u * (blockIdx.x + 1);
            for (int it = 0; it < pairs; it++) {
                const int s = it % STAGES;
                mbar_wait(&empty[s], ((it / STAGES) & 1) ^ 1);
                x = x * 1664525u + 1013904223u; const uint32_t a = (x >> 8) % n_slots;
                x = x * 1664525u + 1013904223u; const uint32_t b = (x >> 8) % n_slots;
                mbar_expect(&full[s], 2 * BLK_BYTES);
                bulk_g2s(smem + (size_t)s * 2 * BLK_BYTES, pool + (size_t)a * BLK_BYTES, BLK_BYTES, &full[s]);
                bulk_g2s(smem + (size_t)s * 2 * BLK_BYTES + BLK_BYTES, pool + (size_t)b * BLK_BYTES, BLK_BYTES, &full[s]);
            }
        }
        return;
    }
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; i++) { c[i][0] = i; c[i][1] = lane; }
    const double av = 1.0000001, bv = 0.9999999;
    for (int it = 0; it < pairs; it++) {
        const int s = it % STAGES;
        mbar_wait(&full[s], (it / STAGES) & 1);
        for (int k = 0; k < dmma_per_pair; k += 8) {
#pragma unroll
            for (int i = 0; i < 8; i++) mma884(c[i][0], c[i][1], av, bv);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
    }
    double sum = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) sum += c[i][0] + c[i][1];
    if (sum == 1.2345e-300) sink[0] = sum;
}

int main() {
    int dev = 0, sms = 0;
    CK(cudaSetDevice(dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const size_t smem = (size_t)STAGES * 2 * BLK_BYTES + 64;
    CK(cudaFuncSetAttribute(k_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    double* sink;
    CK(cudaMalloc(&sink, 8));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const size_t pool_sizes[] = {size_t(48) << 20, size_t(8) << 30};      // fits the L2 / HBM-sized
    const char* pool_names[] = {"48 MB pool (L2 resident)", "8 GB pool (HBM)"};
    for (int ps = 0; ps < 2; ps++) {
        const uint32_t n_slots = (uint32_t)(pool_sizes[ps] / BLK_BYTES);
        char* pool;
        CK(cudaMalloc(&pool, (size_t)n_slots * BLK_BYTES));
        CK(cudaMemset(pool, 0, (size_t)n_slots * BLK_BYTES));
        for (int dm : {0, 128}) {
            const int pairs = 4000;
            k_stream<<<sms, THREADS, smem>>>(pool, n_slots, 200, dm, sink);   // warm-up (fills the L2 for the small pool)
            CK(cudaDeviceSynchronize());
            CK(cudaEventRecord(e0));
            k_stream<<<sms, THREADS, smem>>>(pool, n_slots, pairs, dm, sink);
            CK(cudaEventRecord(e1));
            CK(cudaDeviceSynchronize());
            float ms = 0;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            const double bytes = (double)sms * pairs * 2 * BLK_BYTES;
            printf("%-26s %s: %8.1f GB/s, %.3f us per pair per SM%s\n", pool_names[ps], dm ? "with 128 DMMA / warp / pair" : "delivery only              ",
                   bytes / (ms * 1e-3) * 1e-9, ms * 1e3 / pairs, dm ? "   (2.08 us = DMMA peak)" : "");
        }
        CK(cudaFree(pool));
    }
    return 0;
}
