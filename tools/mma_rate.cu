// mma_rate.cu - micro-benchmark: issue rate of tcgen05.mma (kind::f16, M=128) as a function of N and of the number of
// independent accumulators the MMAs rotate over.  Operands are whatever is in shared memory (values are irrelevant).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/mma_rate tools/mma_rate.cu && tools/bin/mma_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t kdesc(uint32_t saddr, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

__global__ void __launch_bounds__(128, 1) bench(int N, int nacc, int iters, int sbo, int astep, long long* out) {
    extern __shared__ uint8_t raw[];
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    __shared__ uint32_t tmem_addr;
    __shared__ uint64_t bar;
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_addr)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = tmem_addr;
    if (warp == 1 && lane == 0) {
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
        const uint64_t ad = kdesc(base, sbo), bd = kdesc(base + 65536, 1024);
        long long t0 = clock64();
        #pragma unroll 8
        for (int i = 0; i < iters; ++i) {
            // 4 K-steps inside one 128-byte swizzle row, like the conv kernels; `astep` moves A between MMAs (16-B units)
            // nacc is a power of two: no integer division in the issue loop
            mma(tm + (uint32_t)((i & (nacc - 1)) * N), ad + (uint64_t)((i & 3) * 2 + ((i >> 2) & 7) * astep), bd + (uint64_t)((i & 3) * 2), idesc, i >= nacc);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("{\n\t.reg .pred P1;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], 0;\n\t@P1 bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(smem_u32(&bar)) : "memory");
        long long t1 = clock64();
        out[blockIdx.x] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm) : "memory");
}

int main() {
    long long* d;
    cudaMalloc(&d, 148 * sizeof(long long));
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int iters = 4096;
    printf("%-6s %-5s %-6s %-6s %-10s %s\n", "N", "nacc", "sbo", "astep", "grid", "cycles/MMA (max over CTAs)");
    for (int grid : {1, 148})
        for (int N : {16, 64, 256})
            for (int nacc : {1, 2, 4})
                for (int cfg = 0; cfg < 1; ++cfg) {
                    if (N * nacc > 512) continue;
                    int sbo = cfg == 0 ? 1024 : 2048, astep = cfg == 2 ? 8 : 0;
                    bench<<<grid, 128, 190 * 1024>>>(N, nacc, iters, sbo, astep, d);
                    if (cudaDeviceSynchronize() != cudaSuccess) { printf("error %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
                    long long h[148];
                    cudaMemcpy(h, d, grid * sizeof(long long), cudaMemcpyDeviceToHost);
                    long long mx = 0;
                    for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
                    printf("%-6d %-5d %-6d %-6d %-10d %.1f\n", N, nacc, sbo, astep, grid, (double)mx / iters);
                }
    return 0;
}
