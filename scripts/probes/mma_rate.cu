// Probe: issue rate of tcgen05.mma.cta_group::1.kind::i8 (M=128, K=32, A in tensor
// memory, B in shared memory, 128B swizzle) as a function of N.  One CTA per SM,
// one thread issues `reps` x 64 MMAs; prints cycles per MMA.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t desc_of(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

template <int N, int NACC>
__global__ void __launch_bounds__(128, 1) probe(long long *out, int reps) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    unsigned char *sB = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 256 * 128; i += 128) sB[i] = (unsigned char)(i * 7);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;\n" ::"r"(smem_u32(&slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;\n");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n");
    const uint32_t tm = slot;
    if (tid == 0) {
        const uint32_t idesc = (2u << 4) | (0u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint64_t bd0 = desc_of(smem_u32(sB));
        const uint32_t acol0 = tm + NACC * N;          // A behind the accumulators
        long long t0 = clock64();
        uint32_t parity = 0;
        for (int r = 0; r < reps; ++r) {
#pragma unroll
            for (int i = 0; i < 64; ++i) {
                const uint32_t dcol = tm + (i % NACC) * N;
                const uint32_t acol = acol0 + (i & 3) * 8;
                const uint64_t bd = bd0 + (uint64_t)((i & 3) * 2);
                asm volatile(
                    "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                    "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, {%5, %6, %7, %8}, p;\n\t}\n"
                    ::"r"(dcol), "r"(acol), "l"(bd), "r"(idesc), "r"(1u), "r"(0u), "r"(0u), "r"(0u), "r"(0u) : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(&bar)) : "memory");
            uint32_t done = 0;
            while (!done)
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                             : "=r"(done) : "r"(smem_u32(&bar)), "r"(parity) : "memory");
            parity ^= 1;
        }
        long long t1 = clock64();
        if (blockIdx.x == 0) out[0] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(tm));
}

template <int N, int NACC>
void run(long long *d, int grid) {
    const int reps = 200;
    cudaFuncSetAttribute(probe<N, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 256 * 128 + 2048);
    probe<N, NACC><<<grid, 128, 256 * 128 + 2048>>>(d, 10);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    probe<N, NACC><<<grid, 128, 256 * 128 + 2048>>>(d, reps);
    cudaEventRecord(b);
    cudaError_t e = cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, a, b);
    long long cyc = 0; cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
    printf("N=%3d accumulators=%d grid=%3d: %.1f cycles per MMA (%.1f ns), %s\n", N, NACC, grid,
           (double)cyc / (reps * 64.0), ms * 1e6 / (reps * 64.0), cudaGetErrorString(e));
}

int main() {
    long long *d; cudaMalloc(&d, 64);
    run<64, 1>(d, 148); run<80, 1>(d, 148); run<96, 1>(d, 148); run<128, 1>(d, 148);
    run<160, 1>(d, 148); run<192, 1>(d, 148); run<256, 1>(d, 148);
    run<80, 4>(d, 148); run<80, 4>(d, 1); run<64, 4>(d, 148); run<128, 2>(d, 148);
    return 0;
}
