// mma_rate.cu - microbenchmark: issue rate of tcgen05.mma for kind::i8 / kind::f8f6f4 / kind::f16 on this GPU.
// One CTA per SM, one thread issues `reps` MMAs (M=128, N, K=32 bytes) that accumulate into 1, 2 or 4 TMEM tiles in
// turn, commits, waits; cycles per MMA from clock64.   nvcc -gencode arch=compute_100a,code=sm_100a -o mma_rate mma_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t sbo, uint32_t layout) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)1 << 16) | ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | ((uint64_t)1 << 46) | ((uint64_t)layout << 61);
}
template <int KIND> __device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    if (KIND == 0)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
                     ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc), "r"(0u) : "memory");
    else if (KIND == 1)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
                     ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc), "r"(0u) : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
                     ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc), "r"(0u) : "memory");
}
template <int KIND> __global__ void __launch_bounds__(128, 1) rate_kernel(int N, int reps, int ntiles, long long *out) {
    extern __shared__ unsigned char raw[];
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    for (int i = threadIdx.x; i < 16384; i += 128) ((uint32_t *)(raw + (base - smem_u32(raw))))[i] = 0x01010101u * (i & 3);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    if (threadIdx.x == 0) {
        // instruction descriptor: D fmt (i8: s32 = 2, else f32 = 1), A/B formats 0, K-major, N, M = 128
        const uint32_t idesc = ((KIND == 0 ? 2u : 1u) << 4) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
        const uint64_t a = smem_desc(base, 1024, 2), b = smem_desc(base + 16384, 1024, 2);
        const long long t0 = clock64();
        // lean issue loop (a single thread issues: every scalar instruction between two MMAs costs its full latency)
        const uint32_t d0 = tmem, d1 = tmem + (ntiles > 1 ? 256u : 0u);
        mma<KIND>(d0, a, b, idesc, 0u);
        mma<KIND>(d1, a, b, idesc, ntiles > 1 ? 0u : 1u);
#pragma unroll 1
        for (int i = 2; i < reps; i += 8) {
            mma<KIND>(d0, a, b, idesc, 1u); mma<KIND>(d1, a, b, idesc, 1u); mma<KIND>(d0, a, b, idesc, 1u); mma<KIND>(d1, a, b, idesc, 1u);
            mma<KIND>(d0, a, b, idesc, 1u); mma<KIND>(d1, a, b, idesc, 1u); mma<KIND>(d0, a, b, idesc, 1u); mma<KIND>(d1, a, b, idesc, 1u);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        uint32_t done = 0;
        for (int spin = 0; spin < (1 << 24) && !done; ++spin)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
        const long long t1 = clock64();
        if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = done; }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}
template <int KIND> void run(const char *name, int N, int ntiles, long long *d_out) {
    const int reps = 258;
    cudaFuncSetAttribute(rate_kernel<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, 70000);
    for (int w = 0; w < 2; ++w) rate_kernel<KIND><<<148, 128, 70000>>>(N, reps, ntiles, d_out);
    long long h[2] = {0, 0};
    cudaError_t e = cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
    printf("%-8s N=%3d tiles=%d: %8.1f cycles per MMA (ideal N/2 = %d for 8-bit, N for 16-bit)  done=%lld  %s\n", name, N, ntiles, (double)h[0] / reps, N / 2, h[1],
           e == cudaSuccess ? "" : cudaGetErrorString(e));
}
int main() {
    long long *d_out;
    cudaMalloc(&d_out, 16);
    for (int N : {64, 128, 256})
        for (int nt : {1, 2}) {
            if (N * nt > 512) continue;
            run<0>("i8", N, nt, d_out);
            run<1>("f8f6f4", N, nt, d_out);
            run<2>("f16", N, nt, d_out);
        }
    return 0;
}
