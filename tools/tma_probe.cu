// tma_probe.cu - standalone check of the cp.async.bulk.tensor tile load used by k_blur.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <vector>
constexpr int BOXW_MAX = 256, RB = 8;
__constant__ int BOXW;
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void probe(const __grid_constant__ CUtensorMap tmap, int x, int y, int z, uint8_t *out, int mode) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + RB * BOXW_MAX);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (mode & 1) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(RB * BOXW) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                     ::"r"(smem_u32(smem)), "l"(&tmap), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar)) : "memory");
    }
    uint32_t done = 0;
    for (int spin = 0; spin < (1 << 20) && !done; ++spin)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(bar)), "r"(0) : "memory");
    for (int i = threadIdx.x; i < RB * BOXW; i += blockDim.x) out[i] = done ? smem[i] : 0xEE;
}
int main(int argc, char **argv) {
    const int boxw = argc > 1 ? atoi(argv[1]) : 240, xarg = argc > 2 ? atoi(argv[2]) : 76;
    cudaMemcpyToSymbol(BOXW, &boxw, sizeof(int));
    const int W = 1920, H = 1080, B = 4;
    std::vector<uint8_t> h((size_t)W * H * B);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (uint8_t)((i * 2654435761u) >> 24);
    uint8_t *d, *o; cudaMalloc(&d, h.size()); cudaMalloc(&o, RB * BOXW_MAX);
    cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    void *fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    printf("entry point: %s q=%d fn=%p\n", cudaGetErrorString(e), (int)q, fn);
    typedef CUresult (*Enc)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    CUtensorMap map; memset(&map, 0, sizeof(map));
    cuuint64_t dims[3] = {W, H, B}, strides[2] = {W, (cuuint64_t)W * H};
    cuuint32_t box[3] = {(cuuint32_t)boxw, RB, 1}, es[3] = {1, 1, 1};
    CUresult r = ((Enc)fn)(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode: %d\n", (int)r);
    for (int mode = 0; mode < 2; ++mode) {
        const int x = xarg, y = 40, z = 2;
        probe<<<1, 128, RB * BOXW_MAX + 64>>>(map, x, y, z, o, mode);
        e = cudaDeviceSynchronize();
        printf("mode %d: kernel %s\n", mode, cudaGetErrorString(e));
        if (e != cudaSuccess) return 1;
        std::vector<uint8_t> got(RB * boxw); cudaMemcpy(got.data(), o, got.size(), cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int r2 = 0; r2 < RB; ++r2) for (int c = 0; c < boxw; ++c) bad += got[r2 * boxw + c] != h[((size_t)z * H + y + r2) * W + x + c];
        printf("mode %d: mismatches %d (first bytes %02x %02x want %02x %02x)\n", mode, bad, got[0], got[1], h[((size_t)z * H + y) * W + x], h[((size_t)z * H + y) * W + x + 1]);
    }
    return 0;
}
