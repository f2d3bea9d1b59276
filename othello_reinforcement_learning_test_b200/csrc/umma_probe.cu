// umma_probe.cu -- diagnostic entry point: run one accumulation chain of tcgen05.mma on a
// caller-supplied shared-memory image with caller-supplied descriptor fields, and return the
// fp32 accumulator.  tests/test_umma_probe.py uses it to pin the shared-memory descriptor
// semantics the conv kernel relies on (no-swizzle K-major, 16-byte row shifts, SBO = 144 B).
#include "common.cuh"
#include "tc_ptx.cuh"

namespace oth {
namespace tc {

__global__ void __launch_bounds__(128, 1)
k_umma_probe(const uint4* __restrict__ image, int image_units, int n, int k_steps, uint32_t a_off, uint32_t a_lbo,
             uint32_t a_sbo, uint32_t a_kstep, uint32_t b_off, uint32_t b_lbo, uint32_t b_sbo, uint32_t b_kstep,
             float* __restrict__ d_out)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    uint4* img = reinterpret_cast<uint4*>(smem);
    for (int i = threadIdx.x; i < image_units; i += blockDim.x) img[i] = image[i];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t cols = 32;
    while ((int)cols < n) cols <<= 1;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_async_proxy();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    if (threadIdx.x == 0) {
        const uint32_t base = smem_u32(smem);
        const uint32_t idesc = umma_idesc(n);
        for (int k = 0; k < k_steps; ++k) {
            const uint64_t ad = umma_desc(base + a_off + k * a_kstep, a_lbo, a_sbo);
            const uint64_t bd = umma_desc(base + b_off + k * b_kstep, b_lbo, b_sbo);
            umma_bf16(tmem, ad, bd, idesc, k > 0 ? 1u : 0u);
        }
        umma_commit(&bar);
    }
    __syncwarp();
    mbar_wait(&bar, 0);
    tc_fence_after();
    const int m = warp * 32 + lane;
    for (int c0 = 0; c0 < n; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, r);
        for (int j = 0; j < 32 && c0 + j < n; ++j) d_out[(size_t)m * n + c0 + j] = __uint_as_float(r[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(cols) : "memory");
    }
}

}  // namespace tc
}  // namespace oth

using namespace oth;

extern "C" int oth_debug_umma_probe(oth_ctx* ctx, const void* smem_image, int image_bytes, int n, int k_steps,
                                    uint32_t a_off, uint32_t a_lbo, uint32_t a_sbo, uint32_t a_kstep, uint32_t b_off,
                                    uint32_t b_lbo, uint32_t b_sbo, uint32_t b_kstep, float* d_out)
{
    OTH_REQUIRE(ctx && smem_image && d_out, OTH_ERR_ARG, "oth_debug_umma_probe: NULL argument");
    OTH_REQUIRE(image_bytes > 0 && image_bytes % 16 == 0 && image_bytes <= 200 * 1024, OTH_ERR_ARG, "image_bytes must be a multiple of 16, <= 200 KiB");
    OTH_REQUIRE(n >= 16 && n <= 256 && n % 16 == 0 && k_steps >= 1 && k_steps <= 64, OTH_ERR_ARG, "bad n / k_steps");
    OTH_CHECK_CUDA(cudaSetDevice(ctx->device));
    void* d_img = nullptr; float* d_d = nullptr;
    OTH_CHECK_CUDA(cudaMalloc(&d_img, image_bytes));
    OTH_CHECK_CUDA(cudaMalloc((void**)&d_d, (size_t)128 * n * 4));
    OTH_CHECK_CUDA(cudaMemcpyAsync(d_img, smem_image, image_bytes, cudaMemcpyHostToDevice, ctx->stream));
    OTH_CHECK_CUDA(cudaFuncSetAttribute(tc::k_umma_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, image_bytes));
    tc::k_umma_probe<<<1, 128, image_bytes, ctx->stream>>>((const uint4*)d_img, image_bytes / 16, n, k_steps, a_off, a_lbo,
                                                           a_sbo, a_kstep, b_off, b_lbo, b_sbo, b_kstep, d_d);
    ctx->launches++;
    OTH_CHECK_CUDA(cudaGetLastError());
    OTH_CHECK_CUDA(cudaMemcpyAsync(d_out, d_d, (size_t)128 * n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_img); cudaFree(d_d);
    OTH_CHECK_CUDA(e);
    return OTH_OK;
}
