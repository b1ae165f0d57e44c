// Unit test of the CTA-pair MMA path (tcgen05.mma.cta_group::2): D[256,N] = A[256,K] . B[N,K]^T with
// one cluster of two CTAs.  CTA r holds rows 128 r .. of A, rows (N/2) r .. of B (its half of the B
// operand) and receives rows 128 r .. of D in its own TMEM; the leader CTA issues the MMAs for both.
#include "tc_common.cuh"

#include <string.h>

namespace {

using namespace tc;

struct Bars2 {
    uint64_t ops_ready;     // leader: both CTAs have written their operands (2 arrivals)
    uint64_t done;          // each CTA: the MMAs have completed (multicast commit)
    uint32_t tmem_base;
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
umma2_selftest_kernel(const float *A, const float *B, int N, int K, float *D)
{
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = ptx::smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t *smem = smem_raw + (base - raw);
    const int nkb = (K + 63) / 64;
    const int half_n = N / 2;
    const uint32_t b_blk = (uint32_t)half_n * 128u;                       // bytes of this CTA's share of one B block
    const uint32_t a_off = 0, b_off = nkb * A_BLK_BYTES, bar_off = b_off + nkb * b_blk;
    Bars2 *bars = reinterpret_cast<Bars2 *>(smem + ((bar_off + 15u) & ~15u));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = ptx::cluster_ctarank();
    if (threadIdx.x == 0) {
        ptx::mbar_init(&bars->ops_ready, 2);
        ptx::mbar_init(&bars->done, 1);
        ptx::fence_barrier_init();
    }
    ptx::cluster_sync();
    if (warp == 0) ptx::tmem_alloc2(&bars->tmem_base, 256);
    for (int ch = threadIdx.x; ch < nkb * 128 * 8; ch += blockDim.x) {
        const int kb = ch / (128 * 8), r = (ch / 8) % 128, c = ch % 8;
        const int gr = (int)rank * 128 + r;
        uint32_t p[4];
        for (int i = 0; i < 4; ++i) {
            const int k = kb * 64 + c * 8 + 2 * i;
            p[i] = pack_h2(k < K ? A[(size_t)gr * K + k] : 0.f, k + 1 < K ? A[(size_t)gr * K + k + 1] : 0.f);
        }
        *reinterpret_cast<uint4 *>(smem + a_off + kb * A_BLK_BYTES + sw128_off(r, c)) = make_uint4(p[0], p[1], p[2], p[3]);
    }
    for (int ch = threadIdx.x; ch < nkb * half_n * 8; ch += blockDim.x) {
        const int kb = ch / (half_n * 8), r = (ch / 8) % half_n, c = ch % 8;
        const int gr = (int)rank * half_n + r;
        uint32_t p[4];
        for (int i = 0; i < 4; ++i) {
            const int k = kb * 64 + c * 8 + 2 * i;
            p[i] = pack_h2(k < K ? B[(size_t)gr * K + k] : 0.f, k + 1 < K ? B[(size_t)gr * K + k + 1] : 0.f);
        }
        *reinterpret_cast<uint4 *>(smem + b_off + kb * b_blk + sw128_off(r, c)) = make_uint4(p[0], p[1], p[2], p[3]);
    }
    ptx::fence_proxy_async_smem();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = bars->tmem_base;
    if (threadIdx.x == 0) ptx::mbar_arrive_cluster(ptx::map_to_cta(ptx::smem_u32(&bars->ops_ready), 0));
    if (rank == 0 && threadIdx.x == 0) {
        ptx::mbar_wait_cluster(&bars->ops_ready, 0, 80);
        ptx::tc_fence_after();
        const uint32_t idesc = ptx::umma_idesc_f16(256, N);
        for (int kb = 0; kb < nkb; ++kb) {
            const uint64_t da = ptx::umma_desc_sw128(base + a_off + kb * A_BLK_BYTES);
            const uint64_t db = ptx::umma_desc_sw128(base + b_off + kb * b_blk);
            for (int k = 0; k < 4; ++k)
                ptx::umma2_f16(tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb | k) ? 1u : 0u);
        }
        ptx::umma2_commit(&bars->done, 0x3);
    }
    ptx::mbar_wait_cluster(&bars->done, 0, 81);
    ptx::tc_fence_after();
    const int row = warp * 32 + lane;
    for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t r[32];
        ptx::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, r);
        ptx::tmem_ld_wait();
        for (int j = 0; j < 32 && c0 + j < N; ++j) D[(size_t)(rank * 128 + row) * N + c0 + j] = __uint_as_float(r[j]);
    }
    ptx::tc_fence_before();
    ptx::cluster_sync();
    if (warp == 0) ptx::tmem_dealloc2(tmem, 256);
}

// Issue-rate probe (scripts/umma_rate.py): `reps` x 4 MMAs (one 64-wide K block, N = 256) on zeroed operands,
// cycles from first issue to the completion barrier.  pair = 1: cta_group::2 on a cluster of two CTAs (M = 256),
// pair = 0: cta_group::1 in every CTA (M = 128).  All CTAs of the grid run the same loop concurrently.
template <bool pair>
__global__ void __launch_bounds__(128, 1) umma_rate_kernel(int reps, unsigned long long *cycles)
{
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = ptx::smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t *smem = smem_raw + (base - raw);
    Bars2 *bars = reinterpret_cast<Bars2 *>(smem + A_BLK_BYTES + W_BLK_BYTES);
    const int warp = threadIdx.x >> 5;
    const uint32_t rank = pair ? ptx::cluster_ctarank() : 0;
    for (int i = threadIdx.x; i < (A_BLK_BYTES + W_BLK_BYTES) / 16; i += blockDim.x) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) {
        ptx::mbar_init(&bars->done, 1);
        ptx::fence_barrier_init();
    }
    ptx::fence_proxy_async_smem();
    if (pair) ptx::cluster_sync(); else __syncthreads();
    if (warp == 0) { if (pair) ptx::tmem_alloc2(&bars->tmem_base, 512); else ptx::tmem_alloc(&bars->tmem_base, 512); }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = bars->tmem_base;
    if (pair) ptx::cluster_sync();
    if (threadIdx.x == 0 && rank == 0) {
        const uint64_t da = ptx::umma_desc_sw128(base), db = ptx::umma_desc_sw128(base + A_BLK_BYTES);
        const uint32_t idesc = ptx::umma_idesc_f16(pair ? 256 : 128, 256);
        const long long t0 = clock64();
        for (int r = 0; r < reps; ++r)
            for (int k = 0; k < 4; ++k) {
                if (pair) ptx::umma2_f16(tmem + (r & 1) * 256, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, 1u);
                else ptx::umma_f16(tmem + (r & 1) * 256, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, 1u);
            }
        if (pair) ptx::umma2_commit(&bars->done, 0x3); else ptx::umma_commit(&bars->done);
        ptx::mbar_wait_cluster(&bars->done, 0, 82);
        cycles[blockIdx.x] = (unsigned long long)(clock64() - t0);
    } else if (threadIdx.x == 0) {
        ptx::mbar_wait_cluster(&bars->done, 0, 83);
    }
    ptx::tc_fence_before();
    if (pair) ptx::cluster_sync(); else __syncthreads();
    if (warp == 0) { if (pair) ptx::tmem_dealloc2(tmem, 512); else ptx::tmem_dealloc(tmem, 512); }
}

}  // namespace

extern "C" int surs_selftest_umma2(surs_ctx *ctx, const float *A, const float *B, int N, int K, float *D, void *stream)
{
    if (!ctx) return 1;
    SURS_CUDA(ctx, cudaSetDevice(ctx->device));
    if (N < 32 || N > 256 || N % 32 || K < 1 || K > 192) SURS_FAIL(ctx, "surs_selftest_umma2: N in 32..256 step 32, K <= 192");
    const int nkb = (K + 63) / 64;
    const int smem = nkb * (A_BLK_BYTES + (N / 2) * 128) + 64 + 1024;
    SURS_CUDA(ctx, cudaFuncSetAttribute(umma2_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    umma2_selftest_kernel<<<2, 128, smem, (cudaStream_t)stream>>>(A, B, N, K, D);
    SURS_LAUNCH_CHECK(ctx, "umma2_selftest_kernel");
    return 0;
}

/* cycles[b] (leader CTAs / every CTA): clock cycles for reps x 4 MMAs of 128(256) x 256 x 16 */
extern "C" int surs_selftest_umma_rate(surs_ctx *ctx, int pair, int grid, int reps, unsigned long long *cycles, void *stream)
{
    if (!ctx) return 1;
    SURS_CUDA(ctx, cudaSetDevice(ctx->device));
    if (grid < 1 || (pair && (grid & 1)) || reps < 1) SURS_FAIL(ctx, "surs_selftest_umma_rate: bad arguments");
    const int smem = A_BLK_BYTES + W_BLK_BYTES + 64 + 1024;
    SURS_CUDA(ctx, cudaFuncSetAttribute(umma_rate_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    SURS_CUDA(ctx, cudaFuncSetAttribute(umma_rate_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(128);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = pair ? 2 : 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pair ? 1 : 0;
    if (pair) SURS_CUDA(ctx, cudaLaunchKernelEx(&cfg, umma_rate_kernel<true>, reps, cycles));
    else SURS_CUDA(ctx, cudaLaunchKernelEx(&cfg, umma_rate_kernel<false>, reps, cycles));
    SURS_LAUNCH_CHECK(ctx, "umma_rate_kernel");
    return 0;
}
