// Unit test of the CTA-pair MMA path (tcgen05.mma.cta_group::2): D[256,N] = A[256,K] . B[N,K]^T with
// one cluster of two CTAs.  CTA r holds rows 128 r .. of A, rows (N/2) r .. of B (its half of the B
// operand) and receives rows 128 r .. of D in its own TMEM; the leader CTA issues the MMAs for both.
#include "tc_common.cuh"

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
