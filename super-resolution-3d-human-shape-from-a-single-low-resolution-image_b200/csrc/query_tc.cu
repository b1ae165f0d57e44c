// SURS_PREC_FP16: the fused point query on the 5th-generation tensor cores (tcgen05 + TMEM).
//
// One persistent CTA per SM; a tile = 128 points = the 128 TMEM lanes (one point per lane).
// Per tile: projection + mask + depth feature + bilinear gather of both feature maps
// (lib/geometry.py, SuRSNet.py:137-154) write the fp16 A operand F[128 x 336] straight into
// shared memory in the UMMA K-major SWIZZLE_128B layout; then both SurfaceClassifier MLPs
// (SurfaceClassifier.py:45-81) run as a chain of tcgen05.mma (M=128, N=256, K=16; fp16 in,
// fp32 accumulate in TMEM).  Activations never leave the SM: each accumulator is read back
// with tcgen05.ld, gets bias + leaky_relu, is rounded to fp16 and written as the next layer's
// A operand.  The skip concat [y ; f] (SurfaceClassifier.py:63-64) is a second accumulation
// from the resident F tile, never a materialised concat.  The 128->1 layer and the sigmoid
// run in the epilogue of layer 3, whose GEMM carries W4's skip part as a 129th output column.
//
// TMEM holds 512 fp32 columns per lane = two 256-column accumulators.  Layer 1 is 512 wide, so
// it is produced in two halves and layer 0 (1024 wide, in four 256 chunks) is recomputed for
// each half; the first half's activations wait in an L2-resident scratch line (64 KB per CTA)
// and come back by bulk TMA for layer 2.
//
// Warp roles: 0-3 gather + epilogues (warp w owns TMEM lanes 32w..32w+31), 4 = MMA issue
// (one elected lane), 5 = weight stream (cp.async.bulk of pre-swizzled 32 KB blocks laid out in
// HBM in consumption order, 2-stage ring), 6 = scratch reload.  All hand-offs are mbarriers.
#include "common.cuh"
#include "ptx.cuh"

namespace {

constexpr int TILE_M = 128;
constexpr int KBLK = 64;                       // K elements per swizzle-128B block
constexpr int A_BLK_BYTES = TILE_M * 128;      // 16 KB: 128 rows x 64 fp16
constexpr int W_BLK_BYTES = 256 * 128;         // 32 KB: 256 rows x 64 fp16
constexpr int W3_ROWS = 144;                   // layer 3: 128 outputs + W4's skip row + padding to 16
constexpr int W3_BLK_BYTES = W3_ROWS * 128;
constexpr int NF_BLK = 6;                      // F tile: 4 (lr) + 1 (hr) + 1 tail (z, pred_lr)
constexpr int NA_BLK = 4;
constexpr int NSTAGE = 2;
constexpr int BLOCKS_PER_MLP = 8 * 6 + 8 * 4 + 14 + 10;     // 104
constexpr int N256_BLOCKS_PER_MLP = 8 * 6 + 8 * 4 + 14;     // 94
constexpr size_t MLP_BYTES = (size_t)N256_BLOCKS_PER_MLP * W_BLK_BYTES + 10 * (size_t)W3_BLK_BYTES;
constexpr int A_FILLS_PER_MLP = 11;            // 8 x E0, E1(h=1), reload, E2
constexpr int NTHREADS = 7 * 32;

constexpr int SMEM_F = 0;
constexpr int SMEM_A = SMEM_F + NF_BLK * A_BLK_BYTES;
constexpr int SMEM_W = SMEM_A + NA_BLK * A_BLK_BYTES;
constexpr int SMEM_BAR = SMEM_W + NSTAGE * W_BLK_BYTES;
constexpr int SMEM_TOTAL = SMEM_BAR + 256 + 1024;          // + barriers + alignment slack

struct Bars {
    uint64_t full_w[NSTAGE], empty_w[NSTAGE];
    uint64_t a_ready[NA_BLK], a_free[NA_BLK];
    uint64_t acc_full[2], acc_free[2];
    uint64_t f_ready, scr_ready;
    uint32_t tmem_base;
};

struct TcParams {
    const uint8_t *weights;                    // 2 x MLP_BYTES
    const float *bias[2][SURS_NUM_LAYERS];
    const float *w4y[2];                       // [128]: W4[0, 0:128] fp32
    const __half *f_lr, *f_hr;
    int H_lr, W_lr, H_hr, W_hr;
    uint8_t *scratch;                          // 64 KB per CTA
    int64_t ntiles;
};

// byte offset of 16-byte chunk `chunk` of row `row` inside a [rows x 64] fp16 SWIZZLE_128B block
__device__ __forceinline__ uint32_t sw128_off(int row, int chunk)
{
    return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((chunk ^ (row & 7)) << 4));
}

__device__ __forceinline__ uint32_t pack_h2(float a, float b)
{
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&h);
}

__device__ __forceinline__ float leaky(float v) { return fmaxf(v, SURS_LEAKY * v); }

__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint4 v)
{
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

__device__ __forceinline__ void accum_tap(float (&acc)[8], uint4 v, float w)
{
    const __half2 *h = reinterpret_cast<const __half2 *>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float2 f = __half22float2(h[i]);
        acc[2 * i] = fmaf(w, f.x, acc[2 * i]);
        acc[2 * i + 1] = fmaf(w, f.y, acc[2 * i + 1]);
    }
}

// ---- gather: rows 32*warp .. +31 of the F tile -------------------------------------------
__device__ __forceinline__ void gather_rows(const PointIO &io, const TcParams &prm, int64_t tile, int warp, int lane,
                                            uint32_t f_smem, float &zf_out, float &mask_out)
{
    int64_t n = tile * TILE_M + warp * 32 + lane;
    if (n >= io.n) n = io.n - 1;
    float x, y, z;
    pointio_load(io, n, x, y, z);
    const Projected pr = project_point(io, x, y, z);
    zf_out = pr.zf;
    mask_out = pr.mask;
    const Taps tl = make_taps(pr.u, pr.v, prm.H_lr, prm.W_lr);
    const Taps th = make_taps(pr.u, pr.v, prm.H_hr, prm.W_hr);
    // low-res map: 256 channels = 32 lanes x 8 channels, one point per step
#pragma unroll 4
    for (int p = 0; p < 32; ++p) {
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int off = __shfl_sync(0xffffffffu, tl.off[q], p);
            const float w = __shfl_sync(0xffffffffu, tl.w[q], p);
            if (off >= 0) {
                const uint4 v = __ldg(reinterpret_cast<const uint4 *>(prm.f_lr + (size_t)off * SURS_C_LR) + lane);
                accum_tap(acc, v, w);
            }
        }
        const int row = warp * 32 + p;
        const uint4 o = make_uint4(pack_h2(acc[0], acc[1]), pack_h2(acc[2], acc[3]), pack_h2(acc[4], acc[5]), pack_h2(acc[6], acc[7]));
        st_shared_v4(f_smem + (lane >> 3) * A_BLK_BYTES + sw128_off(row, lane & 7), o);
    }
    // high-res map: 64 channels = 8 lanes x 8 channels, four points per step
#pragma unroll 2
    for (int it = 0; it < 8; ++it) {
        const int p = it * 4 + (lane >> 3);
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int off = __shfl_sync(0xffffffffu, th.off[q], p);
            const float w = __shfl_sync(0xffffffffu, th.w[q], p);
            if (off >= 0) {
                const uint4 v = __ldg(reinterpret_cast<const uint4 *>(prm.f_hr + (size_t)off * SURS_C_HR) + (lane & 7));
                accum_tap(acc, v, w);
            }
        }
        const int row = warp * 32 + p;
        const uint4 o = make_uint4(pack_h2(acc[0], acc[1]), pack_h2(acc[2], acc[3]), pack_h2(acc[4], acc[5]), pack_h2(acc[6], acc[7]));
        st_shared_v4(f_smem + 4 * A_BLK_BYTES + sw128_off(row, lane & 7), o);
    }
}

// tail block columns: [z_hi, z_lo, pred_hi, pred_lo, 0...]; hi/lo splits keep ~22 bits of the two scalars
__device__ __forceinline__ void write_tail(uint32_t f_smem, int row, float zf, float pred)
{
    const __half zh = __float2half_rn(zf), ph = __float2half_rn(pred);
    const float zl = zf - __half2float(zh), pl = pred - __half2float(ph);
    const uint4 c0 = make_uint4(pack_h2(__half2float(zh), zl), pack_h2(__half2float(ph), pl), 0u, 0u);
    st_shared_v4(f_smem + 5 * A_BLK_BYTES + sw128_off(row, 0), c0);
    st_shared_v4(f_smem + 5 * A_BLK_BYTES + sw128_off(row, 1), make_uint4(0u, 0u, 0u, 0u));
}

// ---- epilogue of a 256-wide accumulator: +bias, leaky_relu, fp16, K-major swizzled store -------
// dst_smem != 0: into the A operand blocks (hand-off per 64-column block through a_free / a_ready)
// dst_gmem != 0: into the scratch line (same image, reloaded later by bulk TMA)
__device__ __forceinline__ void epilogue_256(uint32_t taddr, const float *__restrict__ bias, int row, int lane,
                                             uint32_t dst_smem, uint8_t *dst_gmem, Bars *bars, uint32_t fill_parity)
{
#pragma unroll 1
    for (int kb = 0; kb < NA_BLK; ++kb) {
        if (dst_smem) ptx::mbar_wait(&bars->a_free[kb], fill_parity ^ 1u, 10 + kb);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            uint32_t r[32];
            ptx::tmem_ld32(taddr + kb * 64 + half * 32, r);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int ch = kb * 64 + half * 32 + j * 8;
                const float4 b0 = __ldg(reinterpret_cast<const float4 *>(bias + ch));
                const float4 b1 = __ldg(reinterpret_cast<const float4 *>(bias + ch + 4));
                const float v0 = leaky(__uint_as_float(r[8 * j + 0]) + b0.x), v1 = leaky(__uint_as_float(r[8 * j + 1]) + b0.y);
                const float v2 = leaky(__uint_as_float(r[8 * j + 2]) + b0.z), v3 = leaky(__uint_as_float(r[8 * j + 3]) + b0.w);
                const float v4 = leaky(__uint_as_float(r[8 * j + 4]) + b1.x), v5 = leaky(__uint_as_float(r[8 * j + 5]) + b1.y);
                const float v6 = leaky(__uint_as_float(r[8 * j + 6]) + b1.z), v7 = leaky(__uint_as_float(r[8 * j + 7]) + b1.w);
                const uint4 o = make_uint4(pack_h2(v0, v1), pack_h2(v2, v3), pack_h2(v4, v5), pack_h2(v6, v7));
                const uint32_t off = kb * A_BLK_BYTES + sw128_off(row, half * 4 + j);
                if (dst_smem) st_shared_v4(dst_smem + off, o);
                else *reinterpret_cast<uint4 *>(dst_gmem + off) = o;
            }
        }
        if (dst_smem) {
            ptx::fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&bars->a_ready[kb]);
        }
    }
}

// K-major MMAs over one 64-wide (or 16-wide tail) K block
__device__ __forceinline__ void mma_block(uint32_t tmem_d, uint32_t a_addr, uint32_t w_addr, int ksteps, uint32_t idesc, bool zero_first)
{
    const uint64_t da = ptx::umma_desc_sw128(a_addr), db = ptx::umma_desc_sw128(w_addr);
#pragma unroll 1
    for (int k = 0; k < ksteps; ++k)
        ptx::umma_f16(tmem_d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (zero_first && k == 0) ? 0u : 1u);
}

__global__ void __launch_bounds__(NTHREADS, 1) query_tc_kernel(PointIO io, TcParams prm)
{
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = ptx::smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t *smem = smem_raw + (base - raw);
    const uint32_t f_smem = base + SMEM_F, a_smem = base + SMEM_A, w_smem = base + SMEM_W;
    Bars *bars = reinterpret_cast<Bars *>(smem + SMEM_BAR);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NSTAGE; ++s) { ptx::mbar_init(&bars->full_w[s], 1); ptx::mbar_init(&bars->empty_w[s], 1); }
        for (int k = 0; k < NA_BLK; ++k) { ptx::mbar_init(&bars->a_ready[k], 4); ptx::mbar_init(&bars->a_free[k], 1); }
        for (int t = 0; t < 2; ++t) { ptx::mbar_init(&bars->acc_full[t], 1); ptx::mbar_init(&bars->acc_free[t], 4); }
        ptx::mbar_init(&bars->f_ready, 4);
        ptx::mbar_init(&bars->scr_ready, 4);
        ptx::fence_barrier_init();
    }
    if (warp == 4) ptx::tmem_alloc(&bars->tmem_base, 512);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = bars->tmem_base;
    const uint32_t T0 = tmem, T1 = tmem + 256;
    uint8_t *scratch = prm.scratch + (size_t)blockIdx.x * (NA_BLK * A_BLK_BYTES);

    if (warp < 4) {
        // =============================== gather + epilogues ===============================
        const int row = warp * 32 + lane;
        const uint32_t lane_t0 = T0 + ((uint32_t)(warp * 32) << 16), lane_t1 = T1 + ((uint32_t)(warp * 32) << 16);
        uint32_t afill = 0, acc0 = 0, acc1 = 0;
        for (int64_t tile = blockIdx.x; tile < prm.ntiles; tile += gridDim.x) {
            float zf, mask;
            gather_rows(io, prm, tile, warp, lane, f_smem, zf, mask);
            write_tail(f_smem, row, zf, 0.0f);
            ptx::fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&bars->f_ready);
            float pred_lr = 0.0f;
#pragma unroll 1
            for (int m = 0; m < 2; ++m) {
#pragma unroll 1
                for (int h = 0; h < 2; ++h) {
#pragma unroll 1
                    for (int c = 0; c < 4; ++c) {           // E0: layer-0 chunk -> A
                        ptx::mbar_wait(&bars->acc_full[0], acc0 & 1u, 20);
                        ptx::tc_fence_after();
                        epilogue_256(lane_t0, prm.bias[m][0] + c * 256, row, lane, a_smem, nullptr, bars, afill & 1u);
                        ++afill;
                        ptx::tc_fence_before();
                        __syncwarp();
                        if (lane == 0) ptx::mbar_arrive(&bars->acc_free[0]);
                        ++acc0;
                    }
                    // E1: layer-1 half -> scratch (h == 0) or A (h == 1)
                    ptx::mbar_wait(&bars->acc_full[1], acc1 & 1u, 21);
                    ptx::tc_fence_after();
                    if (h == 0) {
                        epilogue_256(lane_t1, prm.bias[m][1], row, lane, 0u, scratch, bars, 0u);
                        ptx::fence_proxy_async_all();
                    } else {
                        epilogue_256(lane_t1, prm.bias[m][1] + 256, row, lane, a_smem, nullptr, bars, afill & 1u);
                        ++afill;
                        // Only now may the reload warp start waiting on a_free: mbarrier parity waits
                        // are only meaningful one phase ahead, and its fill is the next one.
                        __syncwarp();
                        if (lane == 0) ptx::mbar_arrive(&bars->scr_ready);
                    }
                    ptx::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive(&bars->acc_free[1]);
                    ++acc1;
                }
                ++afill;                                    // the reload's fill of A (warp 6)
                // E2: layer 2 -> A
                ptx::mbar_wait(&bars->acc_full[0], acc0 & 1u, 22);
                ptx::tc_fence_after();
                epilogue_256(lane_t0, prm.bias[m][2], row, lane, a_smem, nullptr, bars, afill & 1u);
                ++afill;
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(&bars->acc_free[0]);
                ++acc0;
                // E3: layer 3 + layer 4 + sigmoid
                ptx::mbar_wait(&bars->acc_full[1], acc1 & 1u, 23);
                ptx::tc_fence_after();
                float logit = __ldg(prm.bias[m][4]);
#pragma unroll 1
                for (int q = 0; q < 4; ++q) {
                    uint32_t r[32];
                    ptx::tmem_ld32(lane_t1 + q * 32, r);
                    ptx::tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float4 b = __ldg(reinterpret_cast<const float4 *>(prm.bias[m][3] + q * 32 + j));
                        const float4 w = __ldg(reinterpret_cast<const float4 *>(prm.w4y[m] + q * 32 + j));
                        logit = fmaf(w.x, leaky(__uint_as_float(r[j]) + b.x), logit);
                        logit = fmaf(w.y, leaky(__uint_as_float(r[j + 1]) + b.y), logit);
                        logit = fmaf(w.z, leaky(__uint_as_float(r[j + 2]) + b.z), logit);
                        logit = fmaf(w.w, leaky(__uint_as_float(r[j + 3]) + b.w), logit);
                    }
                }
                {
                    uint32_t r[32];
                    ptx::tmem_ld32(lane_t1 + 128, r);       // column 128 = W4's skip part . f
                    ptx::tmem_ld_wait();
                    logit += __uint_as_float(r[0]);
                }
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(&bars->acc_free[1]);
                ++acc1;
                const float pred = mask * (1.0f / (1.0f + expf(-logit)));
                if (m == 0) {
                    pred_lr = pred;
                    write_tail(f_smem, row, zf, pred);      // SuRSNet.py:180: the MASKED pred_lr is channel 321
                    ptx::fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive(&bars->f_ready);
                } else {
                    const int64_t n = tile * TILE_M + row;
                    if (n < io.n) pointio_store(io, n, pred, pred_lr);
                }
            }
        }
    } else if (warp == 4) {
        // =============================== MMA issue ========================================
        if (lane == 0) {
            constexpr uint32_t IDESC256 = ptx::umma_idesc_f16(128, 256);
            constexpr uint32_t IDESC144 = ptx::umma_idesc_f16(128, W3_ROWS);
            uint32_t wblk = 0, afill = 0, acc0 = 0, acc1 = 0, fcnt = 0;
            auto wait_w = [&]() -> uint32_t {
                const uint32_t s = wblk % NSTAGE;
                ptx::mbar_wait(&bars->full_w[s], (wblk / NSTAGE) & 1u, 30);
                ptx::tc_fence_after();
                return w_smem + s * W_BLK_BYTES;
            };
            auto release_w = [&]() {
                ptx::umma_commit(&bars->empty_w[wblk % NSTAGE]);
                ++wblk;
            };
            // accumulate F (6 blocks: 5 x K=64, tail K=16) into tmem_d
            auto mma_F = [&](uint32_t tmem_d, uint32_t idesc, bool zero_first) {
                for (int kb = 0; kb < NF_BLK; ++kb) {
                    const uint32_t w = wait_w();
                    mma_block(tmem_d, f_smem + kb * A_BLK_BYTES, w, kb == NF_BLK - 1 ? 1 : 4, idesc, zero_first && kb == 0);
                    release_w();
                }
            };
            // accumulate the A operand (4 blocks) into tmem_d; hands each block back through a_free
            auto mma_A = [&](uint32_t tmem_d, uint32_t idesc, bool zero_first) {
                for (int kb = 0; kb < NA_BLK; ++kb) {
                    ptx::mbar_wait(&bars->a_ready[kb], afill & 1u, 31);
                    const uint32_t w = wait_w();
                    mma_block(tmem_d, a_smem + kb * A_BLK_BYTES, w, 4, idesc, zero_first && kb == 0);
                    release_w();
                    ptx::umma_commit(&bars->a_free[kb]);
                }
                ++afill;
            };
            for (int64_t tile = blockIdx.x; tile < prm.ntiles; tile += gridDim.x) {
                for (int m = 0; m < 2; ++m) {
                    ptx::mbar_wait(&bars->f_ready, fcnt & 1u, 32);     // gather done / pred_lr written
                    ++fcnt;
                    ptx::tc_fence_after();
                    for (int h = 0; h < 2; ++h) {
                        for (int c = 0; c < 4; ++c) {
                            ptx::mbar_wait(&bars->acc_free[0], (acc0 & 1u) ^ 1u, 33);
                            ptx::tc_fence_after();
                            mma_F(T0, IDESC256, true);                  // layer 0, chunk c
                            ptx::umma_commit(&bars->acc_full[0]);
                            ++acc0;
                            if (c == 0) {
                                ptx::mbar_wait(&bars->acc_free[1], (acc1 & 1u) ^ 1u, 34);
                                ptx::tc_fence_after();
                            }
                            mma_A(T1, IDESC256, c == 0);                // layer 1, half h, K chunk c
                        }
                        ptx::umma_commit(&bars->acc_full[1]);
                        ++acc1;
                    }
                    // layer 2 = W2[:, 256:512] y1[1] + W2[:, 0:256] y1[0] + W2[:, 512:] f
                    ptx::mbar_wait(&bars->acc_free[0], (acc0 & 1u) ^ 1u, 35);
                    ptx::tc_fence_after();
                    mma_A(T0, IDESC256, true);
                    mma_A(T0, IDESC256, false);
                    mma_F(T0, IDESC256, false);
                    ptx::umma_commit(&bars->acc_full[0]);
                    ++acc0;
                    // layer 3 (+ W4's skip row) = W3[:, 0:256] y2 + W3[:, 256:] f
                    ptx::mbar_wait(&bars->acc_free[1], (acc1 & 1u) ^ 1u, 36);
                    ptx::tc_fence_after();
                    mma_A(T1, IDESC144, true);
                    mma_F(T1, IDESC144, false);
                    ptx::umma_commit(&bars->acc_full[1]);
                    ++acc1;
                }
            }
        }
    } else if (warp == 5) {
        // =============================== weight stream ====================================
        if (lane == 0) {
            uint32_t wblk = 0;
            for (int64_t tile = blockIdx.x; tile < prm.ntiles; tile += gridDim.x) {
                const uint8_t *src = prm.weights;
                for (int b = 0; b < 2 * BLOCKS_PER_MLP; ++b) {
                    const uint32_t bytes = (b % BLOCKS_PER_MLP) < N256_BLOCKS_PER_MLP ? W_BLK_BYTES : W3_BLK_BYTES;
                    const uint32_t s = wblk % NSTAGE;
                    ptx::mbar_wait(&bars->empty_w[s], ((wblk / NSTAGE) & 1u) ^ 1u, 40);
                    ptx::mbar_arrive_expect_tx(&bars->full_w[s], bytes);
                    ptx::tma_load_1d(smem + SMEM_W + s * W_BLK_BYTES, src, bytes, &bars->full_w[s]);
                    src += bytes;
                    ++wblk;
                }
            }
        }
    } else {
        // =============================== scratch reload ===================================
        if (lane == 0) {
            uint32_t scr = 0, tiles_done = 0;
            for (int64_t tile = blockIdx.x; tile < prm.ntiles; tile += gridDim.x, ++tiles_done) {
                for (int m = 0; m < 2; ++m) {
                    // this is fill number 9 (0-based) of the MLP's 11 fills of A
                    const uint32_t fill = (tiles_done * 2 + m) * A_FILLS_PER_MLP + 9;
                    ptx::mbar_wait(&bars->scr_ready, scr & 1u, 50);
                    ++scr;
                    for (int kb = 0; kb < NA_BLK; ++kb) {
                        ptx::mbar_wait(&bars->a_free[kb], (fill & 1u) ^ 1u, 51);
                        ptx::mbar_arrive_expect_tx(&bars->a_ready[kb], A_BLK_BYTES);
                        ptx::mbar_arrive(&bars->a_ready[kb]);
                        ptx::mbar_arrive(&bars->a_ready[kb]);
                        ptx::mbar_arrive(&bars->a_ready[kb]);
                        ptx::tma_load_1d(smem + SMEM_A + kb * A_BLK_BYTES, scratch + kb * A_BLK_BYTES, A_BLK_BYTES, &bars->a_ready[kb]);
                    }
                }
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 4) ptx::tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------
// weight packing: fp32 [Cout, Cin] -> the block stream consumed above
// ------------------------------------------------------------------------------------------
struct PackDesc {
    const float *w;        // source layer, row-major [cout][cin]
    const float *w_extra;  // layer 4 (row 128 of the layer-3 skip blocks) or NULL
    int cin, cin_extra;
    int row0, nrows;       // rows taken from w; the block has ntotal rows, the rest is zero
    int ntotal;
    int fblock;            // -1: plain columns k0 .. k0+63; else F-order block index 0..5
    int k0, k0_extra;      // first column (plain) / start of the skip part (F-order)
    int c0;                // 321 / 322: width of the skip input
    uint32_t out_off;
};

// column of the skip input that F-tile position (fblock, kk) holds, or -1 for padding
__device__ __forceinline__ int fmap(int fblock, int kk, int c0)
{
    if (fblock < 4) return fblock * 64 + kk;
    if (fblock == 4) return 256 + kk;
    if (kk < 2) return 320;                      // z_hi, z_lo
    if (kk < 4) return c0 > 321 ? 321 : -1;      // pred_hi, pred_lo (HR MLP only)
    return -1;
}

__global__ void pack_weights_kernel(const PackDesc *descs, uint8_t *out)
{
    const PackDesc d = descs[blockIdx.x];
    for (int ch = threadIdx.x; ch < d.ntotal * 8; ch += blockDim.x) {
        const int r = ch >> 3, c = ch & 7;
        uint32_t packed[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float v[2];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int kk = c * 8 + 2 * i + j;
                float x = 0.0f;
                if (d.fblock < 0) {
                    if (r < d.nrows) x = d.w[(size_t)(d.row0 + r) * d.cin + d.k0 + kk];
                } else {
                    const int col = fmap(d.fblock, kk, d.c0);
                    if (col >= 0) {
                        if (r < d.nrows) x = d.w[(size_t)(d.row0 + r) * d.cin + d.k0 + col];
                        else if (r == 128 && d.w_extra) x = d.w_extra[d.k0_extra + col];
                    }
                }
                v[j] = x;
            }
            packed[i] = pack_h2(v[0], v[1]);
        }
        *reinterpret_cast<uint4 *>(out + d.out_off + sw128_off(r, c)) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
    }
}

__global__ void copy_w4y_kernel(const float *w4, float *dst)
{
    if (threadIdx.x < 128) dst[threadIdx.x] = w4[threadIdx.x];
}

// ------------------------------------------------------------------------------------------
// self test: D[128,N] = A[128,K] . B[N,K]^T through the same descriptors / layouts
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1) umma_selftest_kernel(const float *A, const float *B, int N, int K, int tail16, float *D)
{
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = ptx::smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t *smem = smem_raw + (base - raw);
    const int nkb = (K + 63) / 64;
    const uint32_t a_off = 0, b_off = nkb * A_BLK_BYTES, bar_off = b_off + nkb * W_BLK_BYTES;
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + bar_off);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + bar_off + 8);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { ptx::mbar_init(bar, 1); ptx::fence_barrier_init(); }
    if (warp == 0) ptx::tmem_alloc(tmem_slot, 256);
    // operands -> swizzled K-major fp16 images (zero padded to whole blocks)
    for (int ch = threadIdx.x; ch < nkb * 128 * 8; ch += blockDim.x) {
        const int kb = ch / (128 * 8), r = (ch / 8) % 128, c = ch % 8;
        uint32_t p[4];
        for (int i = 0; i < 4; ++i) {
            const int k = kb * 64 + c * 8 + 2 * i;
            p[i] = pack_h2(k < K ? A[(size_t)r * K + k] : 0.f, k + 1 < K ? A[(size_t)r * K + k + 1] : 0.f);
        }
        *reinterpret_cast<uint4 *>(smem + a_off + kb * A_BLK_BYTES + sw128_off(r, c)) = make_uint4(p[0], p[1], p[2], p[3]);
    }
    for (int ch = threadIdx.x; ch < nkb * 256 * 8; ch += blockDim.x) {
        const int kb = ch / (256 * 8), r = (ch / 8) % 256, c = ch % 8;
        uint32_t p[4];
        for (int i = 0; i < 4; ++i) {
            const int k = kb * 64 + c * 8 + 2 * i;
            p[i] = pack_h2((r < N && k < K) ? B[(size_t)r * K + k] : 0.f, (r < N && k + 1 < K) ? B[(size_t)r * K + k + 1] : 0.f);
        }
        *reinterpret_cast<uint4 *>(smem + b_off + kb * W_BLK_BYTES + sw128_off(r, c)) = make_uint4(p[0], p[1], p[2], p[3]);
    }
    ptx::fence_proxy_async_smem();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    if (threadIdx.x == 0) {
        const uint32_t idesc = ptx::umma_idesc_f16(128, N);
        for (int kb = 0; kb < nkb; ++kb) {
            const int ksteps = (kb == nkb - 1 && tail16) ? 1 : 4;
            mma_block(tmem, base + a_off + kb * A_BLK_BYTES, base + b_off + kb * W_BLK_BYTES, ksteps, idesc, kb == 0);
        }
        ptx::umma_commit(bar);
    }
    ptx::mbar_wait(bar, 0, 60);
    ptx::tc_fence_after();
    const int row = warp * 32 + lane;
    for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t r[32];
        ptx::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, r);
        ptx::tmem_ld_wait();
        for (int j = 0; j < 32 && c0 + j < N; ++j) D[(size_t)row * N + c0 + j] = __uint_as_float(r[j]);
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 0) ptx::tmem_dealloc(tmem, 256);
}

}  // namespace

int surs_tc_pack_weights(surs_ctx *ctx, const float *const w[2][SURS_NUM_LAYERS], cudaStream_t st)
{
    const size_t total = 2 * MLP_BYTES + 2 * 128 * sizeof(float);
    if (!ctx->tc_weights) {
        SURS_CUDA(ctx, cudaMalloc(&ctx->tc_weights, total));
        ctx->tc_weights_bytes = total;
    }
    PackDesc host[2 * BLOCKS_PER_MLP];
    int n = 0;
    for (int m = 0; m < 2; ++m) {
        const int c0 = m == 0 ? SURS_C0_LR : SURS_C0_HR;
        uint32_t off = (uint32_t)(m * MLP_BYTES);
        auto add = [&](int layer, int row0, int nrows, int ntotal, int fblock, int k0, bool extra) {
            PackDesc d;
            memset(&d, 0, sizeof(d));
            d.w = w[m][layer];
            d.cin = ctx->cin[m][layer];
            d.w_extra = extra ? w[m][4] : nullptr;
            d.cin_extra = ctx->cin[m][4];
            d.k0_extra = 128;
            d.row0 = row0; d.nrows = nrows; d.ntotal = ntotal; d.fblock = fblock; d.k0 = k0; d.c0 = c0;
            d.out_off = off;
            off += (uint32_t)ntotal * 128u;
            host[n++] = d;
        };
        for (int h = 0; h < 2; ++h)
            for (int c = 0; c < 4; ++c) {
                for (int kb = 0; kb < NF_BLK; ++kb) add(0, c * 256, 256, 256, kb, 0, false);          // layer 0 chunk c
                for (int kb = 0; kb < NA_BLK; ++kb) add(1, h * 256, 256, 256, -1, c * 256 + kb * 64, false);   // layer 1
            }
        for (int kb = 0; kb < NA_BLK; ++kb) add(2, 0, 256, 256, -1, 256 + kb * 64, false);           // y1[1]
        for (int kb = 0; kb < NA_BLK; ++kb) add(2, 0, 256, 256, -1, kb * 64, false);                 // y1[0]
        for (int kb = 0; kb < NF_BLK; ++kb) add(2, 0, 256, 256, kb, 512, false);                     // skip
        for (int kb = 0; kb < NA_BLK; ++kb) add(3, 0, 128, W3_ROWS, -1, kb * 64, false);             // y2
        for (int kb = 0; kb < NF_BLK; ++kb) add(3, 0, 128, W3_ROWS, kb, 256, true);                  // skip + W4 row
        if (off != (uint32_t)((m + 1) * MLP_BYTES)) SURS_FAIL(ctx, "internal: weight stream size mismatch");
    }
    PackDesc *dev = nullptr;
    SURS_CUDA(ctx, cudaMalloc(&dev, sizeof(host)));
    SURS_CUDA(ctx, cudaMemcpyAsync(dev, host, sizeof(host), cudaMemcpyHostToDevice, st));
    pack_weights_kernel<<<n, 256, 0, st>>>(dev, (uint8_t *)ctx->tc_weights);
    SURS_LAUNCH_CHECK(ctx, "pack_weights_kernel");
    float *w4y = reinterpret_cast<float *>((uint8_t *)ctx->tc_weights + 2 * MLP_BYTES);
    for (int m = 0; m < 2; ++m) {
        copy_w4y_kernel<<<1, 128, 0, st>>>(w[m][4], w4y + 128 * m);
        SURS_LAUNCH_CHECK(ctx, "copy_w4y_kernel");
    }
    SURS_CUDA(ctx, cudaStreamSynchronize(st));
    SURS_CUDA(ctx, cudaFree(dev));
    return 0;
}

int surs_launch_query_tc(surs_ctx *ctx, const PointIO &io, cudaStream_t st)
{
    if (io.n <= 0) return 0;
    TcParams prm;
    prm.weights = (const uint8_t *)ctx->tc_weights;
    const float *w4y = reinterpret_cast<const float *>((const uint8_t *)ctx->tc_weights + 2 * MLP_BYTES);
    for (int m = 0; m < 2; ++m) {
        for (int l = 0; l < SURS_NUM_LAYERS; ++l) prm.bias[m][l] = ctx->b32[m][l];
        prm.w4y[m] = w4y + 128 * m;
    }
    prm.f_lr = ctx->f_lr16; prm.f_hr = ctx->f_hr16;
    prm.H_lr = ctx->H_lr; prm.W_lr = ctx->W_lr; prm.H_hr = ctx->H_hr; prm.W_hr = ctx->W_hr;
    prm.ntiles = (io.n + TILE_M - 1) / TILE_M;
    const int grid = (int)(prm.ntiles < ctx->sm_count ? prm.ntiles : ctx->sm_count);
    if (surs_ensure(ctx, (void **)&ctx->tc_scratch, &ctx->tc_scratch_cap, (size_t)ctx->sm_count * NA_BLK * A_BLK_BYTES)) return 1;
    prm.scratch = (uint8_t *)ctx->tc_scratch;
    SURS_CUDA(ctx, cudaFuncSetAttribute(query_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL));
    query_tc_kernel<<<grid, NTHREADS, SMEM_TOTAL, st>>>(io, prm);
    SURS_LAUNCH_CHECK(ctx, "query_tc_kernel");
    return 0;
}

extern "C" int surs_selftest_umma(surs_ctx *ctx, const float *A, const float *B, int N, int K, int tail16, float *D, void *stream)
{
    if (!ctx) return 1;
    SURS_CUDA(ctx, cudaSetDevice(ctx->device));
    if (N < 16 || N > 256 || N % 16 || K < 1 || K > 192) SURS_FAIL(ctx, "surs_selftest_umma: N in 16..256 step 16, K <= 192");
    const int nkb = (K + 63) / 64;
    const int smem = nkb * (A_BLK_BYTES + W_BLK_BYTES) + 64 + 1024;
    SURS_CUDA(ctx, cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    umma_selftest_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(A, B, N, K, tail16, D);
    SURS_LAUNCH_CHECK(ctx, "umma_selftest_kernel");
    return 0;
}
