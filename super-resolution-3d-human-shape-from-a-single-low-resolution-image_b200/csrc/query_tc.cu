// SURS_PREC_FP16: the fused point query on the 5th-generation tensor cores (tcgen05 + TMEM).
//
// One persistent CTA per SM; a tile = 128 points = the 128 TMEM lanes (one point per lane).
// Per tile: projection + mask + depth feature + bilinear gather of both feature maps
// (lib/geometry.py, SuRSNet.py:137-154) write the fp16 A operand F[128 x 336] straight into
// shared memory in the UMMA K-major SWIZZLE_128B layout; then both SurfaceClassifier MLPs
// (SurfaceClassifier.py:45-81) run as a chain of tcgen05.mma (M=128, N=256, K=16; fp16 in,
// fp32 accumulate in TMEM).  Activations never leave the SM: each accumulator is read back
// with tcgen05.ld, gets leaky_relu, is rounded to fp16 and written as the next layer's A operand.
// The skip concat [y ; f] (SurfaceClassifier.py:63-64) is a second accumulation from the
// resident F tile, never a materialised concat.  Biases ride on the tensor core too: F carries
// two constant-one columns and the weight stream holds each bias as an fp16 hi+lo pair against
// them (the 225 KB of shared memory leave no L1 for bias loads).  The 128->1 layer and the
// sigmoid run in the epilogue of layer 3, whose GEMM carries W4's skip part (+ b4) as a 129th
// output column.
//
// TMEM holds 512 fp32 columns per lane = two 256-column accumulators.  Layer 1 is 512 wide, so
// it is produced in two halves and layer 0 (1024 wide, in four 256 chunks) is recomputed for
// each half; the first half's activations wait in an L2-resident scratch line (64 KB per CTA)
// and come back by bulk TMA for layer 2.
//
// Warp roles: 0-7 gather + epilogues (warp w owns TMEM lanes 32(w%4).. and column half w/4),
// 8 = MMA issue (one elected lane), 9 = weight stream (cp.async.bulk of pre-swizzled 32 KB
// blocks laid out in HBM in consumption order, 3-stage ring), 10 = scratch reload.
// All hand-offs are mbarriers; waits are bounded (a protocol bug traps instead of hanging).
#include "tc_common.cuh"

#include <stdlib.h>

namespace {

using namespace tc;

constexpr int NF_BLK = 6;                      // F tile: 4 (lr) + 1 (hr) + 1 tail (z, pred_lr, ones)
constexpr int NA_SLOT = 2;                     // ring of A-operand K blocks (activations)
constexpr int NSTAGE = 3;                      // weight ring
constexpr int N256_BLOCKS_PER_MLP = 8 * 6 + 8 * 4 + 2 + 14;            // 96 blocks of 256 rows
constexpr int BLOCKS_PER_MLP = N256_BLOCKS_PER_MLP + 10;               // + 10 blocks of 144 rows
constexpr size_t MLP_BYTES = (size_t)N256_BLOCKS_PER_MLP * W_BLK_BYTES + 10 * (size_t)W3_BLK_BYTES;
constexpr int A_FILLS_PER_MLP = 11;            // 8 x E0, E1(h=1), reload, E2 (4 K blocks each)
constexpr int NEPI = 8;                        // gather / epilogue warps
constexpr int NTHREADS = (NEPI + 3) * 32;

constexpr int SMEM_F = 0;
constexpr int SMEM_A = SMEM_F + NF_BLK * A_BLK_BYTES;
constexpr int SMEM_W = SMEM_A + NA_SLOT * A_BLK_BYTES;
constexpr int SMEM_BAR = SMEM_W + NSTAGE * W_BLK_BYTES;
constexpr int SMEM_TOTAL = SMEM_BAR + 256 + 1024;          // + barriers + alignment slack

struct Bars {
    uint64_t full_w[NSTAGE], empty_w[NSTAGE];
    uint64_t a_ready[NA_SLOT], a_free[NA_SLOT];
    uint64_t acc_full[2], acc_free[2];
    uint64_t f_ready, scr_ready;
    uint32_t tmem_base;
};

struct TcParams {
    const uint8_t *weights;                    // 2 x MLP_BYTES
    FeatMaps fm;
    uint8_t *scratch;                          // 64 KB per CTA
    int64_t ntiles;
    float w4y[2][128];                         // W4[0, 0:128] of both MLPs (fp32, used in the last epilogue)
};

__device__ __forceinline__ Projected project_row(const PointIO &io, int64_t tile, int row)
{
    int64_t n = tile * TILE_M + row;
    if (n >= io.n) n = io.n - 1;
    float x, y, z;
    pointio_load(io, n, x, y, z);
    return project_point(io, x, y, z);
}

// tail block columns: [z_hi, z_lo, pred_hi, pred_lo, 1, 1, 0, 0 | 0 x 8]; the hi/lo splits keep
// ~22 bits of the two scalars, the ones multiply the (hi, lo) bias columns of the weight stream
__device__ __forceinline__ void write_tail(uint32_t f_smem, int row, float zf, float pred)
{
    const __half zh = __float2half_rn(zf), ph = __float2half_rn(pred);
    const float zl = zf - __half2float(zh), pl = pred - __half2float(ph);
    const uint4 c0 = make_uint4(pack_h2(__half2float(zh), zl), pack_h2(__half2float(ph), pl), pack_h2(1.0f, 1.0f), 0u);
    st_shared_v4(f_smem + 5 * A_BLK_BYTES + sw128_off(row, 0), c0);
    st_shared_v4(f_smem + 5 * A_BLK_BYTES + sw128_off(row, 1), make_uint4(0u, 0u, 0u, 0u));
}

// ---- epilogue of a 256-wide accumulator: leaky_relu, fp16, K-major swizzled store -------------
// Warp (quarter q, half hsel) converts columns [64 kb + 32 hsel, +32) of rows 32 q.. for kb = 0..3.
// to_smem: into the A ring (K block g0 + kb -> slot (g0 + kb) % NA_SLOT, hand-off a_free / a_ready);
// else into the global scratch image (reloaded later by bulk TMA).
__device__ __forceinline__ void epilogue_256(uint32_t taddr, int row, int hsel, int lane, bool to_smem, uint32_t a_smem,
                                             uint8_t *dst_gmem, Bars *bars, uint32_t g0, unsigned long long *prof)
{
    uint32_t r[2][32];
    ptx::tmem_ld32(taddr + hsel * 32, r[0]);
#pragma unroll
    for (int kb = 0; kb < 4; ++kb) {
        ptx::tmem_ld_wait();
        if (kb < 3) ptx::tmem_ld32(taddr + (kb + 1) * 64 + hsel * 32, r[(kb + 1) & 1]);
        const uint32_t g = g0 + kb, slot = g % NA_SLOT;
        if (to_smem) ptx::mbar_wait(&bars->a_free[slot], ((g / NA_SLOT) & 1u) ^ 1u, 10, prof);
        const uint32_t *v = r[kb & 1];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint4 o = make_uint4(leaky_h2(__uint_as_float(v[8 * j + 0]), __uint_as_float(v[8 * j + 1])),
                                       leaky_h2(__uint_as_float(v[8 * j + 2]), __uint_as_float(v[8 * j + 3])),
                                       leaky_h2(__uint_as_float(v[8 * j + 4]), __uint_as_float(v[8 * j + 5])),
                                       leaky_h2(__uint_as_float(v[8 * j + 6]), __uint_as_float(v[8 * j + 7])));
            const uint32_t off = sw128_off(row, hsel * 4 + j);
            if (to_smem) st_shared_v4(a_smem + slot * A_BLK_BYTES + off, o);
            else *reinterpret_cast<uint4 *>(dst_gmem + kb * A_BLK_BYTES + off) = o;
        }
        if (to_smem) {
            ptx::fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&bars->a_ready[slot]);
        }
    }
}

__device__ unsigned long long g_tc_prof[64];

template <bool PROF>
__global__ void __launch_bounds__(NTHREADS, 1) query_tc_kernel(const __grid_constant__ PointIO io, const __grid_constant__ TcParams prm)
{
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = ptx::smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t *smem = smem_raw + (base - raw);
    const uint32_t f_smem = base + SMEM_F, a_smem = base + SMEM_A, w_smem = base + SMEM_W;
    Bars *bars = reinterpret_cast<Bars *>(smem + SMEM_BAR);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // profiling build: one designated thread per role accumulates its wait cycles into g_tc_prof[tag]
    unsigned long long *prof = nullptr;
    if (PROF && lane == 0 && (warp == 0 || warp >= NEPI)) prof = g_tc_prof;
    const long long t_kernel0 = PROF ? clock64() : 0;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NSTAGE; ++s) { ptx::mbar_init(&bars->full_w[s], 1); ptx::mbar_init(&bars->empty_w[s], 1); }
        for (int k = 0; k < NA_SLOT; ++k) { ptx::mbar_init(&bars->a_ready[k], NEPI); ptx::mbar_init(&bars->a_free[k], 1); }
        for (int t = 0; t < 2; ++t) { ptx::mbar_init(&bars->acc_full[t], 1); ptx::mbar_init(&bars->acc_free[t], NEPI); }
        ptx::mbar_init(&bars->f_ready, NEPI);
        ptx::mbar_init(&bars->scr_ready, NEPI);
        ptx::fence_barrier_init();
    }
    if (warp == NEPI) ptx::tmem_alloc(&bars->tmem_base, 512);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = bars->tmem_base;
    const uint32_t T0 = tmem, T1 = tmem + 256;
    uint8_t *scratch = prm.scratch + (size_t)blockIdx.x * (4 * A_BLK_BYTES);

    if (warp < NEPI) {
        // =============================== gather + epilogues ===============================
        const int quarter = warp & 3, hsel = warp >> 2;
        const int row = quarter * 32 + lane;
        const uint32_t lane_t0 = T0 + ((uint32_t)(quarter * 32) << 16), lane_t1 = T1 + ((uint32_t)(quarter * 32) << 16);
        uint32_t afill = 0, acc0 = 0, acc1 = 0;
        auto release_acc = [&](int t) {
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&bars->acc_free[t]);
        };
        for (int64_t tile = blockIdx.x; tile < prm.ntiles; tile += gridDim.x) {
            const long long t_g0 = PROF ? clock64() : 0;
            {   // warp w fills rows 32 (w % 4) + 16 (w / 4) .. + 15; lanes 16-31 mirror lanes 0-15
                const int row0 = (warp & 3) * 32 + (warp >> 2) * 16;
                gather_rows<16>(prm.fm, project_row(io, tile, row0 + (lane & 15)), row0, lane, f_smem);
            }
            float zf = 0.f, mask = 0.f;
            if (hsel == 0) {                       // warps 0-3 own the per-row scalars and the tail block
                const Projected own = project_row(io, tile, row);
                zf = own.zf;
                mask = own.mask;
                write_tail(f_smem, row, zf, 0.0f);
            }
            ptx::fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&bars->f_ready);
            if (prof) { atomicAdd(prof + 1, (unsigned long long)(clock64() - t_g0)); atomicAdd(prof + 2, 1ull); }
            float pred_lr = 0.0f;
#pragma unroll 1
            for (int m = 0; m < 2; ++m) {
#pragma unroll 1
                for (int h = 0; h < 2; ++h) {
#pragma unroll 1
                    for (int c = 0; c < 4; ++c) {           // E0: layer-0 chunk -> A
                        ptx::mbar_wait(&bars->acc_full[0], acc0 & 1u, 20, prof);
                        ptx::tc_fence_after();
                        epilogue_256(lane_t0, row, hsel, lane, true, a_smem, nullptr, bars, afill * 4, prof);
                        ++afill;
                        release_acc(0);
                        ++acc0;
                    }
                    // E1: layer-1 half -> scratch (h == 0) or A (h == 1)
                    ptx::mbar_wait(&bars->acc_full[1], acc1 & 1u, 21, prof);
                    ptx::tc_fence_after();
                    if (h == 0) {
                        epilogue_256(lane_t1, row, hsel, lane, false, 0u, scratch, bars, 0u, prof);
                        ptx::fence_proxy_async_all();
                    } else {
                        epilogue_256(lane_t1, row, hsel, lane, true, a_smem, nullptr, bars, afill * 4, prof);
                        ++afill;
                        // Only now may the reload warp start waiting on a_free: mbarrier parity waits
                        // are only meaningful one phase ahead, and its fill is the next one.
                        __syncwarp();
                        if (lane == 0) ptx::mbar_arrive(&bars->scr_ready);
                    }
                    release_acc(1);
                    ++acc1;
                }
                ++afill;                                    // the reload's fill of A (warp NEPI + 2)
                // E2: layer 2 -> A
                ptx::mbar_wait(&bars->acc_full[0], acc0 & 1u, 22, prof);
                ptx::tc_fence_after();
                epilogue_256(lane_t0, row, hsel, lane, true, a_smem, nullptr, bars, afill * 4, prof);
                ++afill;
                release_acc(0);
                ++acc0;
                // E3: layer 3 + layer 4 + sigmoid (warps 0-3; the others only hand the accumulator back)
                ptx::mbar_wait(&bars->acc_full[1], acc1 & 1u, 23, prof);
                ptx::tc_fence_after();
                float logit = 0.0f;
                if (hsel == 0) {
#pragma unroll 1
                    for (int q = 0; q < 4; ++q) {
                        uint32_t r[32];
                        ptx::tmem_ld32(lane_t1 + q * 32, r);
                        ptx::tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 32; ++j) logit = fmaf(prm.w4y[m][q * 32 + j], leaky(__uint_as_float(r[j])), logit);
                    }
                    uint32_t r[32];
                    ptx::tmem_ld32(lane_t1 + 128, r);       // column 128 = W4's skip part . f + b4
                    ptx::tmem_ld_wait();
                    logit += __uint_as_float(r[0]);
                }
                release_acc(1);
                ++acc1;
                if (hsel == 0) {
                    const float pred = mask * (1.0f / (1.0f + expf(-logit)));
                    if (m == 0) {
                        pred_lr = pred;
                        write_tail(f_smem, row, zf, pred);  // SuRSNet.py:180: the MASKED pred_lr is channel 321
                        ptx::fence_proxy_async_smem();
                    } else {
                        const int64_t n = tile * TILE_M + row;
                        if (n < io.n) pointio_store(io, n, pred, pred_lr);
                    }
                }
                if (m == 0) {
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive(&bars->f_ready);
                }
            }
        }
    } else if (warp == NEPI) {
        // =============================== MMA issue ========================================
        if (lane == 0) {
            constexpr uint32_t IDESC256 = ptx::umma_idesc_f16(128, 256);
            constexpr uint32_t IDESC144 = ptx::umma_idesc_f16(128, W3_ROWS);
            uint32_t wblk = 0, ablk = 0, acc0 = 0, acc1 = 0, fcnt = 0;
            auto wait_w = [&]() -> uint32_t {
                const uint32_t s = wblk % NSTAGE;
                ptx::mbar_wait(&bars->full_w[s], (wblk / NSTAGE) & 1u, 30, prof);
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
            // only the tail block of F (bias of a layer whose input is not F)
            auto mma_F_tail = [&](uint32_t tmem_d, uint32_t idesc) {
                const uint32_t w = wait_w();
                mma_block(tmem_d, f_smem + (NF_BLK - 1) * A_BLK_BYTES, w, 1, idesc, false);
                release_w();
            };
            // accumulate 4 K blocks of the A ring into tmem_d; hands each slot back through a_free
            auto mma_A = [&](uint32_t tmem_d, uint32_t idesc, bool zero_first) {
                for (int kb = 0; kb < 4; ++kb) {
                    const uint32_t slot = ablk % NA_SLOT;
                    ptx::mbar_wait(&bars->a_ready[slot], (ablk / NA_SLOT) & 1u, 31, prof);
                    ptx::tc_fence_after();
                    const uint32_t w = wait_w();
                    mma_block(tmem_d, a_smem + slot * A_BLK_BYTES, w, 4, idesc, zero_first && kb == 0);
                    release_w();
                    ptx::umma_commit(&bars->a_free[slot]);
                    ++ablk;
                }
            };
            for (int64_t tile = blockIdx.x; tile < prm.ntiles; tile += gridDim.x) {
                for (int m = 0; m < 2; ++m) {
                    ptx::mbar_wait(&bars->f_ready, fcnt & 1u, 32, prof);     // gather done / pred_lr written
                    ++fcnt;
                    ptx::tc_fence_after();
                    for (int h = 0; h < 2; ++h) {
                        for (int c = 0; c < 4; ++c) {
                            ptx::mbar_wait(&bars->acc_free[0], (acc0 & 1u) ^ 1u, 33, prof);
                            ptx::tc_fence_after();
                            mma_F(T0, IDESC256, true);                  // layer 0, chunk c (+ b0)
                            ptx::umma_commit(&bars->acc_full[0]);
                            ++acc0;
                            if (c == 0) {
                                ptx::mbar_wait(&bars->acc_free[1], (acc1 & 1u) ^ 1u, 34, prof);
                                ptx::tc_fence_after();
                            }
                            mma_A(T1, IDESC256, c == 0);                // layer 1, half h, K chunk c
                        }
                        mma_F_tail(T1, IDESC256);                       // + b1 (half h)
                        ptx::umma_commit(&bars->acc_full[1]);
                        ++acc1;
                    }
                    // layer 2 = W2[:, 256:512] y1[1] + W2[:, 0:256] y1[0] + W2[:, 512:] f (+ b2)
                    ptx::mbar_wait(&bars->acc_free[0], (acc0 & 1u) ^ 1u, 35, prof);
                    ptx::tc_fence_after();
                    mma_A(T0, IDESC256, true);
                    mma_A(T0, IDESC256, false);
                    mma_F(T0, IDESC256, false);
                    ptx::umma_commit(&bars->acc_full[0]);
                    ++acc0;
                    // layer 3 (+ W4's skip row) = W3[:, 0:256] y2 + W3[:, 256:] f (+ b3, b4)
                    ptx::mbar_wait(&bars->acc_free[1], (acc1 & 1u) ^ 1u, 36, prof);
                    ptx::tc_fence_after();
                    mma_A(T1, IDESC144, true);
                    mma_F(T1, IDESC144, false);
                    ptx::umma_commit(&bars->acc_full[1]);
                    ++acc1;
                }
            }
        }
    } else if (warp == NEPI + 1) {
        // =============================== weight stream ====================================
        if (lane == 0) {
            uint32_t wblk = 0;
            for (int64_t tile = blockIdx.x; tile < prm.ntiles; tile += gridDim.x) {
                const uint8_t *src = prm.weights;
                for (int b = 0; b < 2 * BLOCKS_PER_MLP; ++b) {
                    const uint32_t bytes = (b % BLOCKS_PER_MLP) < N256_BLOCKS_PER_MLP ? W_BLK_BYTES : W3_BLK_BYTES;
                    const uint32_t s = wblk % NSTAGE;
                    ptx::mbar_wait(&bars->empty_w[s], ((wblk / NSTAGE) & 1u) ^ 1u, 40, prof);
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
                    // fill number 9 (0-based) of the MLP's 11 fills of A = K blocks 36..39 of its 44
                    const uint32_t g0 = ((tiles_done * 2 + m) * A_FILLS_PER_MLP + 9) * 4;
                    ptx::mbar_wait(&bars->scr_ready, scr & 1u, 50, prof);
                    ++scr;
                    for (int kb = 0; kb < 4; ++kb) {
                        const uint32_t g = g0 + kb, slot = g % NA_SLOT;
                        ptx::mbar_wait(&bars->a_free[slot], ((g / NA_SLOT) & 1u) ^ 1u, 51, prof);
                        ptx::mbar_arrive_expect_tx(&bars->a_ready[slot], A_BLK_BYTES);
                        for (int i = 1; i < NEPI; ++i) ptx::mbar_arrive(&bars->a_ready[slot]);
                        ptx::tma_load_1d(smem + SMEM_A + slot * A_BLK_BYTES, scratch + kb * A_BLK_BYTES, A_BLK_BYTES, &bars->a_ready[slot]);
                    }
                }
            }
        }
    }
    if (PROF && threadIdx.x == 0) atomicAdd(g_tc_prof + 0, (unsigned long long)(clock64() - t_kernel0));
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == NEPI) ptx::tmem_dealloc(tmem, 512);
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
    const size_t total = 2 * MLP_BYTES;
    if (!ctx->tc_weights) {
        SURS_CUDA(ctx, cudaMalloc(&ctx->tc_weights, total));
        ctx->tc_weights_bytes = total;
    }
    PackDesc host[2 * BLOCKS_PER_MLP];
    int n = 0;
    for (int m = 0; m < 2; ++m) {
        const int c0 = m == 0 ? SURS_C0_LR : SURS_C0_HR;
        uint32_t off = (uint32_t)(m * MLP_BYTES);
        auto add = [&](int layer, int row0, int nrows, int ntotal, int fblock, int k0, bool extra, bool bias_only) {
            PackDesc d;
            memset(&d, 0, sizeof(d));
            d.w = w[m][layer];
            d.cin = ctx->cin[m][layer];
            d.w_extra = extra ? w[m][4] : nullptr;
            d.bias = ctx->b32[m][layer];
            d.bias_extra = extra ? ctx->b32[m][4] : nullptr;
            d.k0_extra = 128;
            d.row0 = row0; d.nrows = nrows; d.ntotal = ntotal; d.fblock = fblock; d.k0 = k0; d.c0 = c0;
            d.bias_only = bias_only ? 1 : 0;
            d.out_off = off;
            off += (uint32_t)ntotal * 128u;
            host[n++] = d;
        };
        for (int h = 0; h < 2; ++h) {
            for (int c = 0; c < 4; ++c) {
                for (int kb = 0; kb < NF_BLK; ++kb) add(0, c * 256, 256, 256, kb, 0, false, false);            // layer 0 chunk c
                for (int kb = 0; kb < 4; ++kb) add(1, h * 256, 256, 256, -1, c * 256 + kb * 64, false, false);   // layer 1
            }
            add(1, h * 256, 256, 256, NF_BLK - 1, 0, false, true);                                             // b1
        }
        for (int kb = 0; kb < 4; ++kb) add(2, 0, 256, 256, -1, 256 + kb * 64, false, false);        // y1[1]
        for (int kb = 0; kb < 4; ++kb) add(2, 0, 256, 256, -1, kb * 64, false, false);              // y1[0]
        for (int kb = 0; kb < NF_BLK; ++kb) add(2, 0, 256, 256, kb, 512, false, false);             // skip (+ b2)
        for (int kb = 0; kb < 4; ++kb) add(3, 0, 128, W3_ROWS, -1, kb * 64, false, false);          // y2
        for (int kb = 0; kb < NF_BLK; ++kb) add(3, 0, 128, W3_ROWS, kb, 256, true, false);          // skip + W4 row (+ b3, b4)
        if (off != (uint32_t)((m + 1) * MLP_BYTES)) SURS_FAIL(ctx, "internal: weight stream size mismatch");
    }
    PackDesc *dev = nullptr;
    SURS_CUDA(ctx, cudaMalloc(&dev, sizeof(host)));
    SURS_CUDA(ctx, cudaMemcpyAsync(dev, host, sizeof(host), cudaMemcpyHostToDevice, st));
    pack_weights_kernel<<<n, 256, 0, st>>>(dev, (uint8_t *)ctx->tc_weights);
    SURS_LAUNCH_CHECK(ctx, "pack_weights_kernel");
    for (int m = 0; m < 2; ++m)
        SURS_CUDA(ctx, cudaMemcpyAsync(ctx->tc_w4y[m], w[m][4], 128 * sizeof(float), cudaMemcpyDeviceToHost, st));
    SURS_CUDA(ctx, cudaStreamSynchronize(st));
    SURS_CUDA(ctx, cudaFree(dev));
    return 0;
}

int surs_launch_query_tc(surs_ctx *ctx, const PointIO &io, cudaStream_t st)
{
    if (io.n <= 0) return 0;
    TcParams prm;
    prm.weights = (const uint8_t *)ctx->tc_weights;
    memcpy(prm.w4y, ctx->tc_w4y, sizeof(prm.w4y));
    prm.fm.f_lr = ctx->f_lr16; prm.fm.f_hr = ctx->f_hr16;
    prm.fm.H_lr = ctx->H_lr; prm.fm.W_lr = ctx->W_lr; prm.fm.H_hr = ctx->H_hr; prm.fm.W_hr = ctx->W_hr;
    prm.ntiles = (io.n + TILE_M - 1) / TILE_M;
    const int grid = (int)(prm.ntiles < ctx->sm_count ? prm.ntiles : ctx->sm_count);
    if (surs_ensure(ctx, (void **)&ctx->tc_scratch, &ctx->tc_scratch_cap, (size_t)ctx->sm_count * 4 * A_BLK_BYTES)) return 1;
    prm.scratch = (uint8_t *)ctx->tc_scratch;
    static const bool profile = getenv("SURS_TC_PROFILE") != nullptr;
    if (!profile) {
        SURS_CUDA(ctx, cudaFuncSetAttribute(query_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL));
        query_tc_kernel<false><<<grid, NTHREADS, SMEM_TOTAL, st>>>(io, prm);
        SURS_LAUNCH_CHECK(ctx, "query_tc_kernel");
        return 0;
    }
    // SURS_TC_PROFILE=1: synchronous launch of the instrumented kernel + a table of wait cycles per role
    unsigned long long zero[64] = {0}, h[64];
    SURS_CUDA(ctx, cudaMemcpyToSymbol(g_tc_prof, zero, sizeof(zero)));
    SURS_CUDA(ctx, cudaFuncSetAttribute(query_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL));
    query_tc_kernel<true><<<grid, NTHREADS, SMEM_TOTAL, st>>>(io, prm);
    SURS_LAUNCH_CHECK(ctx, "query_tc_kernel<profile>");
    SURS_CUDA(ctx, cudaStreamSynchronize(st));
    SURS_CUDA(ctx, cudaMemcpyFromSymbol(h, g_tc_prof, sizeof(h)));
    const double tiles = (double)(h[2] ? h[2] : 1), k = 1e-3 / tiles;
    fprintf(stderr, "[surs tc profile] tiles=%llu grid=%d kcycles/tile: total %.1f | epi warp0: gather %.1f wait_acc_full(E0 %.1f E1 %.1f E2 %.1f E3 %.1f) "
                    "wait_a_free %.1f | mma: wait_w %.1f wait_a_ready %.1f wait_f %.1f wait_acc_free(%.1f %.1f %.1f %.1f) | loader wait_empty %.1f | reload %.1f %.1f\n",
            h[2], grid, h[0] * k, h[1] * k, h[20] * k, h[21] * k, h[22] * k, h[23] * k, h[10] * k,
            h[30] * k, h[31] * k, h[32] * k, h[33] * k, h[34] * k, h[35] * k, h[36] * k, h[40] * k, h[50] * k, h[51] * k);
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
