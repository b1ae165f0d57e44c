// CTA-pair variant of the one-pass column-factored grid kernel (query_col.cu, P = 1, dense slabs).
//
// Two CTAs of one cluster (the two SMs of a TPC) evaluate two consecutive 128-node tiles with ONE stream of
// tcgen05.mma.cta_group::2 instructions (M = 256: rows 0-127 in the leader's TMEM, 128-255 in the peer's).  Each CTA
// keeps only ITS HALF of every weight block in shared memory (rows N/2 r .. of the block), so per node
//   * the weight traffic L2 -> shared memory halves (2.7 MB per 128-node tile -> 1.35 MB),
//   * the tensor core reads half as many B-operand bytes from shared memory,
//   * one thread issues for two SMs.
// The kernel is power-bound (sw_power_cap at ~1.5 GHz of 1.97), so moving fewer bytes is what buys time.
// Arithmetic, tile order inside a column, K order and accumulation are exactly those of query_col_kernel<P = 1>.
//
// Warp roles (both CTAs): 0-7 produce y0 and run the epilogues, 9 = weight stream (this CTA's halves) + column
// vectors; leader only: 8 = MMA issue (T0 half of layer 1, layers 2 and 3), 10 = MMA issue (T1 half of layer 1);
// peer only: 11 forwards "my half has landed" to the leader's pair barrier.
// Barriers counted across the pair live in the LEADER's shared memory (remote arrivals from the peer); barriers
// completed by tcgen05.commit are multicast to both CTAs.
#include "col_epi.cuh"

#include <stdlib.h>
#include <string.h>

namespace {

using namespace col;

__shared__ int s_prof;
__device__ unsigned long long g_pair_prof[128];       // [64 rank + tag - 100]
// profiling (SURS_COL_ABLATE & 128): cycles spent in the wait with this tag, by thread 0 / the single-thread roles
#define PW(tag, stmt)                                                                    \
    do {                                                                                 \
        if (s_prof && (threadIdx.x == 0 || threadIdx.x >= 256)) {                        \
            const long long t_ = clock64();                                              \
            stmt;                                                                        \
            atomicAdd(g_pair_prof + ((tag) - 100) + 64 * (blockIdx.x & 1), (unsigned long long)(clock64() - t_)); \
        } else {                                                                         \
            stmt;                                                                        \
        }                                                                                \
    } while (0)
// arrive on a barrier that lives in the leader CTA (shared::cluster address).  The plain form: the .release.cluster
// form of the remote arrive made the whole kernel 11 % slower (521 instead of 468 ms at 512^3); the data the arrival
// publishes is read by this CTA's own tensor core, behind fence.proxy.async + the leader's acquire
__device__ __forceinline__ void arrive_x(uint32_t addr)
{
    ptx::mbar_arrive_cluster(addr);
}
__device__ __forceinline__ void wait_x(uint64_t *bar, uint32_t parity, int tag, bool spin)
{
    if (spin) ptx::mbar_wait_cluster(bar, parity, tag);
    else ptx::mbar_wait_cluster_hint(bar, parity, tag);
}

#ifndef SURS_PAIR_NSTAGE
#define SURS_PAIR_NSTAGE 6
#define SURS_PAIR_NASLOT 4
#endif
constexpr int NSTAGE = SURS_PAIR_NSTAGE;                              // weight stages (half blocks)
constexpr int NA_SLOT = SURS_PAIR_NASLOT;
constexpr int WH_BYTES = W_BLK_BYTES / 2;              // this CTA's 128 rows of a 256 x 64 block
constexpr int WH128_BYTES = W128_BLK_BYTES / 2;        // this CTA's 64 rows of a 128 x 64 block (layer 3)
constexpr int SMEM_W = 0;
constexpr int SMEM_A = SMEM_W + NSTAGE * WH_BYTES;
constexpr int SMEM_CV = SMEM_A + NA_SLOT * A_BLK_BYTES;
constexpr int SMEM_GV = SMEM_CV + 2 * CV_BYTES;
constexpr int SMEM_BAR = SMEM_GV + GV_BYTES;
constexpr int SMEM_PREDX = SMEM_BAR + 512;
constexpr int SMEM_TOTAL = SMEM_PREDX + 512 + 1024;
static_assert(SMEM_TOTAL <= 232448, "shared memory budget");
constexpr int BLOCKS_PER_MLP = 32 + 8 + 4;
constexpr int NEPI = 8;
constexpr int NTHREADS = (NEPI + 4) * 32;

struct Bars {
    uint64_t full_w[NSTAGE];                 // peer: its TMA half has landed (the forwarder waits here)
    uint64_t pair_w[NSTAGE], pair_wb[NSTAGE];// leader: both halves have landed (leader's TMA + the peer's forwarder)
    uint64_t empty_w[NSTAGE];                // each CTA: multicast commit of the consuming issuer
    uint64_t a_ready[NA_SLOT], a_ready_b[NA_SLOT];   // leader: 2 x NEPI warps
    uint64_t a_free[NA_SLOT];                // each CTA: two (multicast) arrivals per phase
    uint64_t acc_full[2];                    // each CTA: multicast commit
    uint64_t acc_free[2], t1_free_b;         // leader: 2 x NEPI warps
    uint64_t cv_full[2], cv_empty[2];        // each CTA
    uint32_t tmem_base;
};
static_assert(sizeof(Bars) <= 512, "barrier block");

struct Epi {
    uint32_t a_ready0, a_ready_b0, acc_free0;            // the leader's barrier arrays (shared::cluster addresses)
    uint64_t *a_free;
    uint32_t a_smem;
    int row, hsel, lane;
    float zf, pred;
    uint32_t g;
};

__device__ __forceinline__ uint32_t ring_acquire(Epi &e)
{
    const uint32_t slot = e.g % NA_SLOT;
    PW(110, ptx::mbar_wait(&e.a_free[slot], ((e.g / NA_SLOT) & 1u) ^ 1u, 110));
    return slot;
}
__device__ __forceinline__ void ring_publish(Epi &e, uint32_t slot, bool layer0)
{
    ptx::fence_proxy_async_smem();
    __syncwarp();
    if (e.lane == 0) {
        arrive_x(e.a_ready0 + 8u * slot);
        if (layer0) arrive_x(e.a_ready_b0 + 8u * slot);
    }
    ++e.g;
}

// epilogue of a 256-column accumulator into 4 K blocks of the A ring (query_col.cu: epilogue_256)
template <bool HAS_Z, bool HAS_P>
__device__ __forceinline__ void epilogue_256(Epi &e, uint32_t taddr, int acc_id, const float *add, const float *wz, const float *wp)
{
    uint32_t r[2][32];
    ptx::tmem_ld32(taddr + e.hsel * 32, r[0]);
#pragma unroll
    for (int kb = 0; kb < 4; ++kb) {
        ptx::tmem_ld_wait();
        if (kb < 3) {
            ptx::tmem_ld32(taddr + (kb + 1) * 64 + e.hsel * 32, r[(kb + 1) & 1]);
        } else {
            ptx::tc_fence_before();
            __syncwarp();
            if (e.lane == 0) arrive_x(e.acc_free0 + 8u * acc_id);
        }
        const int c = kb * 64 + e.hsel * 32;
        const uint32_t slot = ring_acquire(e);
        finish32<1, true, HAS_Z, HAS_P>(r[kb & 1], add + c, wz + c, wp + c, e.zf, e.pred, e.a_smem + slot * A_BLK_BYTES, e.row, e.hsel, 0);
        ring_publish(e, slot, false);
    }
}

__device__ __forceinline__ void mma_block2(uint32_t tmem_d, uint32_t a_addr, uint32_t w_addr, uint32_t idesc, bool zero_first)
{
    const uint64_t da = ptx::umma_desc_sw128(a_addr), db = ptx::umma_desc_sw128(w_addr);
#pragma unroll 1
    for (int k = 0; k < 4; ++k)
        ptx::umma2_f16(tmem_d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (zero_first && k == 0) ? 0u : 1u);
}

__global__ void __launch_bounds__(NTHREADS, 1) query_col_pair_kernel(const __grid_constant__ PointIO io, const __grid_constant__ ColParams prm)
{
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = ptx::smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t *smem = smem_raw + (base - raw);
    const uint32_t a_smem = base + SMEM_A, w_smem = base + SMEM_W;
    float *cv_s = reinterpret_cast<float *>(smem + SMEM_CV);
    float *gv_s = reinterpret_cast<float *>(smem + SMEM_GV);
    Bars *bars = reinterpret_cast<Bars *>(smem + SMEM_BAR);
    float *pred_x = reinterpret_cast<float *>(smem + SMEM_PREDX);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = ptx::cluster_ctarank();
    const bool leader = rank == 0;
    const int64_t npairs = (prm.ntiles + 1) / 2;
    const int64_t cl = blockIdx.x >> 1, ncl = gridDim.x >> 1;
    const long long t_kernel0 = clock64();

    if (threadIdx.x == 0) {
        s_prof = (prm.ablate & 128) != 0;
        for (int s = 0; s < NSTAGE; ++s) {
            ptx::mbar_init(&bars->full_w[s], 1); ptx::mbar_init(&bars->empty_w[s], 1);
            ptx::mbar_init(&bars->pair_w[s], 2); ptx::mbar_init(&bars->pair_wb[s], 2);
        }
        for (int k = 0; k < NA_SLOT; ++k) {
            ptx::mbar_init(&bars->a_ready[k], 2 * NEPI); ptx::mbar_init(&bars->a_ready_b[k], 2 * NEPI);
            ptx::mbar_init(&bars->a_free[k], 2);
        }
        ptx::mbar_init(&bars->t1_free_b, 2 * NEPI);
        for (int t = 0; t < 2; ++t) {
            ptx::mbar_init(&bars->acc_full[t], 1); ptx::mbar_init(&bars->acc_free[t], 2 * NEPI);
            ptx::mbar_init(&bars->cv_full[t], 1); ptx::mbar_init(&bars->cv_empty[t], NEPI);
        }
        ptx::fence_barrier_init();
    }
    for (int i = threadIdx.x; i < GV_BYTES / 16; i += NTHREADS)          // constant vectors: resident for the whole kernel
        reinterpret_cast<uint4 *>(gv_s)[i] = __ldg(reinterpret_cast<const uint4 *>(prm.gv) + i);
    ptx::cluster_sync();                                                 // both CTAs' barriers exist before any remote arrive
    if (warp == NEPI) ptx::tmem_alloc2(&bars->tmem_base, 512);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = bars->tmem_base;
    const uint32_t T0 = tmem, T1 = tmem + 256;

    if (warp < NEPI) {
        // =============================== y0 production + epilogues ========================
        Epi e;
        e.a_ready0 = ptx::map_to_cta(ptx::smem_u32(&bars->a_ready[0]), 0);
        e.a_ready_b0 = ptx::map_to_cta(ptx::smem_u32(&bars->a_ready_b[0]), 0);
        e.acc_free0 = ptx::map_to_cta(ptx::smem_u32(&bars->acc_free[0]), 0);
        const uint32_t t1_free0 = ptx::map_to_cta(ptx::smem_u32(&bars->t1_free_b), 0);
        e.a_free = bars->a_free; e.a_smem = a_smem; e.lane = lane; e.g = 0;
        const int quarter = warp & 3;
        e.hsel = warp >> 2;
        e.row = quarter * 32 + lane;
        const uint32_t lane_t0 = T0 + ((uint32_t)(quarter * 32) << 16), lane_t1 = T1 + ((uint32_t)(quarter * 32) << 16);
        uint32_t acc0 = 0, acc1 = 0, it = 0;
        for (int64_t tp = cl; tp < npairs; tp += ncl, ++it) {
            int64_t tile = 2 * tp + rank;
            const bool valid = tile < prm.ntiles;
            if (!valid) tile = prm.ntiles - 1;                           // odd tile count: the peer recomputes the last tile, stores nothing
            const int64_t col = tile / prm.nseg;
            const int seg = (int)(tile - col * prm.nseg);
            const int k = seg * TILE_M + e.row;
            const int kc = k < prm.R2 ? k : prm.R2 - 1;
            const int i = prm.plane_lo + (int)(col / prm.R1), j = (int)(col % prm.R1);
            const Projected pr = project_point(io, (float)io.axis[0][i], (float)io.axis[1][j], (float)io.axis[2][kc]);
            float zf4[4], pred4[4] = {0.0f, 0.0f, 0.0f, 0.0f};            // layer 0 works on rows lane + 32 r
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int kr = seg * TILE_M + lane + 32 * r;
                zf4[r] = project_point(io, (float)io.axis[0][i], (float)io.axis[1][j], (float)io.axis[2][kr < prm.R2 ? kr : prm.R2 - 1]).zf;
            }
            const uint32_t cvb = it & 1u;
            PW(111, ptx::mbar_wait(&bars->cv_full[cvb], (it >> 1) & 1u, 111));
            const float *cv = cv_s + cvb * CV_FLOATS;
            e.zf = pr.zf;
            float pred_lr = 0.0f;
#pragma unroll 1
            for (int m = 0; m < 2; ++m) {
                const float *cvm = cv + m * CV_STRIDE, *gvm = gv_s + m * GV_STRIDE;
                e.pred = pred_lr;
                // layer 0 on the CUDA cores: 16 K blocks of y0 = leaky(C0 + w_z z (+ w_p pred_lr))
#pragma unroll 1
                for (int kb = 0; kb < 16; ++kb) {
                    const uint32_t slot = ring_acquire(e);
                    const int c = kb * 64 + warp * 8;                    // warp w fills 16-byte chunk w of all 128 rows
                    const uint32_t dst = a_smem + slot * A_BLK_BYTES;
                    if (m == 0) produce8<1, false>(cvm + CV_C0 + c, gvm + GV_WZ0 + c, nullptr, zf4, pred4, dst, lane, warp, 0);
                    else produce8<1, true>(cvm + CV_C0 + c, gvm + GV_WZ0 + c, gvm + GV_WP0 + c, zf4, pred4, dst, lane, warp, 0);
                    ring_publish(e, slot, true);
                }
                // E1: layer 1, both halves (bias b1) -> A ring
                PW(120, ptx::mbar_wait(&bars->acc_full[0], acc0 & 1u, 120));
                ptx::tc_fence_after();
                epilogue_256<false, false>(e, lane_t0, 0, gvm + GV_B1, nullptr, nullptr);
                ++acc0;
                PW(121, ptx::mbar_wait(&bars->acc_full[1], acc1 & 1u, 121));
                ptx::tc_fence_after();
                epilogue_256<false, false>(e, lane_t1, 1, gvm + GV_B1 + 256, nullptr, nullptr);
                ++acc1;
                // E2: layer 2 (T0) + skip terms -> A ring
                PW(122, ptx::mbar_wait(&bars->acc_full[0], acc0 & 1u, 122));
                ptx::tc_fence_after();
                if (m == 0) epilogue_256<true, false>(e, lane_t0, 0, cvm + CV_C2, gvm + GV_WZ2, nullptr);
                else epilogue_256<true, true>(e, lane_t0, 0, cvm + CV_C2, gvm + GV_WZ2, gvm + GV_WP2);
                ++acc0;
                // E3: layer 3 (T1) + skip terms, layer 4 (each warp 64 of the 128 channels), sigmoid (warps 0-3)
                PW(123, ptx::mbar_wait(&bars->acc_full[1], acc1 & 1u, 123));
                ptx::tc_fence_after();
                float logit = 0.0f;
                {
                    uint32_t r[2][32];
                    const int cb = e.hsel * 64;
                    ptx::tmem_ld32(lane_t1 + cb, r[0]);
                    ptx::tmem_ld32(lane_t1 + cb + 32, r[1]);
                    ptx::tmem_ld_wait();
                    float lg[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
#pragma unroll
                        for (int j4 = 0; j4 < 8; ++j4) {
                            const int c = cb + q * 32 + 4 * j4;
                            const float4 a = *reinterpret_cast<const float4 *>(cvm + CV_C3 + c), z = *reinterpret_cast<const float4 *>(gvm + GV_WZ3 + c);
                            const float4 w4 = *reinterpret_cast<const float4 *>(gvm + GV_W4Y + c);
                            float v0 = __uint_as_float(r[q][4 * j4]) + a.x + z.x * e.zf, v1 = __uint_as_float(r[q][4 * j4 + 1]) + a.y + z.y * e.zf;
                            float v2 = __uint_as_float(r[q][4 * j4 + 2]) + a.z + z.z * e.zf, v3 = __uint_as_float(r[q][4 * j4 + 3]) + a.w + z.w * e.zf;
                            if (m == 1) {
                                const float4 p = *reinterpret_cast<const float4 *>(gvm + GV_WP3 + c);
                                v0 = fmaf(p.x, e.pred, v0); v1 = fmaf(p.y, e.pred, v1); v2 = fmaf(p.z, e.pred, v2); v3 = fmaf(p.w, e.pred, v3);
                            }
                            lg[0] = fmaf(w4.x, leaky(v0), lg[0]); lg[1] = fmaf(w4.y, leaky(v1), lg[1]);
                            lg[2] = fmaf(w4.z, leaky(v2), lg[2]); lg[3] = fmaf(w4.w, leaky(v3), lg[3]);
                        }
                    }
                    logit = (lg[0] + lg[1]) + (lg[2] + lg[3]);
                }
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    arrive_x(e.acc_free0 + 8u);
                    arrive_x(t1_free0);
                }
                ++acc1;
                if (e.hsel == 1) pred_x[e.row] = logit;
                asm volatile("bar.sync %0, 64;" ::"r"(2 + quarter) : "memory");
                if (e.hsel == 0) {
                    logit = (cvm[CV_C4] + gvm[GV_WZ4] * e.zf + (m == 1 ? gvm[GV_WP4] * e.pred : 0.0f)) + (logit + pred_x[e.row]);
                    const float pred = pr.mask * (1.0f / (1.0f + expf(-logit)));
                    if (m == 0) {
                        pred_lr = pred;
                        pred_x[e.row] = pred;
                    } else if (valid && k < prm.R2) {
                        const int64_t n = col * prm.R2 + k;
                        io.out_hr[n] = pred;
                        io.out_lr[n] = pred_lr;
                    }
                }
                // the HR pass needs the pred_lr of rows lane + 32 r (layer 0) and of the warp's own row
                if (m == 0) {
                    asm volatile("bar.sync 1, 256;" ::: "memory");
                    pred_lr = pred_x[e.row];
#pragma unroll
                    for (int r = 0; r < 4; ++r) pred4[r] = pred_x[lane + 32 * r];
                }
            }
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&bars->cv_empty[cvb]);
        }
    } else if (warp == NEPI) {
        // =============================== MMA issue (leader): T0 half of layer 1, layers 2 and 3 ============
        if (lane == 0 && leader) {
            const bool spin = prm.ablate & 32;
            constexpr uint32_t IDESC256 = ptx::umma_idesc_f16(256, 256);
            constexpr uint32_t IDESC128 = ptx::umma_idesc_f16(256, 128);
            const uint32_t a_free_peer = ptx::map_to_cta(ptx::smem_u32(&bars->a_free[0]), 1);
            uint32_t wblk = 0, ablk = 0, acc0 = 0, acc1 = 0, wph = 0;     // wph: per-slot phase of pair_w (this thread's fills only)
            auto wait_w = [&]() -> uint32_t {
                const uint32_t s = wblk % NSTAGE;
                PW(130, wait_x(&bars->pair_w[s], (wph >> s) & 1u, 130, spin));
                wph ^= 1u << s;
                ptx::tc_fence_after();
                return w_smem + s * WH_BYTES;
            };
            auto release_w = [&]() {
                ptx::umma2_commit(&bars->empty_w[wblk % NSTAGE], 0x3);
                ++wblk;
            };
            auto wait_a = [&]() -> uint32_t {
                const uint32_t slot = ablk % NA_SLOT;
                PW(131, wait_x(&bars->a_ready[slot], (ablk / NA_SLOT) & 1u, 131, spin));
                ptx::tc_fence_after();
                return slot;
            };
            // a_free expects two arrivals per CTA: in layer 1 one multicast commit from each issuing thread, elsewhere a
            // plain arrive in both CTAs plus this thread's commit (the phase completes when the commit lands)
            auto release_a = [&](uint32_t slot, int commits) {
                if (commits == 2) {
                    ptx::mbar_arrive(&bars->a_free[slot]);
                    ptx::mbar_arrive_cluster(a_free_peer + 8u * slot);
                }
                ptx::umma2_commit(&bars->a_free[slot], 0x3);
                ++ablk;
            };
            for (int64_t tp = cl; tp < npairs; tp += ncl) {
                for (int m = 0; m < 2; ++m) {
                    long long tp0 = s_prof ? clock64() : 0;
                    auto phase = [&](int slot) {
                        if (s_prof) {
                            const long long t1 = clock64();
                            atomicAdd(g_pair_prof + slot, (unsigned long long)(t1 - tp0));
                            tp0 = t1;
                        }
                    };
                    // layer 1, T0 half: K = 1024 (16 blocks), N = 256
                    PW(133, wait_x(&bars->acc_free[0], (acc0 & 1u) ^ 1u, 133, spin));
                    PW(134, wait_x(&bars->acc_free[1], (acc1 & 1u) ^ 1u, 134, spin));    // keeps this thread's phase count of acc_free[1]
                    ptx::tc_fence_after();
                    for (int kb = 0; kb < 16; ++kb) {
                        const uint32_t slot = wait_a();
                        const uint32_t w = wait_w();
                        if (!(prm.ablate & 4)) mma_block2(T0, a_smem + slot * A_BLK_BYTES, w, IDESC256, kb == 0);
                        release_w();
                        ++wblk;                                   // the odd block belongs to the other thread
                        release_a(slot, 1);
                    }
                    ptx::umma2_commit(&bars->acc_full[0], 0x3);
                    ++acc0; ++acc1;
                    phase(50);
                    // layer 2: K = 512 (y1 halves from E1), N = 256 -> T0
                    PW(135, wait_x(&bars->acc_free[0], (acc0 & 1u) ^ 1u, 135, spin));
                    ptx::tc_fence_after();
                    for (int kb = 0; kb < 8; ++kb) {
                        const uint32_t slot = wait_a();
                        const uint32_t w = wait_w();
                        mma_block2(T0, a_smem + slot * A_BLK_BYTES, w, IDESC256, kb == 0);
                        release_w();
                        release_a(slot, 2);
                    }
                    ptx::umma2_commit(&bars->acc_full[0], 0x3);
                    ++acc0;
                    phase(51);
                    // layer 3: K = 256, N = 128 -> T1
                    PW(136, wait_x(&bars->acc_free[1], (acc1 & 1u) ^ 1u, 136, spin));
                    ptx::tc_fence_after();
                    for (int kb = 0; kb < 4; ++kb) {
                        const uint32_t slot = wait_a();
                        const uint32_t w = wait_w();
                        mma_block2(T1, a_smem + slot * A_BLK_BYTES, w, IDESC128, kb == 0);
                        release_w();
                        release_a(slot, 2);
                    }
                    ptx::umma2_commit(&bars->acc_full[1], 0x3);
                    ++acc1;
                    phase(52);
                }
            }
        }
    } else if (warp == NEPI + 2) {
        // =============================== MMA issue (leader), T1 half of layer 1 ======================
        if (lane == 0 && leader) {
            const bool spin = prm.ablate & 32;
            constexpr uint32_t IDESC256 = ptx::umma_idesc_f16(256, 256);
            uint32_t wblk = 1, ablk = 0, wph = 0, aph = 0, tph = 0;
            for (int64_t tp = cl; tp < npairs; tp += ncl) {
                for (int m = 0; m < 2; ++m) {
                    PW(137, wait_x(&bars->t1_free_b, tph ^ 1u, 137, spin));          // E3 of the previous pass has read T1 (both CTAs)
                    tph ^= 1u;
                    ptx::tc_fence_after();
                    for (int kb = 0; kb < 16; ++kb) {
                        const uint32_t slot = ablk % NA_SLOT, s = wblk % NSTAGE;
                        PW(138, wait_x(&bars->a_ready_b[slot], (aph >> slot) & 1u, 138, spin));
                        aph ^= 1u << slot;
                        PW(139, wait_x(&bars->pair_wb[s], (wph >> s) & 1u, 139, spin));
                        wph ^= 1u << s;
                        ptx::tc_fence_after();
                        if (!(prm.ablate & 4)) mma_block2(T1, a_smem + slot * A_BLK_BYTES, w_smem + s * WH_BYTES, IDESC256, kb == 0);
                        ptx::umma2_commit(&bars->empty_w[s], 0x3);
                        ptx::umma2_commit(&bars->a_free[slot], 0x3);
                        wblk += 2;
                        ++ablk;
                    }
                    ptx::umma2_commit(&bars->acc_full[1], 0x3);
                    wblk += 12; ablk += 12;                       // layers 2 and 3
                }
            }
        }
    } else if (warp == NEPI + 1) {
        // =============================== weight stream (this CTA's halves) + column vectors ====================
        if (lane == 0) {
            uint32_t wblk = 0, it = 0;
            for (int64_t tp = cl; tp < npairs; tp += ncl, ++it) {
                int64_t tile = 2 * tp + rank;
                if (tile >= prm.ntiles) tile = prm.ntiles - 1;
                const uint32_t cvb = it & 1u;
                PW(141, ptx::mbar_wait(&bars->cv_empty[cvb], ((it >> 1) & 1u) ^ 1u, 141));
                ptx::mbar_arrive_expect_tx(&bars->cv_full[cvb], CV_BYTES);
                ptx::tma_load_1d(smem + SMEM_CV + cvb * CV_BYTES, prm.table + (tile / prm.nseg) * CV_ROW_FLOATS, CV_BYTES, &bars->cv_full[cvb]);
                const uint8_t *src = prm.weights;
                for (int b = 0; b < 2 * BLOCKS_PER_MLP; ++b) {
                    const int bm = b % BLOCKS_PER_MLP;
                    const uint32_t half = bm < 40 ? WH_BYTES : WH128_BYTES;
                    const uint32_t s = wblk % NSTAGE;
                    PW(140, ptx::mbar_wait(&bars->empty_w[s], ((wblk / NSTAGE) & 1u) ^ 1u, 140));
                    // the leader's half completes on the pair barrier directly; the peer's on full_w, forwarded by warp NEPI + 3
                    uint64_t *full = !leader ? &bars->full_w[s] : (bm < 32 && (b & 1)) ? &bars->pair_wb[s] : &bars->pair_w[s];
                    ptx::mbar_arrive_expect_tx(full, half);
                    ptx::tma_load_1d(smem + SMEM_W + s * WH_BYTES, src + rank * half, half, full);
                    src += 2 * half;
                    ++wblk;
                }
            }
        }
    } else {
        // =============================== peer: "my half has landed" -> the leader's pair barrier ==============
        if (lane == 0 && !leader) {
            const uint32_t pair_w0 = ptx::map_to_cta(ptx::smem_u32(&bars->pair_w[0]), 0);
            const uint32_t pair_wb0 = ptx::map_to_cta(ptx::smem_u32(&bars->pair_wb[0]), 0);
            uint32_t wblk = 0;
            for (int64_t tp = cl; tp < npairs; tp += ncl) {
                for (int b = 0; b < 2 * BLOCKS_PER_MLP; ++b) {
                    const int bm = b % BLOCKS_PER_MLP;
                    const uint32_t s = wblk % NSTAGE;
                    PW(142, ptx::mbar_wait(&bars->full_w[s], (wblk / NSTAGE) & 1u, 142));
                    arrive_x(((bm < 32 && (b & 1)) ? pair_wb0 : pair_w0) + 8u * s);
                    ++wblk;
                }
            }
        }
    }
    if (s_prof && threadIdx.x == 0 && leader) atomicAdd(g_pair_prof + 0, (unsigned long long)(clock64() - t_kernel0));
    ptx::tc_fence_before();
    ptx::cluster_sync();                       // the peer's barriers and TMEM stay alive until the leader's last commit has landed
    if (warp == NEPI) ptx::tmem_dealloc2(tmem, 512);
}

}  // namespace

// Dense slab, one pass (SURS_PREC_FP16 and the first pass of SURS_PREC_FP16R): prm as prepared by
// surs_launch_query_col (table built, weights = the one-pass main stream).
int surs_launch_query_col_pair(surs_ctx *ctx, const PointIO &io, const col::ColParams &prm, cudaStream_t st)
{
    const int64_t npairs = (prm.ntiles + 1) / 2;
    const int64_t max_cl = ctx->sm_count / 2;
    const int ncl = (int)(npairs < max_cl ? npairs : max_cl);
    SURS_CUDA(ctx, cudaFuncSetAttribute(query_col_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL));
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(2 * ncl);
    cfg.blockDim = dim3(NTHREADS);
    cfg.dynamicSmemBytes = SMEM_TOTAL;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    SURS_CUDA(ctx, cudaLaunchKernelEx(&cfg, query_col_pair_kernel, io, prm));
    SURS_LAUNCH_CHECK(ctx, "query_col_pair_kernel");
    if (prm.ablate & 128) {
        unsigned long long h[128];
        SURS_CUDA(ctx, cudaStreamSynchronize(st));
        SURS_CUDA(ctx, cudaMemcpyFromSymbol(h, g_pair_prof, sizeof(h)));
        const double k = 1e-3 / (double)npairs;
        fprintf(stderr, "[surs pair profile] pairs=%lld clusters=%d kcycles/pair: total %.1f | mma phases: L1 %.1f L2 %.1f L3 %.1f\n", (long long)npairs, ncl, h[0] * k, h[50] * k, h[51] * k, h[52] * k);
        for (int r = 0; r < 2; ++r) {
            const unsigned long long *q = h + 64 * r;
            fprintf(stderr, "[surs pair profile] rank %d: epi thread 0: a_free %.1f cv %.1f acc_full(E1a %.1f E1b %.1f E2 %.1f E3 %.1f) | issuer A: w %.1f a %.1f acc_free(%.1f %.1f %.1f %.1f) | "
                            "issuer B: t1 %.1f a %.1f w %.1f | loader: empty %.1f cv %.1f | forwarder: full %.1f\n",
                    r, q[10] * k, q[11] * k, q[20] * k, q[21] * k, q[22] * k, q[23] * k, q[30] * k, q[31] * k, q[33] * k, q[34] * k, q[35] * k, q[36] * k,
                    q[37] * k, q[38] * k, q[39] * k, q[40] * k, q[41] * k, q[42] * k);
        }
        unsigned long long zero[128] = {0};
        SURS_CUDA(ctx, cudaMemcpyToSymbol(g_pair_prof, zero, sizeof(zero)));
    }
    return 0;
}
