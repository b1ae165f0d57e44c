// Dense grid evaluation with layer 1 updated incrementally along each column (SURS_PREC_FP16,
// surs_eval_grid without a transform; same preconditions as query_col.cu).
//
// Along a k-column the layer-0 pre-activation of channel c is  pre_c(k) = C0_c + wz_c z(k) + wp_c p(k)
// (C0 from the per-column table; z = depth feature, p = masked LR prediction, HR MLP only), and
// leaky(x) = 0.01 x + 0.99 relu(x) (lib/model/SurfaceClassifier.py:57: F.leaky_relu).  Layer 1 is
// therefore
//     y1pre[n](k) = P_n + Q_n z(k) + R_n p(k),   (P, Q, R)_n = sum_c s_c W1[n,c] (C0_c, wz_c, wp_c) (+ b1),
// with s_c = 1 where pre_c(k) > 0 and 0.01 elsewhere.  (P, Q, R) only change when a channel changes
// sign ("event"): a rank-1 update with column c of W1.  A column of 512 nodes sees a few hundred
// events per MLP instead of 512 x 1024 x 512 MACs, so the K = 1024 GEMM of layer 1 -- 76 % of the
// MACs of query_col.cu -- disappears from the tensor cores and layer 1 is exact fp32 arithmetic.
// Layers 2 and 3 stay on tcgen05 as in query_col.cu; the skip terms enter through the column table.
//
// Per tile (128 consecutive k of one column; a CTA walks whole columns so the state carries over):
//   detect : sign changes of all 1024 channels over the tile -> events, sorted by (k, c)
//   emit   : thread (warp w, lane l) owns output channels 64 w + 2 l, +1: walks k, applies the events
//            of step k (fp32 FMAs with W1^T rows read from L2), writes leaky(y1pre) as fp16 into
//            K block w of the UMMA operand tile
//   L2, E2, L3, E3 : tensor-core layers 2 / 3 and their epilogues (layer 4 + sigmoid in E3)
// Every column starts from the all-negative state (P = C1 of the table, Q / R constant vectors).
//
// Warp roles: 0-7 workers (detect, emit, epilogues), 8 = MMA issue, 9 = weight stream + column vectors.
#include "col_common.cuh"

#include <stdlib.h>

namespace {

using namespace col;

constexpr int NWORK = 8;
constexpr int NWT = NWORK * 32;
constexpr int NTHREADS = NWT + 64;
constexpr int NSTAGE = 3;
constexpr int EV_CAP = 1024;               // >= 1024 so that a single step always fits
constexpr int EV_PAD = 16;
constexpr int EV_GROUP = 16;
constexpr int NGRP = 4;                    // detection works on groups of 32 rows
constexpr int SCAN_CAP = 1024;
constexpr uint32_t EV_SENTINEL = 128u << 11;

constexpr int SM_Y = 0;                                        // 8 K blocks: y1, then y2 in blocks 0-3
constexpr int SM_W = SM_Y + 8 * A_BLK_BYTES;
constexpr int SM_CV = SM_W + NSTAGE * W128_BLK_BYTES;
constexpr int SM_XV = SM_CV + CV_BYTES;
constexpr int SM_EV = SM_XV + XV_BYTES;                        // sorted events (scan list during detection)
constexpr int SM_STAGE = SM_EV + (EV_CAP + EV_PAD) * 4;
constexpr int SM_BUCKET = SM_STAGE + EV_CAP * 4;
constexpr int SM_CNT = SM_BUCKET + EV_CAP * 4;                 // cnt[132], fill[132], start[132], ctr[4], bm[4][32], pre[4][36]
constexpr int SM_SGN = SM_CNT + (3 * 132 + 4 + 128 + 144) * 4;           // sign of layer-0 channels: [mlp][2][1024] bytes
constexpr int SM_ZP = SM_SGN + 4096;                           // float2 (z_feat, pred_lr) per row
constexpr int SM_MASK = SM_ZP + 1024;                          // mask[128], partial logits[128]
constexpr int SM_BAR = SM_MASK + 1024;
constexpr int SM_TOTAL = SM_BAR + 256 + 1024;
static_assert(SM_TOTAL <= 232448, "shared memory budget");
static_assert(SM_XV % 16 == 0 && SM_EV % 16 == 0 && SM_CV % 16 == 0, "alignment");

struct Bars {
    uint64_t full_w[NSTAGE], empty_w[NSTAGE];
    uint64_t y1_ready, acc2_full, y2_ready, acc3_full;
    uint64_t cv_full, cv_empty;
    uint32_t tmem_base;
};

struct IncParams {
    const uint8_t *weights;        // 2 x XW_MLP_BYTES
    const float *xv;               // [XV_FLOATS]
    const float *table;            // [ncols][CV_ROW_FLOATS]
    const __half *w1h[2];          // fp16 W1^T: [1024][512]
    const float *qstar, *rstar;    // [2][512]
    int64_t ncols;
    int nseg;                      // tiles per column = ceil(R2 / 128)
    int R1, R2, plane_lo;
};

struct Smem {
    uint32_t y, w;
    const float *cv, *xv;
    uint32_t *ev, *stage, *bucket;
    uint16_t *scan;
    int *cnt, *fill, *start, *ctr;
    uint8_t *sgn;
    float2 *zp;
    float *mask, *part;
    Bars *bars;
};

__device__ __forceinline__ void bar_workers() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

__device__ __forceinline__ void push_event(const Smem &s, int k, int c, int sign)
{
    const int pos = atomicAdd(&s.ctr[0], 1);
    if (pos < EV_CAP) {
        s.stage[pos] = ((uint32_t)k << 11) | ((uint32_t)c << 1) | (uint32_t)sign;
        atomicAdd(&s.cnt[k], 1);
    }
}

// Sign changes of the 1024 layer-0 channels over rows [k0, k1) -> s.ev, sorted by row; inside a row
// first the channels settled by the interval tests (by channel), then the row-scanned ones (by
// channel): a deterministic summation order for the fp32 updates.  EV_PAD sentinels follow.
// Returns the number of events or -1 when they do not fit.
//
// Rows are handled in groups of 32.  For each (channel, group) an interval bound of the
// pre-activation settles the sign for the whole group, or the pair is scanned row by row.  The
// tracked sign (s.sgn) is what the state (P, Q, R) reflects; because leaky() is continuous at 0, a
// sign taken from the bound instead of the row's own rounding costs O(ulp) and nothing more.
template <int M>
__device__ __forceinline__ int detect(const Smem &s, int k0, int k1, int cur, int tid)
{
    const int warp = tid >> 5, lane = tid & 31;
    const float *cvm = s.cv + M * CV_STRIDE, *xvm = s.xv + M * XV_LR_FLOATS;
    const uint8_t *sc = s.sgn + M * 2048 + cur * 1024;
    uint8_t *sn = s.sgn + M * 2048 + (cur ^ 1) * 1024;
    uint32_t *bm = reinterpret_cast<uint32_t *>(s.ctr + 4);      // [NGRP][32]: channels that flip at the first row of a group
    int *pre = s.ctr + 4 + NGRP * 32;                            // [NGRP][36]: exclusive prefix of popc(bm)
    bar_workers();                                               // previous users of the scratch are done
    for (int i = tid; i < 2 * 132; i += NWT) s.cnt[i] = 0;       // cnt and fill are adjacent
    if (tid < 4) s.ctr[tid] = 0;
    int ga[NGRP], gb[NGRP];
    float zlo[NGRP], zhi[NGRP], plo[NGRP], phi[NGRP];
#pragma unroll
    for (int g = 0; g < NGRP; ++g) {
        ga[g] = 32 * g > k0 ? 32 * g : k0;
        gb[g] = 32 * g + 32 < k1 ? 32 * g + 32 : k1;
        zlo[g] = zhi[g] = plo[g] = phi[g] = 0.0f;
        if (ga[g] < gb[g]) {
            const float za = s.zp[ga[g]].x, zb = s.zp[gb[g] - 1].x;      // z_feat is monotonic along the column
            zlo[g] = fminf(za, zb); zhi[g] = fmaxf(za, zb);
            if (M) {
                const int kk = 32 * g + lane;
                const bool in = kk >= ga[g] && kk < gb[g];
                float lo = in ? s.zp[kk].y : 2.0f, hi = in ? s.zp[kk].y : -1.0f;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
                    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
                }
                plo[g] = lo; phi[g] = hi;
            }
        }
    }
    bar_workers();
    uint32_t flip_bits = 0, sign_bits = 0;                       // bit 4 j + g
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int c = tid + NWT * j;
        const float a = cvm[CV_C0 + c], b = xvm[XV_WZ0 + c], p = M ? xvm[XV_WP0 + c] : 0.0f;
        int prev = sc[c];
#pragma unroll
        for (int g = 0; g < NGRP; ++g) {
            bool flip = false;
            int now = prev;
            if (ga[g] < gb[g]) {
                const float b0 = b * zlo[g], b1 = b * zhi[g];
                float lo = a + fminf(b0, b1), hi = a + fmaxf(b0, b1);
                if (M) {
                    const float p0 = p * plo[g], p1 = p * phi[g];
                    lo += fminf(p0, p1); hi += fmaxf(p0, p1);
                }
                if (lo > 0.0f) {
                    now = 1;
                } else if (hi <= 0.0f) {
                    now = 0;
                } else {                                          // scan the rows; the group ends with its last row's sign
                    const int idx = atomicAdd(&s.ctr[1], 1);
                    if (idx < SCAN_CAP) s.scan[idx] = (uint16_t)(c | (g << 10) | (prev << 12));
                    const float2 zp = s.zp[gb[g] - 1];
                    now = fmaf(p, zp.y, fmaf(b, zp.x, a)) > 0.0f;
                    prev = now;                                   // flips inside the group are the scanner's
                }
                flip = now != prev;
            }
            const uint32_t word = __ballot_sync(0xffffffffu, flip);
            if (lane == 0) bm[g * 32 + warp + NWORK * j] = word;  // bit l of word w = channel 32 w + l
            if (flip) { flip_bits |= 1u << (4 * j + g); sign_bits |= (uint32_t)now << (4 * j + g); }
            prev = now;
        }
        sn[c] = (uint8_t)prev;
    }
    bar_workers();
    const int nscan = s.ctr[1];
    if (nscan > SCAN_CAP) return -1;
    if (warp == 0) {
#pragma unroll
        for (int g = 0; g < NGRP; ++g) {
            const int n = __popc(bm[g * 32 + lane]);
            int inc = n;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += t;
            }
            pre[g * 36 + lane] = inc - n;
            if (lane == 31) pre[g * 36 + 32] = inc;
        }
    }
    for (int i = tid; i < nscan; i += NWT) {
        const uint32_t e = s.scan[i];
        const int c = (int)(e & 1023u), g = (int)((e >> 10) & 3u);
        int prev = (int)(e >> 12);
        const float a = cvm[CV_C0 + c], b = xvm[XV_WZ0 + c], p = M ? xvm[XV_WP0 + c] : 0.0f;
        const int r0 = 32 * g > k0 ? 32 * g : k0, r1 = 32 * g + 32 < k1 ? 32 * g + 32 : k1;
        for (int kk = r0; kk < r1; ++kk) {
            const float2 zp = s.zp[kk];
            const int sg = fmaf(p, zp.y, fmaf(b, zp.x, a)) > 0.0f;
            if (sg != prev) push_event(s, kk, c, sg);
            prev = sg;
        }
    }
    bar_workers();
    const int Eg = s.ctr[0];
    const int E = Eg + pre[32] + pre[36 + 32] + pre[72 + 32] + pre[108 + 32];
    if (Eg > EV_CAP || E > EV_CAP) return -1;
    if (warp == 0) {                                             // row offsets: settled events of a group start first, then the scanned ones
        int v[4], x[4], sum = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int k = 4 * lane + q;
            x[q] = 0;
#pragma unroll
            for (int g = 0; g < NGRP; ++g)
                if (ga[g] < gb[g] && k == ga[g]) x[q] = pre[g * 36 + 32];
            v[q] = s.cnt[k] + x[q];
            sum += v[q];
        }
        int inc = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        int run = inc - sum;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            s.start[4 * lane + q] = run;
            s.cnt[4 * lane + q] = run + x[q];                    // first scanned event of the row
            run += v[q];
        }
        if (lane == 31) { s.start[128] = run; s.start[129] = run; }
    }
    bar_workers();
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int g = 0; g < NGRP; ++g)
            if (flip_bits & (1u << (4 * j + g))) {
                const int w = warp + NWORK * j;
                const int pos = s.start[ga[g]] + pre[g * 36 + w] + __popc(bm[g * 32 + w] & ((1u << lane) - 1u));
                s.ev[pos] = ((uint32_t)ga[g] << 11) | ((uint32_t)(tid + NWT * j) << 1) | ((sign_bits >> (4 * j + g)) & 1u);
            }
    for (int i = tid; i < Eg; i += NWT) {
        const uint32_t w = s.stage[i];
        const int k = (int)(w >> 11);
        s.bucket[s.cnt[k] + atomicAdd(&s.fill[k], 1)] = w;
    }
    bar_workers();
    for (int i = tid; i < Eg; i += NWT) {
        const uint32_t w = s.stage[i];
        const int k = (int)(w >> 11);
        const uint32_t c = (w >> 1) & 1023u;
        const int b0 = s.cnt[k], b1 = s.start[k + 1];
        int r = 0;
        for (int j = b0; j < b1; ++j) r += (((s.bucket[j] >> 1) & 1023u) < c) ? 1 : 0;
        s.ev[b0 + r] = w;
    }
    if (tid < EV_PAD) s.ev[E + tid] = EV_SENTINEL;
    bar_workers();
    return E;
}

// layer-1 state of two output channels
struct State {
    float P0, P1, Q0, Q1, R0, R1;
};

// Rows [k0, k1): apply the events of each row to the state, then write leaky(y1pre) of the row.
// Column c of W1 (fp16, 1 KB) is one coalesced 128-byte read per warp; EV_GROUP events are in
// flight per thread while the previous EV_GROUP are applied.
template <int M>
__device__ __forceinline__ void apply_emit(const Smem &s, State &st, const uint32_t *w1h_n, int k0, int k1, int E, int warp, int lane)
{
    const float *c0v = s.cv + M * CV_STRIDE + CV_C0, *wzv = s.xv + M * XV_LR_FLOATS + XV_WZ0, *wpv = s.xv + M * XV_LR_FLOATS + XV_WP0;
    const uint32_t ydst = s.y + warp * A_BLK_BYTES + (lane & 3) * 4;
    const int chunk = lane >> 2;
    int k = k0;
    auto emit = [&](int kk) {
        const float2 zp = s.zp[kk];
        float v0 = fmaf(st.Q0, zp.x, st.P0), v1 = fmaf(st.Q1, zp.x, st.P1);
        if (M) { v0 = fmaf(st.R0, zp.y, v0); v1 = fmaf(st.R1, zp.y, v1); }
        const uint32_t h = leaky_h2(v0, v1);
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(ydst + sw128_off(kk, chunk)), "r"(h) : "memory");
    };
    const int ngroups = (E + EV_GROUP - 1) / EV_GROUP;
    uint32_t wc[EV_GROUP];
    if (ngroups > 0) {
#pragma unroll
        for (int i = 0; i < EV_GROUP; ++i) wc[i] = __ldg(w1h_n + (size_t)((s.ev[i] >> 1) & 1023u) * 256);
    }
#pragma unroll 1
    for (int g = 0; g < ngroups; ++g) {
        uint32_t wn[EV_GROUP];
        if (g + 1 < ngroups) {
#pragma unroll
            for (int i = 0; i < EV_GROUP; ++i) wn[i] = __ldg(w1h_n + (size_t)((s.ev[(g + 1) * EV_GROUP + i] >> 1) & 1023u) * 256);
        }
#pragma unroll
        for (int i = 0; i < EV_GROUP; ++i) {
            const uint32_t e = s.ev[g * EV_GROUP + i];            // sentinels behind the last event
            const int ke = (int)(e >> 11);
            const int kend = ke < k1 ? ke : k1;
            while (k < kend) emit(k++);
            const int c = (int)((e >> 1) & 1023u);
            const float sc = ke >= 128 ? 0.0f : ((e & 1u) ? (1.0f - SURS_LEAKY) : (SURS_LEAKY - 1.0f));
            const float2 w = __half22float2(*reinterpret_cast<const __half2 *>(&wc[i]));
            const float a = c0v[c] * sc, b = wzv[c] * sc;
            st.P0 = fmaf(a, w.x, st.P0); st.P1 = fmaf(a, w.y, st.P1);
            st.Q0 = fmaf(b, w.x, st.Q0); st.Q1 = fmaf(b, w.y, st.Q1);
            if (M) {
                const float p = wpv[c] * sc;
                st.R0 = fmaf(p, w.x, st.R0); st.R1 = fmaf(p, w.y, st.R1);
            }
        }
#pragma unroll
        for (int i = 0; i < EV_GROUP; ++i) wc[i] = wn[i];
    }
    while (k < k1) emit(k++);
}

// 32 channels of one row -> fp16 -> 16-byte chunks chunk0 .. chunk0 + 3 of a K block.
template <int M>
__device__ __forceinline__ void finish32(const uint32_t *acc, const float *add, const float *wz, const float *wp,
                                         float zf, float pred, uint32_t dst, int row, int chunk0)
{
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float v[8];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const float4 a = *reinterpret_cast<const float4 *>(add + 8 * j + 4 * q);
            const float4 z = *reinterpret_cast<const float4 *>(wz + 8 * j + 4 * q);
            v[4 * q + 0] = fmaf(z.x, zf, a.x); v[4 * q + 1] = fmaf(z.y, zf, a.y);
            v[4 * q + 2] = fmaf(z.z, zf, a.z); v[4 * q + 3] = fmaf(z.w, zf, a.w);
            if (M) {
                const float4 p = *reinterpret_cast<const float4 *>(wp + 8 * j + 4 * q);
                v[4 * q + 0] = fmaf(p.x, pred, v[4 * q + 0]); v[4 * q + 1] = fmaf(p.y, pred, v[4 * q + 1]);
                v[4 * q + 2] = fmaf(p.z, pred, v[4 * q + 2]); v[4 * q + 3] = fmaf(p.w, pred, v[4 * q + 3]);
            }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] += __uint_as_float(acc[8 * j + i]);
        const uint4 o = make_uint4(leaky_h2(v[0], v[1]), leaky_h2(v[2], v[3]),
                                   leaky_h2(v[4], v[5]), leaky_h2(v[6], v[7]));
        st_shared_v4(dst + sw128_off(row, chunk0 + j), o);
    }
}

__device__ unsigned long long g_inc_prof[64];

// One MLP over one tile (worker warps).
template <int M, bool PROF>
__device__ __forceinline__ void mlp_pass(const Smem &s, const PointIO &io, const IncParams &prm, State &st, int &cur, uint32_t ph,
                                         uint32_t tmem, int64_t col, int seg, int tid, unsigned long long *prof)
{
    const int warp = tid >> 5, lane = tid & 31;
    const float *cvm = s.cv + M * CV_STRIDE, *xvm = s.xv + M * XV_LR_FLOATS;
    long long t0 = PROF ? clock64() : 0;
    auto lap = [&](int slot) {
        if (PROF && prof) {
            const long long t1 = clock64();
            atomicAdd(prof + slot, (unsigned long long)(t1 - t0));
            t0 = t1;
        }
    };
    // ---- layer 1: events + emission -----------------------------------------------------------
    const uint32_t *w1h_n = reinterpret_cast<const uint32_t *>(prm.w1h[M]) + 32 * warp + lane;
    int k0 = 0, span = TILE_M;
    while (k0 < TILE_M) {
        const int k1 = k0 + span < TILE_M ? k0 + span : TILE_M;
        const int E = detect<M>(s, k0, k1, cur, tid);
        if (E < 0) {                                             // too many events: halve the row range and retry
            span = (k1 - k0) > 1 ? (k1 - k0) / 2 : 1;
            continue;
        }
        lap(1);
        if (PROF && prof) atomicAdd(prof + 8, (unsigned long long)E);
        apply_emit<M>(s, st, w1h_n, k0, k1, E, warp, lane);
        lap(2);
        cur ^= 1;
        k0 = k1;
        span = TILE_M;
    }
    ptx::fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) ptx::mbar_arrive(&s.bars->y1_ready);
    // ---- E2: layer 2 + skip terms -> y2 (K blocks 0-3) ----------------------------------------
    const int quarter = warp & 3, h2 = warp >> 2;
    const int row = quarter * 32 + lane;
    const float2 zp = s.zp[row];
    const uint32_t lane_base = tmem + ((uint32_t)(quarter * 32) << 16);
    ptx::mbar_wait(&s.bars->acc2_full, ph & 1u, 20, prof);
    ptx::tc_fence_after();
    lap(3);
    {
        uint32_t r[2][32];
        const uint32_t taddr = lane_base + 128 * h2;
        ptx::tmem_ld32(taddr, r[0]);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            ptx::tmem_ld_wait();
            if (g < 3) ptx::tmem_ld32(taddr + 32 * (g + 1), r[(g + 1) & 1]);
            const int c = 128 * h2 + 32 * g;
            finish32<M>(r[g & 1], cvm + CV_C2 + c, xvm + XV_WZ2 + c, xvm + XV_WP2 + c, zp.x, zp.y,
                        s.y + (2 * h2 + (g >> 1)) * A_BLK_BYTES, row, (g & 1) * 4);
        }
    }
    ptx::tc_fence_before();
    ptx::fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) ptx::mbar_arrive(&s.bars->y2_ready);
    lap(4);
    // ---- E3: layer 3 + skip terms, layer 4, sigmoid ---------------------------------------------
    ptx::mbar_wait(&s.bars->acc3_full, ph & 1u, 21, prof);
    ptx::tc_fence_after();
    lap(5);
    float partial = 0.0f;
#pragma unroll 1
    for (int g = 0; g < 2; ++g) {
        uint32_t r[32];
        ptx::tmem_ld32(lane_base + 256 + 64 * h2 + 32 * g, r);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int jj = 0; jj < 32; ++jj) {
            const int c = 64 * h2 + 32 * g + jj;
            float v = __uint_as_float(r[jj]) + fmaf(xvm[XV_WZ3 + c], zp.x, cvm[CV_C3 + c]);
            if (M) v = fmaf(xvm[XV_WP3 + c], zp.y, v);
            partial = fmaf(xvm[XV_W4Y + c], leaky(v), partial);
        }
    }
    ptx::tc_fence_before();
    if (h2 == 1) s.part[row] = partial;
    asm volatile("bar.sync %0, 64;" ::"r"(2 + quarter) : "memory");
    if (h2 == 0) {
        float logit = fmaf(xvm[XV_WZ4], zp.x, cvm[CV_C4]);
        if (M) logit = fmaf(xvm[XV_WP4], zp.y, logit);
        logit += partial + s.part[row];
        const float pred = s.mask[row] * (1.0f / (1.0f + expf(-logit)));
        const int k = seg * TILE_M + row;
        if (M == 0) s.zp[row].y = pred;
        if (k < prm.R2) (M == 0 ? io.out_lr : io.out_hr)[col * prm.R2 + k] = pred;
    }
    lap(6);
}

template <bool PROF>
__global__ void __launch_bounds__(NTHREADS, 1) query_inc_kernel(const __grid_constant__ PointIO io, const __grid_constant__ IncParams prm)
{
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = ptx::smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t *smem = smem_raw + (base - raw);
    Smem s;
    s.y = base + SM_Y; s.w = base + SM_W;
    s.cv = reinterpret_cast<const float *>(smem + SM_CV);
    s.xv = reinterpret_cast<const float *>(smem + SM_XV);
    s.ev = reinterpret_cast<uint32_t *>(smem + SM_EV);
    s.scan = reinterpret_cast<uint16_t *>(smem + SM_EV);
    s.stage = reinterpret_cast<uint32_t *>(smem + SM_STAGE);
    s.bucket = reinterpret_cast<uint32_t *>(smem + SM_BUCKET);
    s.cnt = reinterpret_cast<int *>(smem + SM_CNT);
    s.fill = s.cnt + 132; s.start = s.cnt + 264; s.ctr = s.cnt + 396;
    s.sgn = smem + SM_SGN;
    s.zp = reinterpret_cast<float2 *>(smem + SM_ZP);
    s.mask = reinterpret_cast<float *>(smem + SM_MASK);
    s.part = s.mask + 128;
    Bars *bars = reinterpret_cast<Bars *>(smem + SM_BAR);
    s.bars = bars;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned long long *prof = nullptr;
    if (PROF && lane == 0 && (warp == 0 || warp >= NWORK)) prof = g_inc_prof;
    const long long t_kernel0 = PROF ? clock64() : 0;

    if (threadIdx.x == 0) {
        for (int i = 0; i < NSTAGE; ++i) { ptx::mbar_init(&bars->full_w[i], 1); ptx::mbar_init(&bars->empty_w[i], 1); }
        ptx::mbar_init(&bars->y1_ready, NWORK); ptx::mbar_init(&bars->y2_ready, NWORK);
        ptx::mbar_init(&bars->acc2_full, 1); ptx::mbar_init(&bars->acc3_full, 1);
        ptx::mbar_init(&bars->cv_full, 1); ptx::mbar_init(&bars->cv_empty, NWORK);
        ptx::fence_barrier_init();
    }
    if (warp == NWORK) ptx::tmem_alloc(&bars->tmem_base, 512);
    for (int i = threadIdx.x; i < XV_BYTES / 16; i += NTHREADS)          // constant vectors: resident for the whole kernel
        reinterpret_cast<uint4 *>(smem + SM_XV)[i] = __ldg(reinterpret_cast<const uint4 *>(prm.xv) + i);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = bars->tmem_base;

    if (warp < NWORK) {
        // =============================== workers ===========================================
        const int tid = threadIdx.x;
        const int n0 = 64 * warp + 2 * lane;
        State st[2];
        int cur[2] = {0, 0};
        uint32_t ph = 0, colit = 0;
        for (int64_t col = blockIdx.x; col < prm.ncols; col += gridDim.x, ++colit) {
            ptx::mbar_wait(&bars->cv_full, colit & 1u, 11, prof);
            // every column starts from the all-negative state
            {
                const float2 c1l = __ldg(reinterpret_cast<const float2 *>(prm.table + col * CV_ROW_FLOATS + CV_C1 + n0));
                const float2 c1h = __ldg(reinterpret_cast<const float2 *>(prm.table + col * CV_ROW_FLOATS + CV_C1 + 512 + n0));
                const float2 ql = __ldg(reinterpret_cast<const float2 *>(prm.qstar + n0)), qh = __ldg(reinterpret_cast<const float2 *>(prm.qstar + 512 + n0));
                const float2 rh = __ldg(reinterpret_cast<const float2 *>(prm.rstar + 512 + n0));
                st[0].P0 = c1l.x; st[0].P1 = c1l.y; st[0].Q0 = ql.x; st[0].Q1 = ql.y; st[0].R0 = 0.f; st[0].R1 = 0.f;
                st[1].P0 = c1h.x; st[1].P1 = c1h.y; st[1].Q0 = qh.x; st[1].Q1 = qh.y; st[1].R0 = rh.x; st[1].R1 = rh.y;
                cur[0] = cur[1] = 0;
            }
            const int i = prm.plane_lo + (int)(col / prm.R1), j = (int)(col % prm.R1);
            for (int seg = 0; seg < prm.nseg; ++seg) {
                bar_workers();                                   // previous tile is done with zp / mask / sgn
                if (tid < TILE_M) {
                    const int k = seg * TILE_M + tid;
                    const int kc = k < prm.R2 ? k : prm.R2 - 1;
                    const Projected pr = project_point(io, (float)io.axis[0][i], (float)io.axis[1][j], (float)io.axis[2][kc]);
                    s.zp[tid] = make_float2(pr.zf, 0.0f);
                    s.mask[tid] = pr.mask;
                }
                if (seg == 0) reinterpret_cast<uint4 *>(s.sgn)[tid] = make_uint4(0, 0, 0, 0);
                // (detect() starts with a barrier)
                mlp_pass<0, PROF>(s, io, prm, st[0], cur[0], ph, tmem, col, seg, tid, prof);
                ++ph;
                mlp_pass<1, PROF>(s, io, prm, st[1], cur[1], ph, tmem, col, seg, tid, prof);
                ++ph;
            }
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&bars->cv_empty);
        }
    } else if (warp == NWORK) {
        // =============================== MMA issue ========================================
        if (lane == 0) {
            constexpr uint32_t IDESC128 = ptx::umma_idesc_f16(128, 128);
            uint32_t wblk = 0, ph = 0;
            auto wait_w = [&]() -> uint32_t {
                const uint32_t sl = wblk % NSTAGE;
                ptx::mbar_wait(&bars->full_w[sl], (wblk / NSTAGE) & 1u, 30, prof);
                ptx::tc_fence_after();
                return s.w + sl * W128_BLK_BYTES;
            };
            auto release_w = [&]() {
                ptx::umma_commit(&bars->empty_w[wblk % NSTAGE]);
                ++wblk;
            };
            const int64_t ncol_mine = prm.ncols > blockIdx.x ? (prm.ncols - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
            const int64_t npass = ncol_mine * prm.nseg * 2;
            for (int64_t it = 0; it < npass; ++it, ++ph) {
                ptx::mbar_wait(&bars->y1_ready, ph & 1u, 31, prof);
                ptx::tc_fence_after();
                for (int kb = 0; kb < 8; ++kb)
                    for (int h = 0; h < 2; ++h) {
                        const uint32_t w = wait_w();
                        mma_block(tmem + 128 * h, s.y + kb * A_BLK_BYTES, w, 4, IDESC128, kb == 0);
                        release_w();
                    }
                ptx::umma_commit(&bars->acc2_full);
                ptx::mbar_wait(&bars->y2_ready, ph & 1u, 32, prof);
                ptx::tc_fence_after();
                for (int kb = 0; kb < 4; ++kb) {
                    const uint32_t w = wait_w();
                    mma_block(tmem + 256, s.y + kb * A_BLK_BYTES, w, 4, IDESC128, kb == 0);
                    release_w();
                }
                ptx::umma_commit(&bars->acc3_full);
            }
        }
    } else {
        // =============================== weight stream + column vectors ====================
        if (lane == 0) {
            uint32_t wblk = 0, colit = 0;
            for (int64_t col = blockIdx.x; col < prm.ncols; col += gridDim.x, ++colit) {
                ptx::mbar_wait(&bars->cv_empty, (colit & 1u) ^ 1u, 41, prof);
                ptx::mbar_arrive_expect_tx(&bars->cv_full, CV_BYTES);
                ptx::tma_load_1d(smem + SM_CV, prm.table + col * CV_ROW_FLOATS, CV_BYTES, &bars->cv_full);
                for (int seg = 0; seg < prm.nseg; ++seg) {
                    const uint8_t *src = prm.weights;
                    for (int b = 0; b < 2 * XW_BLOCKS_PER_MLP; ++b) {
                        const uint32_t sl = wblk % NSTAGE;
                        ptx::mbar_wait(&bars->empty_w[sl], ((wblk / NSTAGE) & 1u) ^ 1u, 40, prof);
                        ptx::mbar_arrive_expect_tx(&bars->full_w[sl], W128_BLK_BYTES);
                        ptx::tma_load_1d(smem + SM_W + sl * W128_BLK_BYTES, src, W128_BLK_BYTES, &bars->full_w[sl]);
                        src += W128_BLK_BYTES;
                        ++wblk;
                    }
                }
            }
        }
    }
    if (PROF && threadIdx.x == 0) atomicAdd(g_inc_prof + 0, (unsigned long long)(clock64() - t_kernel0));
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == NWORK) ptx::tmem_dealloc(tmem, 512);
}

}  // namespace

int surs_launch_query_inc(surs_ctx *ctx, const PointIO &io, int R1, int R2, int plane_lo, int nplanes, cudaStream_t st)
{
    const int64_t ncols = (int64_t)nplanes * R1;
    if (ncols <= 0) return 0;
    if (surs_col_build_table(ctx, io, R1, plane_lo, ncols, st)) return 1;
    uint8_t *base = (uint8_t *)ctx->col_weights;
    IncParams prm;
    prm.weights = base + OFF_XW;
    prm.xv = reinterpret_cast<const float *>(base + OFF_XV);
    prm.table = (const float *)ctx->col_table;
    prm.w1h[0] = reinterpret_cast<const __half *>(base + OFF_W1H); prm.w1h[1] = prm.w1h[0] + 1024 * 512;
    prm.qstar = reinterpret_cast<const float *>(base + OFF_QSTAR);
    prm.rstar = reinterpret_cast<const float *>(base + OFF_RSTAR);
    prm.ncols = ncols;
    prm.nseg = (R2 + TILE_M - 1) / TILE_M;
    prm.R1 = R1; prm.R2 = R2; prm.plane_lo = plane_lo;
    const int grid = (int)(ncols < ctx->sm_count ? ncols : ctx->sm_count);
    static const bool profile = getenv("SURS_TC_PROFILE") != nullptr;
    if (!profile) {
        SURS_CUDA(ctx, cudaFuncSetAttribute(query_inc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL));
        query_inc_kernel<false><<<grid, NTHREADS, SM_TOTAL, st>>>(io, prm);
        SURS_LAUNCH_CHECK(ctx, "query_inc_kernel");
        return 0;
    }
    unsigned long long zero[64] = {0}, h[64];
    SURS_CUDA(ctx, cudaMemcpyToSymbol(g_inc_prof, zero, sizeof(zero)));
    SURS_CUDA(ctx, cudaFuncSetAttribute(query_inc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL));
    query_inc_kernel<true><<<grid, NTHREADS, SM_TOTAL, st>>>(io, prm);
    SURS_LAUNCH_CHECK(ctx, "query_inc_kernel<profile>");
    SURS_CUDA(ctx, cudaStreamSynchronize(st));
    SURS_CUDA(ctx, cudaMemcpyFromSymbol(h, g_inc_prof, sizeof(h)));
    const double npass = (double)ncols * prm.nseg * 2;
    const double k = 1e-3 / npass;
    fprintf(stderr, "[surs inc profile] tile-MLP passes=%.0f grid=%d events/pass %.1f | kcycles/pass: total %.2f | warp0: detect %.2f emit %.2f wait_L2 %.2f E2 %.2f wait_L3 %.2f E3 %.2f "
                    "| mma: wait_w %.2f wait_y1 %.2f wait_y2 %.2f | loader wait_empty %.2f wait_cv_empty %.2f\n",
            npass, grid, (double)h[8] / npass, h[0] * k * 1.0, h[1] * k, h[2] * k, h[3] * k, h[4] * k, h[5] * k, h[6] * k,
            h[30] * k, h[31] * k, h[32] * k, h[40] * k, h[41] * k);
    return 0;
}
