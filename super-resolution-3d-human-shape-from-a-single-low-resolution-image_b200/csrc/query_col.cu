// Column-factored grid evaluation (SURS_PREC_FP16 and SURS_PREC_FP16X3): surs_eval_grid and the levels of
// surs_eval_grid_octree without a transform; in SURS_PREC_FP16X3 also every other point source, through per-point
// tables (MODE 2 below).
//
// With the calibration used by gen_mesh (lib/train_util.py:63-66: diag(2,-2,2,1)) -- in general
// whenever calib[0][2] == calib[1][2] == 0 and the grid is not transformed -- the image
// coordinates (u,v) of a grid node depend only on its (i,j) and all 320 gathered feature
// channels are identical along a k-column (lib/sdf.py:17-24 + lib/geometry.py:15-31).  Every
// product of a weight matrix with the image features is therefore computed ONCE PER COLUMN:
//
//   table kernel  (col_table_kernel): for each column, C0 = W0[:, :320] f + b0 (1024),
//       C2 = W2[:, 512:832] f + b2 (256), C3 = W3[:, 256:576] f + b3 (128),
//       C4 = W4[:, 128:448] f + b4 (1), for both MLPs -- a tcgen05 GEMM over 128-column tiles.
//   main kernel   (query_col_kernel): tile = 128 consecutive k of one column.  Layer 0 collapses
//       to y0 = leaky(C0 + w_z z_feat (+ w_p pred_lr)) on the CUDA cores; layers 1-3 run on the
//       tensor cores exactly as in query_tc.cu but without layer-0 / skip GEMMs:
//       L1 (K=1024, N=512: both TMEM accumulators), L2 (K=512, N=256), L3 (K=256, N=128);
//       the skip terms enter the epilogues as C2 / C3 / C4 + rank-1 updates in z_feat and pred_lr.
//
// Identical arithmetic up to rounding order (it is an exact refactoring of W.[y; f; z; p]),
// executed MACs per point drop from 3.0 M (query_tc.cu) to 1.38 M and the weight stream per tile
// from 6.7 MB to 2.7 MB.  The algorithmic FLOP count used for roofline reporting stays
// 4 564 998 per point.
//
// Warp roles (main kernel): 0-7 produce y0 and run the epilogues (warp w: TMEM lanes 32 (w%4)..,
// column half w/4), 8 = MMA issue (T0 half of layer 1, layers 2 and 3), 9 = weight stream + per-column vectors
// (bulk TMA), 10 = MMA issue of the T1 half of layer 1 (one-pass mode only).
#include "col_epi.cuh"

#include <new>
#include <stdlib.h>

namespace {

using namespace col;

// ring depths: P = 1 (one pass) 4 weight stages + 3 A slots; P = 3 (split operands) 3 + 4: a K block is a
// (hi, lo) slot pair there, so four slots hold two K blocks
// (measured at 512^3: 4 weight stages + 3 A slots 342 ms; 3 + 4 in the same shared memory 352 ms)
#ifndef SURS_P1_NSTAGE
#define SURS_P1_NSTAGE 4
#define SURS_P1_NASLOT 3
#endif
template <int P> struct Ring {
    static constexpr int NSTAGE = P == 1 ? SURS_P1_NSTAGE : 3;
    static constexpr int NA_SLOT = P == 1 ? SURS_P1_NASLOT : 4;
    static constexpr int AP = P == 1 ? 1 : 2;            // A blocks per K block (hi | hi, lo)
    static constexpr int WP = P == 1 ? 1 : 2;            // weight blocks per K block (hi | hi, lo)
    static constexpr int SMEM_W = 0;
    static constexpr int SMEM_A = SMEM_W + NSTAGE * W_BLK_BYTES;
    static constexpr int SMEM_CV = SMEM_A + NA_SLOT * A_BLK_BYTES;
    static constexpr int SMEM_GV = SMEM_CV + 2 * CV_BYTES;
    static constexpr int SMEM_BAR = SMEM_GV + GV_BYTES;
    static constexpr int SMEM_PREDX = SMEM_BAR + 256;   // layer-4 partial sums and pred_lr hand-over between the warps
    static constexpr int SMEM_TOTAL = SMEM_PREDX + 512 + 1024;
    static_assert(SMEM_TOTAL <= 232448, "shared memory budget");
};
constexpr int BLOCKS_PER_MLP = 32 + 8 + 4;
constexpr int NEPI = 8;
constexpr int NTHREADS = (NEPI + 3) * 32;            // + MMA issue, weight stream, second MMA issue

template <int NSTAGE, int NA_SLOT>
struct BarsT {
    uint64_t full_w[NSTAGE], empty_w[NSTAGE];
    uint64_t a_ready[NA_SLOT], a_free[NA_SLOT];
    uint64_t acc_full[2], acc_free[2];
    uint64_t cv_full[2], cv_empty[2];
    // second MMA-issuing thread (T1 half of layer 1): it only ever sees its own fills, so that
    // its parity waits stay exactly one phase behind (a thread that skips fills loses the count)
    uint64_t full_wb[NSTAGE], a_ready_b[NA_SLOT], t1_free_b;
    uint32_t tmem_base;
};
static_assert(sizeof(BarsT<4, 3>) <= 256 && sizeof(BarsT<3, 4>) <= 256, "barrier block");

// Indexed mode (octree lists): the four rows of a lane belong to arbitrary columns, so each row brings its
// own C0 from the column table (global memory; rows of one column share the cache line).
struct C0Rows {
    float4 v[4][2];
};
__device__ __forceinline__ void load_c0_rows(C0Rows &c, const float *const (&trow)[4], int off)
{
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        c.v[r][0] = __ldg(reinterpret_cast<const float4 *>(trow[r] + off));
        c.v[r][1] = __ldg(reinterpret_cast<const float4 *>(trow[r] + off + 4));
    }
}
template <int P, bool HAS_P>
__device__ __forceinline__ void produce8x(const C0Rows &c0, const float *wz, const float *wp, const float (&zf)[4], const float (&pred)[4],
                                          uint32_t dst, int lane, int chunk, int part)
{
    uint64_t z[4], p[4];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const float4 zv = *reinterpret_cast<const float4 *>(wz + 4 * q);
        z[2 * q] = pk2(zv.x, zv.y); z[2 * q + 1] = pk2(zv.z, zv.w);
        if (HAS_P) {
            const float4 pv = *reinterpret_cast<const float4 *>(wp + 4 * q);
            p[2 * q] = pk2(pv.x, pv.y); p[2 * q + 1] = pk2(pv.z, pv.w);
        }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const uint64_t a[4] = {pk2(c0.v[r][0].x, c0.v[r][0].y), pk2(c0.v[r][0].z, c0.v[r][0].w), pk2(c0.v[r][1].x, c0.v[r][1].y), pk2(c0.v[r][1].z, c0.v[r][1].w)};
        const uint64_t zz = pk2(zf[r], zf[r]), pp = pk2(pred[r], pred[r]);
        float f[8];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            uint64_t v = ffma2(z[i], zz, a[i]);
            if (HAS_P) v = ffma2(p[i], pp, v);
            up2(v, f[2 * i], f[2 * i + 1]);
        }
        const uint4 o = make_uint4(act_h2<P>(f[0], f[1], part), act_h2<P>(f[2], f[3], part),
                                   act_h2<P>(f[4], f[5], part), act_h2<P>(f[6], f[7], part));
        st_shared_v4(dst + sw128_off(lane + 32 * r, chunk), o);
    }
}

struct EpiCtx {
    uint64_t *a_ready, *a_ready_b, *a_free, *acc_free;   // barrier arrays of the kernel's BarsT
    uint32_t nslot;
    uint32_t a_smem;
    int row, hsel, lane;
    float zf, pred;
    uint32_t g;                    // running A-ring K block number
    int ablate;                    // profiling builds only (SURS_COL_ABLATE): 8 no proxy fence, 16 no epilogue math / stores
    unsigned long long *prof;
};

__device__ __forceinline__ uint32_t ring_acquire(EpiCtx &e)
{
    const uint32_t slot = e.g % e.nslot;
    ptx::mbar_wait(&e.a_free[slot], ((e.g / e.nslot) & 1u) ^ 1u, 10, e.prof);
    return slot;
}
__device__ __forceinline__ void ring_publish(EpiCtx &e, uint32_t slot, bool layer0 = false)   // layer0: the block feeds the second issuing thread too
{
    if (!(e.ablate & 8)) ptx::fence_proxy_async_smem();
    __syncwarp();
    if (e.lane == 0) {
        ptx::mbar_arrive(&e.a_ready[slot]);
        if (layer0) ptx::mbar_arrive(&e.a_ready_b[slot]);          // layer-0 blocks feed both halves of layer 1
    }
    ++e.g;
}

// epilogue of a 256-column accumulator into 4 K blocks of the A ring; the accumulator is handed
// back (acc_free) as soon as its last column has been read, before the last block is written
template <int P, bool HAS_Z, bool HAS_P>
__device__ __forceinline__ void epilogue_256(EpiCtx &e, uint32_t taddr, int acc_id, const float *add, const float *wz, const float *wp)
{
    uint32_t r[2][32];
    ptx::tmem_ld32(taddr + e.hsel * 32, r[0]);
#pragma unroll
    for (int kb = 0; kb < 4; ++kb) {
        const long long t_a = e.prof ? clock64() : 0;
        ptx::tmem_ld_wait();
        if (e.prof) atomicAdd(e.prof + 24, (unsigned long long)(clock64() - t_a));     // wait for the TMEM load
        if (kb < 3) {
            ptx::tmem_ld32(taddr + (kb + 1) * 64 + e.hsel * 32, r[(kb + 1) & 1]);
        } else {
            ptx::tc_fence_before();
            __syncwarp();
            if (e.lane == 0) ptx::mbar_arrive(&e.acc_free[acc_id]);
        }
        const int c = kb * 64 + e.hsel * 32;
#pragma unroll 1
        for (int part = 0; part < Ring<P>::AP; ++part) {       // P = 3: the K block goes out as a (hi, lo) slot pair
            const uint32_t slot = ring_acquire(e);
            const long long t_b = e.prof ? clock64() : 0;
            if (!(e.ablate & 16))
                finish32<P, true, HAS_Z, HAS_P>(r[kb & 1], add + c, wz + c, wp + c, e.zf, e.pred, e.a_smem + slot * A_BLK_BYTES, e.row, e.hsel, part);
            const long long t_c = e.prof ? clock64() : 0;
            ring_publish(e, slot);
            if (e.prof) {
                atomicAdd(e.prof + 25, (unsigned long long)(t_c - t_b));                   // bias / skip terms, leaky, fp16, stores
                atomicAdd(e.prof + 26, (unsigned long long)(clock64() - t_c));             // fence + arrive
                atomicAdd(e.prof + 27, 1ull);
            }
        }
    }
}

__device__ unsigned long long g_col_prof[64];

// INDEXED: the tile is 128 consecutive entries of io.idx_list (octree levels) instead of 128 consecutive k of
// one column; column vectors are read per row from the table of ALL columns (plane_lo = 0), results are
// scattered with pointio_store.
// P = 3 (SURS_PREC_FP16X3): every K block of layers 1-3 is issued three times -- A_hi.W_hi, A_lo.W_hi, A_hi.W_lo --
// from an A ring that carries the block as a (hi, lo) slot pair and a weight stream that carries it as (hi, lo).
// The schedule differs from P = 1: an accumulator may only be released after its last column has been read, and
// with two ring slots per K block E1 can no longer park three K blocks in the ring while layer 2 waits for the
// accumulator it is reading.  So layer 1 runs as two sequential N halves into T0 (layer 0 is produced twice), layer 2
// accumulates into T1 behind each half, layer 3 goes to T0 -- the accumulator being drained is never the one the
// consumer of the drained blocks writes to.  One issuing thread; warp NEPI + 2 idles.
// MODE 0: dense slab; 1 (octree levels): rows = entries of io.idx_list, table rows = grid columns; 2 (arbitrary
// point sources, SURS_PREC_FP16X3): rows = points prm.n0 + ..., the "column" table has one row PER POINT
// (col_table_kernel in point mode), so nothing is shared between rows but the kernels and their arithmetic are reused.
template <bool PROF, int MODE, int P>
__global__ void __launch_bounds__(NTHREADS, 1) query_col_kernel(const __grid_constant__ PointIO io, const __grid_constant__ ColParams prm)
{
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = ptx::smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t *smem = smem_raw + (base - raw);
    constexpr bool INDEXED = MODE != 0;
    using RG = Ring<P>;
    constexpr int NSTAGE = RG::NSTAGE, NA_SLOT = RG::NA_SLOT, AP = RG::AP;
    constexpr int SMEM_W = RG::SMEM_W, SMEM_A = RG::SMEM_A, SMEM_CV = RG::SMEM_CV, SMEM_GV = RG::SMEM_GV;
    using Bars = BarsT<NSTAGE, NA_SLOT>;
    const uint32_t a_smem = base + SMEM_A, w_smem = base + SMEM_W;
    float *cv_s = reinterpret_cast<float *>(smem + SMEM_CV);
    float *gv_s = reinterpret_cast<float *>(smem + SMEM_GV);
    Bars *bars = reinterpret_cast<Bars *>(smem + RG::SMEM_BAR);
    float *pred_x = reinterpret_cast<float *>(smem + RG::SMEM_PREDX);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned long long *prof = nullptr;
    if (PROF && lane == 0 && (warp == 0 || warp == NEPI || warp == NEPI + 1)) prof = g_col_prof;
    const long long t_kernel0 = PROF ? clock64() : 0;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NSTAGE; ++s) { ptx::mbar_init(&bars->full_w[s], 1); ptx::mbar_init(&bars->empty_w[s], 1); }
        for (int k = 0; k < NA_SLOT; ++k) {
            ptx::mbar_init(&bars->a_ready[k], NEPI); ptx::mbar_init(&bars->a_ready_b[k], NEPI);
            ptx::mbar_init(&bars->a_free[k], P == 1 ? 2 : 1);          // P = 1: one commit from each issuing thread
        }
        for (int s = 0; s < NSTAGE; ++s) ptx::mbar_init(&bars->full_wb[s], 1);
        ptx::mbar_init(&bars->t1_free_b, NEPI);
        for (int t = 0; t < 2; ++t) {
            ptx::mbar_init(&bars->acc_full[t], 1); ptx::mbar_init(&bars->acc_free[t], NEPI);
            ptx::mbar_init(&bars->cv_full[t], 1); ptx::mbar_init(&bars->cv_empty[t], NEPI);
        }
        ptx::fence_barrier_init();
    }
    if (warp == NEPI) ptx::tmem_alloc(&bars->tmem_base, 512);
    for (int i = threadIdx.x; i < GV_BYTES / 16; i += NTHREADS)          // constant vectors: resident for the whole kernel
        reinterpret_cast<uint4 *>(gv_s)[i] = __ldg(reinterpret_cast<const uint4 *>(prm.gv) + i);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = bars->tmem_base;
    const uint32_t T0 = tmem, T1 = tmem + 256;

    if (warp < NEPI) {
        // =============================== y0 production + epilogues ========================
        EpiCtx e;
        e.a_ready = bars->a_ready; e.a_ready_b = bars->a_ready_b; e.a_free = bars->a_free; e.acc_free = bars->acc_free;
        e.nslot = NA_SLOT; e.a_smem = a_smem; e.lane = lane; e.prof = prof; e.g = 0;
        e.ablate = PROF ? prm.ablate : 0;
        const int quarter = warp & 3;
        e.hsel = warp >> 2;
        e.row = quarter * 32 + lane;
        const uint32_t lane_t0 = T0 + ((uint32_t)(quarter * 32) << 16), lane_t1 = T1 + ((uint32_t)(quarter * 32) << 16);
        uint32_t acc0 = 0, acc1 = 0, it = 0;
        for (int64_t tile = blockIdx.x; tile < prm.ntiles; tile += gridDim.x, ++it) {
            int64_t col = 0, n_own = 0;
            int k = 0;
            Projected pr;
            float zf4[4], pred4[4] = {0.0f, 0.0f, 0.0f, 0.0f};      // layer 0 works on rows lane + 32 r
            const float *trow4[4] = {nullptr, nullptr, nullptr, nullptr};
            const float *cv = nullptr;
            uint32_t cvb = 0;
            if (!INDEXED) {
                col = tile / prm.nseg;
                const int seg = (int)(tile - col * prm.nseg);
                k = seg * TILE_M + e.row;
                const int kc = k < prm.R2 ? k : prm.R2 - 1;
                const int i = prm.plane_lo + (int)(col / prm.R1), j = (int)(col % prm.R1);
                pr = project_point(io, (float)io.axis[0][i], (float)io.axis[1][j], (float)io.axis[2][kc]);
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const int kr = seg * TILE_M + lane + 32 * r;
                    zf4[r] = project_point(io, (float)io.axis[0][i], (float)io.axis[1][j], (float)io.axis[2][kr < prm.R2 ? kr : prm.R2 - 1]).zf;
                }
                cvb = it & 1u;
                ptx::mbar_wait(&bars->cv_full[cvb], (it >> 1) & 1u, 11, prof);
                cv = cv_s + cvb * CV_FLOATS;
            } else {
                auto node = [&](int64_t n, int64_t &c, Projected &q) {
                    if (MODE == 2) {
                        const int64_t nn = n < prm.n_end ? n : prm.n_end - 1;
                        float x, y, z;
                        pointio_load(io, nn, x, y, z);
                        q = project_point(io, x, y, z);
                        c = nn - prm.n0;
                        return;
                    }
                    const int64_t lin = io.idx_list[n < io.n ? n : io.n - 1];
                    c = lin / prm.R2;
                    const int kk = (int)(lin - c * prm.R2), jj = (int)(c % prm.R1), ii = (int)(c / prm.R1);
                    q = project_point(io, (float)io.axis[0][ii], (float)io.axis[1][jj], (float)io.axis[2][kk]);
                };
                const int64_t nbase = MODE == 2 ? prm.n0 : 0;
                n_own = nbase + tile * TILE_M + e.row;
                node(n_own, col, pr);
                if (MODE == 1) col -= prm.col0;
                cv = prm.table + col * CV_ROW_FLOATS;               // this row's column vectors, in global memory
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    int64_t c;
                    Projected q;
                    node(nbase + tile * TILE_M + lane + 32 * r, c, q);
                    zf4[r] = q.zf;
                    if (MODE == 1) c -= prm.col0;
                    trow4[r] = prm.table + c * CV_ROW_FLOATS;
                }
            }
            e.zf = pr.zf;
            float pred_lr = 0.0f;
#pragma unroll 1
            for (int m = 0; m < prm.nmlp; ++m) {
                const float *cvm = cv + m * CV_STRIDE, *gvm = gv_s + m * GV_STRIDE;
                e.pred = pred_lr;
                // layer 0 on the CUDA cores: 16 K blocks of y0 = leaky(C0 + w_z z (+ w_p pred_lr))
                C0Rows c0rows;
#pragma unroll 1
                for (int half = 0; half < (P == 1 ? 1 : 2); ++half) {
                if (INDEXED) load_c0_rows(c0rows, trow4, m * CV_STRIDE + CV_C0 + warp * 8);
#pragma unroll 1
                for (int kp = 0; kp < 16 * AP; ++kp) {
                    const int kb = kp / AP, part = kp - kb * AP;
                    const uint32_t slot = ring_acquire(e);
                    const int c = kb * 64 + warp * 8;                    // warp w fills 16-byte chunk w of all 128 rows
                    const uint32_t dst = a_smem + slot * A_BLK_BYTES;
                    if (PROF && (prm.ablate & 2)) {
                        // profiling only: no layer-0 arithmetic (the A block keeps stale data)
                    } else if (INDEXED) {
                        if (m == 0) produce8x<P, false>(c0rows, gvm + GV_WZ0 + c, nullptr, zf4, pred4, dst, lane, warp, part);
                        else produce8x<P, true>(c0rows, gvm + GV_WZ0 + c, gvm + GV_WP0 + c, zf4, pred4, dst, lane, warp, part);
                        if (kb < 15 && part == AP - 1) load_c0_rows(c0rows, trow4, m * CV_STRIDE + CV_C0 + c + 64);   // lands while the next slot is awaited
                    } else if (m == 0) produce8<P, false>(cvm + CV_C0 + c, gvm + GV_WZ0 + c, nullptr, zf4, pred4, dst, lane, warp, part);
                    else produce8<P, true>(cvm + CV_C0 + c, gvm + GV_WZ0 + c, gvm + GV_WP0 + c, zf4, pred4, dst, lane, warp, part);
                    ring_publish(e, slot, P == 1);
                }
                if (P != 1) {
                    // E1 of this half of layer 1 (always T0) -> A ring; layer 2 accumulates into T1 behind it
                    ptx::mbar_wait(&bars->acc_full[0], acc0 & 1u, 20, prof);
                    ptx::tc_fence_after();
                    epilogue_256<P, false, false>(e, lane_t0, 0, gvm + GV_B1 + half * 256, nullptr, nullptr);
                    ++acc0;
                }
                }
                // P = 1: layer 1 in T0 | T1, layer 2 -> T0, layer 3 -> T1;  P = 3: layer 2 -> T1, layer 3 -> T0
                const uint32_t lane_l2 = P == 1 ? lane_t0 : lane_t1, lane_l3 = P == 1 ? lane_t1 : lane_t0;
                uint32_t &acc_l2 = P == 1 ? acc0 : acc1, &acc_l3 = P == 1 ? acc1 : acc0;
                constexpr int ID_L2 = P == 1 ? 0 : 1, ID_L3 = P == 1 ? 1 : 0;
                if (P == 1) {
                    // E1: layer 1, both halves (bias b1) -> A ring
                    ptx::mbar_wait(&bars->acc_full[0], acc0 & 1u, 20, prof);
                    ptx::tc_fence_after();
                    epilogue_256<P, false, false>(e, lane_t0, 0, gvm + GV_B1, nullptr, nullptr);
                    ++acc0;
                    ptx::mbar_wait(&bars->acc_full[1], acc1 & 1u, 21, prof);
                    ptx::tc_fence_after();
                    epilogue_256<P, false, false>(e, lane_t1, 1, gvm + GV_B1 + 256, nullptr, nullptr);
                    ++acc1;
                }
                // E2: layer 2 + skip terms -> A ring
                ptx::mbar_wait(&bars->acc_full[ID_L2], acc_l2 & 1u, 22, prof);
                ptx::tc_fence_after();
                if (m == 0) epilogue_256<P, true, false>(e, lane_l2, ID_L2, cvm + CV_C2, gvm + GV_WZ2, nullptr);
                else epilogue_256<P, true, true>(e, lane_l2, ID_L2, cvm + CV_C2, gvm + GV_WZ2, gvm + GV_WP2);
                ++acc_l2;
                // E3: layer 3 + skip terms, layer 4 (each warp 64 of the 128 channels, four partial sums), sigmoid (warps 0-3)
                ptx::mbar_wait(&bars->acc_full[ID_L3], acc_l3 & 1u, 23, prof);
                ptx::tc_fence_after();
                float logit = 0.0f;
                {
                    uint32_t r[2][32];
                    const int cb = e.hsel * 64;
                    ptx::tmem_ld32(lane_l3 + cb, r[0]);
                    ptx::tmem_ld32(lane_l3 + cb + 32, r[1]);
                    ptx::tmem_ld_wait();
                    // packed fp32 pairs: v = ((acc + C3) + wz3 z) (+ wp3 pred), partial sums (lg0, lg1) and (lg2, lg3)
                    uint64_t lg01 = pk2(0.0f, 0.0f), lg23 = lg01;
                    const uint64_t zz = pk2(e.zf, e.zf), pp = pk2(e.pred, e.pred), slope = pk2(SURS_LEAKY, SURS_LEAKY);
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
#pragma unroll
                        for (int j4 = 0; j4 < 8; ++j4) {
                            const int c = cb + q * 32 + 4 * j4;
                            const float4 a = *reinterpret_cast<const float4 *>(cvm + CV_C3 + c), z = *reinterpret_cast<const float4 *>(gvm + GV_WZ3 + c);
                            const float4 w4 = *reinterpret_cast<const float4 *>(gvm + GV_W4Y + c);
                            uint64_t v01 = fadd2(pk2(__uint_as_float(r[q][4 * j4]), __uint_as_float(r[q][4 * j4 + 1])), pk2(a.x, a.y));
                            uint64_t v23 = fadd2(pk2(__uint_as_float(r[q][4 * j4 + 2]), __uint_as_float(r[q][4 * j4 + 3])), pk2(a.z, a.w));
                            v01 = ffma2(pk2(z.x, z.y), zz, v01);
                            v23 = ffma2(pk2(z.z, z.w), zz, v23);
                            if (m == 1) {
                                const float4 p = *reinterpret_cast<const float4 *>(gvm + GV_WP3 + c);
                                v01 = ffma2(pk2(p.x, p.y), pp, v01);
                                v23 = ffma2(pk2(p.z, p.w), pp, v23);
                            }
                            float v0, v1, v2, v3, s0, s1, s2, s3;
                            up2(v01, v0, v1); up2(v23, v2, v3);
                            up2(fmul2(v01, slope), s0, s1); up2(fmul2(v23, slope), s2, s3);
                            lg01 = ffma2(pk2(w4.x, w4.y), pk2(fmaxf(v0, s0), fmaxf(v1, s1)), lg01);
                            lg23 = ffma2(pk2(w4.z, w4.w), pk2(fmaxf(v2, s2), fmaxf(v3, s3)), lg23);
                        }
                    }
                    float lg[4];
                    up2(lg01, lg[0], lg[1]); up2(lg23, lg[2], lg[3]);
                    logit = (lg[0] + lg[1]) + (lg[2] + lg[3]);
                }
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    ptx::mbar_arrive(&bars->acc_free[ID_L3]);
                    if (P == 1) ptx::mbar_arrive(&bars->t1_free_b);
                }
                ++acc_l3;
                // the upper half's partial sum goes to the quarter's lower warp (fixed order: lower + upper) through the
                // row's pred_x cell: its last readers (the LR -> HR hand-over) are a whole pass of ring traffic behind
                if (e.hsel == 1) pred_x[e.row] = logit;
                asm volatile("bar.sync %0, 64;" ::"r"(2 + quarter) : "memory");
                if (e.hsel == 0) logit = (cvm[CV_C4] + gvm[GV_WZ4] * e.zf + (m == 1 ? gvm[GV_WP4] * e.pred : 0.0f)) + (logit + pred_x[e.row]);
                if (e.hsel == 0) {
                    const float pred = pr.mask * (1.0f / (1.0f + expf(-logit)));
                    if (m == 0) {
                        pred_lr = pred;
                        pred_x[e.row] = pred;
                        if (INDEXED && prm.nmlp == 1 && n_own < (MODE == 2 ? prm.n_end : io.n)) pointio_store_lr(io, n_own, pred);
                    } else if (INDEXED) {
                        if (n_own < (MODE == 2 ? prm.n_end : io.n)) pointio_store(io, n_own, pred, pred_lr);
                    } else if (k < prm.R2) {
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
            if (!INDEXED && lane == 0) ptx::mbar_arrive(&bars->cv_empty[cvb]);
        }
    } else if (warp == NEPI) {
        // =============================== MMA issue ========================================
        // Every wait / commit of an issuing thread is ~100 cycles of serial latency -- with three waits,
        // three commits and eight MMAs per K block one thread needs ~1.7 kcycles where the tensor pipe
        // needs 1024.  So layer 1 is issued by two threads: this one the T0 half (even weight blocks)
        // plus layers 2 and 3, warp NEPI + 2 the T1 half (odd weight blocks).
        if (lane == 0) {
            constexpr uint32_t IDESC256 = ptx::umma_idesc_f16(128, 256);
            constexpr uint32_t IDESC128 = ptx::umma_idesc_f16(128, 128);
            uint32_t wblk = 0, ablk = 0, acc0 = 0, acc1 = 0, wph = 0;     // wph: per-slot phase of full_w (this thread's fills only)
            auto wait_w = [&]() -> uint32_t {
                const uint32_t s = wblk % NSTAGE;
                ptx::mbar_wait(&bars->full_w[s], (wph >> s) & 1u, 30, prof);
                wph ^= 1u << s;
                ptx::tc_fence_after();
                return w_smem + s * W_BLK_BYTES;
            };
            auto release_w = [&]() {
                ptx::umma_commit(&bars->empty_w[wblk % NSTAGE]);
                ++wblk;
            };
            auto wait_a = [&]() -> uint32_t {
                const uint32_t slot = ablk % NA_SLOT;
                ptx::mbar_wait(&bars->a_ready[slot], (ablk / NA_SLOT) & 1u, 31, prof);
                ptx::tc_fence_after();
                return slot;
            };
            // a_free expects two arrivals (one per issuing thread in layer 1).  Where this thread is the only
            // consumer the second arrival is a plain mbarrier arrive: a tcgen05.commit costs ~150 cycles of this
            // thread's serial chain, and the phase still completes only when the commit's arrival lands
            auto release_a = [&](uint32_t slot, int commits) {
                if (commits == 2) ptx::mbar_arrive(&bars->a_free[slot]);
                ptx::umma_commit(&bars->a_free[slot]);
                ++ablk;
            };
            // P = 3: one K block = A slots (hi, lo) x weight blocks (hi, lo): A_hi.W_hi, A_lo.W_hi, A_hi.W_lo --
            // twelve MMAs behind four waits and four commits
            auto kblock3 = [&](uint32_t tmem_d, uint32_t idesc, bool zero_first) {
                const uint32_t s_hi = wait_a();
                const uint32_t w_hi = wait_w();
                mma_block(tmem_d, a_smem + s_hi * A_BLK_BYTES, w_hi, 4, idesc, zero_first);
                ++ablk;
                const uint32_t s_lo = wait_a();
                mma_block(tmem_d, a_smem + s_lo * A_BLK_BYTES, w_hi, 4, idesc, false);
                release_w();
                const uint32_t w_lo = wait_w();
                mma_block(tmem_d, a_smem + s_hi * A_BLK_BYTES, w_lo, 4, idesc, false);
                release_w();
                ptx::umma_commit(&bars->a_free[s_hi]);
                ptx::umma_commit(&bars->a_free[s_lo]);
                ++ablk;
            };
            for (int64_t tile = blockIdx.x; P != 1 && tile < prm.ntiles; tile += gridDim.x) {
                for (int m = 0; m < prm.nmlp; ++m) {
                    for (int half = 0; half < 2; ++half) {
                        // layer 1, one N half: K = 1024 -> T0
                        ptx::mbar_wait(&bars->acc_free[0], (acc0 & 1u) ^ 1u, 33, prof);
                        ptx::tc_fence_after();
                        for (int kb = 0; kb < 16; ++kb) kblock3(T0, IDESC256, kb == 0);
                        ptx::umma_commit(&bars->acc_full[0]);
                        ++acc0;
                        // layer 2, the K half that E1 drains from T0 -> T1
                        if (half == 0) {
                            ptx::mbar_wait(&bars->acc_free[1], (acc1 & 1u) ^ 1u, 35, prof);
                            ptx::tc_fence_after();
                        }
                        for (int kb = 0; kb < 4; ++kb) kblock3(T1, IDESC256, half == 0 && kb == 0);
                    }
                    ptx::umma_commit(&bars->acc_full[1]);
                    ++acc1;
                    // layer 3: K = 256, N = 128 -> T0
                    ptx::mbar_wait(&bars->acc_free[0], (acc0 & 1u) ^ 1u, 36, prof);
                    ptx::tc_fence_after();
                    for (int kb = 0; kb < 4; ++kb) kblock3(T0, IDESC128, kb == 0);
                    ptx::umma_commit(&bars->acc_full[0]);
                    ++acc0;
                }
            }
            for (int64_t tile = blockIdx.x; P == 1 && tile < prm.ntiles; tile += gridDim.x) {
                for (int m = 0; m < prm.nmlp; ++m) {
                    long long tp = PROF ? clock64() : 0;
                    auto phase = [&](int slot) {
                        if (PROF) {
                            const long long t1 = clock64();
                            atomicAdd(g_col_prof + slot, (unsigned long long)(t1 - tp));
                            tp = t1;
                        }
                    };
                    // layer 1, T0 half: K = 1024 (16 blocks), N = 256
                    ptx::mbar_wait(&bars->acc_free[0], (acc0 & 1u) ^ 1u, 33, prof);
                    ptx::mbar_wait(&bars->acc_free[1], (acc1 & 1u) ^ 1u, 34, prof);    // keeps this thread's phase count of acc_free[1]
                    ptx::tc_fence_after();
                    for (int kb = 0; kb < 16; ++kb) {
                        const uint32_t slot = wait_a();
                        const uint32_t w = wait_w();
                        if (!(PROF && (prm.ablate & 4))) mma_block(T0, a_smem + slot * A_BLK_BYTES, w, 4, IDESC256, kb == 0);
                        release_w();
                        ++wblk;                                   // the odd block belongs to the other thread
                        release_a(slot, 1);
                    }
                    ptx::umma_commit(&bars->acc_full[0]);
                    ++acc0; ++acc1;
                    phase(50);
                    // layer 2: K = 512 (y1 halves from E1), N = 256 -> T0
                    ptx::mbar_wait(&bars->acc_free[0], (acc0 & 1u) ^ 1u, 35, prof);
                    ptx::tc_fence_after();
                    for (int kb = 0; kb < 8; ++kb) {
                        const uint32_t slot = wait_a();
                        const uint32_t w = wait_w();
                        mma_block(T0, a_smem + slot * A_BLK_BYTES, w, 4, IDESC256, kb == 0);
                        release_w();
                        release_a(slot, 2);
                    }
                    ptx::umma_commit(&bars->acc_full[0]);
                    ++acc0;
                    phase(51);
                    // layer 3: K = 256, N = 128 -> T1
                    ptx::mbar_wait(&bars->acc_free[1], (acc1 & 1u) ^ 1u, 36, prof);
                    ptx::tc_fence_after();
                    for (int kb = 0; kb < 4; ++kb) {
                        const uint32_t slot = wait_a();
                        const uint32_t w = wait_w();
                        mma_block(T1, a_smem + slot * A_BLK_BYTES, w, 4, IDESC128, kb == 0);
                        release_w();
                        release_a(slot, 2);
                    }
                    ptx::umma_commit(&bars->acc_full[1]);
                    ++acc1;
                    phase(52);
                }
            }
        }
    } else if (warp == NEPI + 2) {
        // =============================== MMA issue, T1 half of layer 1 ======================
        if (lane == 0 && P == 1) {
            constexpr uint32_t IDESC256 = ptx::umma_idesc_f16(128, 256);
            uint32_t wblk = 1, ablk = 0, wph = 0, aph = 0, tph = 0;
            for (int64_t tile = blockIdx.x; tile < prm.ntiles; tile += gridDim.x) {
                for (int m = 0; m < prm.nmlp; ++m) {
                    ptx::mbar_wait(&bars->t1_free_b, tph ^ 1u, 37, nullptr);          // E3 of the previous pass has read T1
                    tph ^= 1u;
                    ptx::tc_fence_after();
                    for (int kb = 0; kb < 16; ++kb) {
                        const uint32_t slot = ablk % NA_SLOT, s = wblk % NSTAGE;
                        ptx::mbar_wait(&bars->a_ready_b[slot], (aph >> slot) & 1u, 38, nullptr);
                        aph ^= 1u << slot;
                        ptx::mbar_wait(&bars->full_wb[s], (wph >> s) & 1u, 39, nullptr);
                        wph ^= 1u << s;
                        ptx::tc_fence_after();
                        if (!(PROF && (prm.ablate & 4))) mma_block(T1, a_smem + slot * A_BLK_BYTES, w_smem + s * W_BLK_BYTES, 4, IDESC256, kb == 0);
                        ptx::umma_commit(&bars->empty_w[s]);
                        ptx::umma_commit(&bars->a_free[slot]);
                        wblk += 2;
                        ++ablk;
                    }
                    ptx::umma_commit(&bars->acc_full[1]);
                    wblk += 12; ablk += 12;                       // layers 2 and 3
                }
            }
        }
    } else {
        // =============================== weight stream + column vectors ====================
        if (lane == 0) {
            uint32_t wblk = 0, it = 0;
            for (int64_t tile = blockIdx.x; tile < prm.ntiles; tile += gridDim.x, ++it) {
                if (!INDEXED) {
                    const uint32_t cvb = it & 1u;
                    ptx::mbar_wait(&bars->cv_empty[cvb], ((it >> 1) & 1u) ^ 1u, 41, prof);
                    ptx::mbar_arrive_expect_tx(&bars->cv_full[cvb], CV_BYTES);
                    ptx::tma_load_1d(smem + SMEM_CV + cvb * CV_BYTES, prm.table + (tile / prm.nseg) * CV_ROW_FLOATS, CV_BYTES, &bars->cv_full[cvb]);
                }
                const uint8_t *src = prm.weights;
                for (int b = 0; b < prm.nmlp * BLOCKS_PER_MLP * RG::WP; ++b) {
                    const int bm = b % (BLOCKS_PER_MLP * RG::WP);
                    const uint32_t bytes = bm < 40 * RG::WP ? W_BLK_BYTES : W128_BLK_BYTES;
                    const uint32_t s = wblk % NSTAGE;
                    ptx::mbar_wait(&bars->empty_w[s], ((wblk / NSTAGE) & 1u) ^ 1u, 40, prof);
                    // odd layer-1 blocks go to the second issuing thread
                    uint64_t *full = (P == 1 && bm < 32 && (b & 1)) ? &bars->full_wb[s] : &bars->full_w[s];
                    if (PROF && (prm.ablate & 1)) {
                        ptx::mbar_arrive(full);
                    } else {
                        ptx::mbar_arrive_expect_tx(full, bytes);
                        ptx::tma_load_1d(smem + SMEM_W + s * W_BLK_BYTES, src, bytes, full);
                    }
                    src += bytes;
                    ++wblk;
                }
            }
        }
    }
    if (PROF && threadIdx.x == 0) atomicAdd(g_col_prof + 0, (unsigned long long)(clock64() - t_kernel0));
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == NEPI) ptx::tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------
// table kernel: per-column products of the weight matrices with the image features
// ------------------------------------------------------------------------------------------
constexpr int TB_NSTAGE = 3;
constexpr int TB_THREADS = 6 * 32;
constexpr int TB_CHUNKS = 2 * TB_CHUNKS_PER_MLP;
// P = 3 (split operands): F as hi + lo tiles (10 K blocks), two weight stages
template <int P>
struct TbLayout {
    static constexpr int NSTAGE = P == 1 ? TB_NSTAGE : 2;
    static constexpr int SMEM_F = 0;
    static constexpr int SMEM_W = (P == 1 ? 5 : 10) * A_BLK_BYTES;
    static constexpr int SMEM_BAR = SMEM_W + NSTAGE * W_BLK_BYTES;
    static constexpr int SMEM_TOTAL = SMEM_BAR + 256 + 1024;
    static_assert(SMEM_TOTAL <= 232448, "shared memory budget");
};

struct TbBars {
    uint64_t full_w[TB_NSTAGE], empty_w[TB_NSTAGE];
    uint64_t acc_full[2], acc_free[2];
    uint64_t f_ready;
    uint32_t tmem_base;
};

struct TbParams {
    const uint8_t *weights;        // 2 x TB_MLP_BYTES
    const float *bias[2][SURS_NUM_LAYERS];
    const float *g0;               // [2][512]: bias of the C1 block
    FeatMaps fm;
    float *table;
    int64_t ncols;
    int64_t point0;                // >= 0: point mode -- table row r belongs to point point0 + r of io (any point source)
    int R1, plane_lo;
    int skip_c1;                   // the C1 block is only read by query_inc.cu (opt-in SURS_COL_INC=1)
};

template <int P>
__global__ void __launch_bounds__(TB_THREADS, 1) col_table_kernel(const __grid_constant__ PointIO io, const __grid_constant__ TbParams prm)
{
    using L = TbLayout<P>;
    constexpr int TB_SMEM_F = L::SMEM_F, TB_SMEM_W = L::SMEM_W, TB_SMEM_BAR = L::SMEM_BAR, NST = L::NSTAGE;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = ptx::smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t *smem = smem_raw + (base - raw);
    const uint32_t f_smem = base + TB_SMEM_F, w_smem = base + TB_SMEM_W;
    TbBars *bars = reinterpret_cast<TbBars *>(smem + TB_SMEM_BAR);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < NST; ++s) { ptx::mbar_init(&bars->full_w[s], 1); ptx::mbar_init(&bars->empty_w[s], 1); }
        for (int t = 0; t < 2; ++t) { ptx::mbar_init(&bars->acc_full[t], 1); ptx::mbar_init(&bars->acc_free[t], 4); }
        ptx::mbar_init(&bars->f_ready, 4);
        ptx::fence_barrier_init();
    }
    if (warp == 4) ptx::tmem_alloc(&bars->tmem_base, 512);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = bars->tmem_base;
    const int64_t ntiles = (prm.ncols + TILE_M - 1) / TILE_M;
    // the C1 block (layer 1 of the all-negative state) only serves query_inc.cu on dense grids: not computed in point mode
    const bool skip_c1 = prm.point0 >= 0 || prm.skip_c1;

    if (warp < 4) {
        const int row = warp * 32 + lane;
        uint32_t acc[2] = {0, 0};
        for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            int64_t col = tile * TILE_M + row;
            const bool valid = col < prm.ncols;
            if (!valid) col = prm.ncols - 1;
            Projected pr;
            if (prm.point0 >= 0) {
                float x, y, z;
                pointio_load(io, prm.point0 + col, x, y, z);
                pr = project_point(io, x, y, z);
            } else {
                const int i = prm.plane_lo + (int)(col / prm.R1), j = (int)(col % prm.R1);
                pr = project_point(io, (float)io.axis[0][i], (float)io.axis[1][j], (float)io.axis[2][0]);
            }
            if (P == 1) gather_rows<32>(prm.fm, pr, warp * 32, lane, f_smem);
            else gather_rows_x3<32>(prm.fm, pr, warp * 32, lane, f_smem);
            ptx::fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&bars->f_ready);
            float *dst_col = prm.table + col * CV_ROW_FLOATS;
#pragma unroll 1
            for (int ch = 0; ch < TB_CHUNKS; ++ch) {
                const int m = ch / TB_CHUNKS_PER_MLP, cc = ch % TB_CHUNKS_PER_MLP, t = ch & 1;
                if (skip_c1 && (cc == 4 || cc == 5)) continue;       // an (even, odd) chunk pair: accumulator parity is kept
                ptx::mbar_wait(&bars->acc_full[t], acc[t] & 1u, 70);
                ptx::tc_fence_after();
                const uint32_t taddr = tmem + t * 256 + ((uint32_t)(warp * 32) << 16);
                const int ncols_out = cc < 7 ? 256 : 128;
                const float *bias = cc < 4 ? prm.bias[m][0] + cc * 256
                                  : cc < 6 ? prm.g0 + m * 512 + (cc - 4) * 256 : (cc == 6 ? prm.bias[m][2] : prm.bias[m][3]);
                float *dst = cc < 4 ? dst_col + m * CV_STRIDE + CV_C0 + cc * 256
                           : cc < 6 ? dst_col + CV_C1 + m * 512 + (cc - 4) * 256
                                    : dst_col + m * CV_STRIDE + (cc == 6 ? CV_C2 : CV_C3);
#pragma unroll 1
                for (int c0 = 0; c0 < ncols_out; c0 += 32) {
                    uint32_t r[32];
                    ptx::tmem_ld32(taddr + c0, r);
                    ptx::tmem_ld_wait();
                    if (valid) {
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            const float4 b = __ldg(reinterpret_cast<const float4 *>(bias + c0 + 4 * q));
                            *reinterpret_cast<float4 *>(dst + c0 + 4 * q) =
                                make_float4(__uint_as_float(r[4 * q]) + b.x, __uint_as_float(r[4 * q + 1]) + b.y,
                                            __uint_as_float(r[4 * q + 2]) + b.z, __uint_as_float(r[4 * q + 3]) + b.w);
                        }
                    }
                }
                if (cc == 7) {                                     // column 128 = W4's skip part . f
                    uint32_t r[32];
                    ptx::tmem_ld32(taddr + 128, r);
                    ptx::tmem_ld_wait();
                    if (valid) dst_col[m * CV_STRIDE + CV_C4] = __uint_as_float(r[0]) + __ldg(prm.bias[m][4]);
                }
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(&bars->acc_free[t]);
                ++acc[t];
            }
        }
    } else if (warp == 4) {
        if (lane == 0) {
            constexpr uint32_t IDESC256 = ptx::umma_idesc_f16(128, 256);
            constexpr uint32_t IDESC144 = ptx::umma_idesc_f16(128, W3_ROWS);
            uint32_t wblk = 0, acc[2] = {0, 0}, fcnt = 0;
            for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                ptx::mbar_wait(&bars->f_ready, fcnt & 1u, 71);
                ++fcnt;
                ptx::tc_fence_after();
                for (int ch = 0; ch < TB_CHUNKS; ++ch) {
                    const int t = ch & 1;
                    if (skip_c1 && (ch % TB_CHUNKS_PER_MLP == 4 || ch % TB_CHUNKS_PER_MLP == 5)) continue;
                    ptx::mbar_wait(&bars->acc_free[t], (acc[t] & 1u) ^ 1u, 72);
                    ptx::tc_fence_after();
                    for (int kp = 0; kp < 5 * P; ++kp) {
                        const int kb = kp / P, part = kp - kb * P;           // P = 3: F_hi.W_hi, F_lo.W_hi, F_hi.W_lo
                        const uint32_t s = wblk % NST;
                        ptx::mbar_wait(&bars->full_w[s], (wblk / NST) & 1u, 73);
                        ptx::tc_fence_after();
                        mma_block(tmem + t * 256, f_smem + (kb + (part == 1 ? 5 : 0)) * A_BLK_BYTES, w_smem + s * W_BLK_BYTES, 4,
                                  (ch % TB_CHUNKS_PER_MLP) == 7 ? IDESC144 : IDESC256, kp == 0);
                        ptx::umma_commit(&bars->empty_w[s]);
                        ++wblk;
                    }
                    ptx::umma_commit(&bars->acc_full[t]);
                    ++acc[t];
                }
            }
        }
    } else {
        if (lane == 0) {
            uint32_t wblk = 0;
            for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                const uint8_t *src = prm.weights;
                for (int ch = 0; ch < TB_CHUNKS; ++ch)
                    for (int kp = 0; kp < 5 * P; ++kp) {
                        const uint32_t bytes = (ch % TB_CHUNKS_PER_MLP) == 7 ? W3_BLK_BYTES : W_BLK_BYTES;
                        if (skip_c1 && (ch % TB_CHUNKS_PER_MLP == 4 || ch % TB_CHUNKS_PER_MLP == 5)) { src += bytes; continue; }
                        const uint32_t s = wblk % NST;
                        ptx::mbar_wait(&bars->empty_w[s], ((wblk / NST) & 1u) ^ 1u, 74);
                        ptx::mbar_arrive_expect_tx(&bars->full_w[s], bytes);
                        ptx::tma_load_1d(smem + TB_SMEM_W + s * W_BLK_BYTES, src, bytes, &bars->full_w[s]);
                        src += bytes;
                        ++wblk;
                    }
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 4) ptx::tmem_dealloc(tmem, 512);
}

// per-MLP constant vectors from the fp32 weights
struct GvSrc {
    const float *w[SURS_NUM_LAYERS];
    const float *b1;
    int cin[SURS_NUM_LAYERS];
    int has_pred;
};
__global__ void build_gv_kernel(GvSrc s0, GvSrc s1, float *gv)
{
    const GvSrc &s = blockIdx.x == 0 ? s0 : s1;
    float *o = gv + blockIdx.x * GV_STRIDE;
    for (int c = threadIdx.x; c < 1024; c += blockDim.x) {
        o[GV_WZ0 + c] = s.w[0][(size_t)c * s.cin[0] + 320];
        o[GV_WP0 + c] = s.has_pred ? s.w[0][(size_t)c * s.cin[0] + 321] : 0.0f;
    }
    for (int c = threadIdx.x; c < 512; c += blockDim.x) o[GV_B1 + c] = s.b1[c];
    for (int c = threadIdx.x; c < 256; c += blockDim.x) {
        o[GV_WZ2 + c] = s.w[2][(size_t)c * s.cin[2] + 512 + 320];
        o[GV_WP2 + c] = s.has_pred ? s.w[2][(size_t)c * s.cin[2] + 512 + 321] : 0.0f;
    }
    for (int c = threadIdx.x; c < 128; c += blockDim.x) {
        o[GV_WZ3 + c] = s.w[3][(size_t)c * s.cin[3] + 256 + 320];
        o[GV_WP3 + c] = s.has_pred ? s.w[3][(size_t)c * s.cin[3] + 256 + 321] : 0.0f;
        o[GV_W4Y + c] = s.w[4][c];
    }
    if (threadIdx.x == 0) {
        o[GV_WZ4] = s.w[4][128 + 320];
        o[GV_WP4] = s.has_pred ? s.w[4][128 + 321] : 0.0f;
        o[GV_WP4 + 1] = 0.0f;
        o[GV_WP4 + 2] = 0.0f;
    }
}


// G = 0.01 W1 W0 (512 x cin0): layer 1 applied to the negative-slope branch of layer 0, plus its
// constant part g0 = 0.01 W1 b0 + b1 and the columns of G that multiply z_feat / pred_lr.
struct GSrc {
    const float *w0t, *w1t;        // transposed fp32: [cin0][1024], [1024][512]
    const float *b0, *b1;
    int cin0;
};
__global__ void build_g_kernel(GSrc s0, GSrc s1, float *G, float *g0, float *qstar, float *rstar)
{
    const int m = blockIdx.y, n = blockIdx.x;
    const GSrc &s = m == 0 ? s0 : s1;
    const int j = threadIdx.x;                                   // column of G; j == G_STRIDE: the constant part
    float acc = 0.0f;
    if (j < s.cin0)
        for (int c = 0; c < 1024; ++c) acc = fmaf(s.w1t[(size_t)c * 512 + n], s.w0t[(size_t)j * 1024 + c], acc);
    else if (j == G_STRIDE)
        for (int c = 0; c < 1024; ++c) acc = fmaf(s.w1t[(size_t)c * 512 + n], s.b0[c], acc);
    acc *= SURS_LEAKY;
    if (j < G_STRIDE) G[((size_t)m * 512 + n) * G_STRIDE + j] = j < s.cin0 ? acc : 0.0f;
    else g0[m * 512 + n] = acc + s.b1[n];
    if (j == 320) qstar[m * 512 + n] = acc;
    if (j == 321) rstar[m * 512 + n] = j < s.cin0 ? acc : 0.0f;
}

__global__ void build_w1h_kernel(const float *w1t_lr, const float *w1t_hr, __half *out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;         // 2 x 1024 x 512
    if (i < 2 * 1024 * 512) out[i] = __float2half_rn(i < 1024 * 512 ? w1t_lr[i] : w1t_hr[i - 1024 * 512]);
}

__global__ void build_xv_kernel(GvSrc s0, GvSrc s1, float *xv)
{
    const GvSrc &s = blockIdx.x == 0 ? s0 : s1;
    float *o = xv + blockIdx.x * XV_LR_FLOATS;
    const bool hp = s.has_pred != 0;
    for (int c = threadIdx.x; c < 1024; c += blockDim.x) {
        o[XV_WZ0 + c] = s.w[0][(size_t)c * s.cin[0] + 320];
        if (hp) o[XV_WP0 + c] = s.w[0][(size_t)c * s.cin[0] + 321];
    }
    for (int c = threadIdx.x; c < 256; c += blockDim.x) {
        o[XV_WZ2 + c] = s.w[2][(size_t)c * s.cin[2] + 512 + 320];
        if (hp) o[XV_WP2 + c] = s.w[2][(size_t)c * s.cin[2] + 512 + 321];
    }
    for (int c = threadIdx.x; c < 128; c += blockDim.x) {
        o[XV_WZ3 + c] = s.w[3][(size_t)c * s.cin[3] + 256 + 320];
        if (hp) o[XV_WP3 + c] = s.w[3][(size_t)c * s.cin[3] + 256 + 321];
        o[XV_W4Y + c] = s.w[4][c];
    }
    if (threadIdx.x < 4) {
        o[XV_WZ4 + threadIdx.x] = threadIdx.x == 0 ? s.w[4][128 + 320] : 0.0f;
        if (hp) o[XV_WP4 + threadIdx.x] = threadIdx.x == 0 ? s.w[4][128 + 321] : 0.0f;
    }
}

}  // namespace

int surs_col_pack_weights(surs_ctx *ctx, const float *const w[2][SURS_NUM_LAYERS], cudaStream_t st)
{
    if (!ctx->col_weights) SURS_CUDA(ctx, cudaMalloc(&ctx->col_weights, COL_WEIGHTS_BYTES));
    uint8_t *base = (uint8_t *)ctx->col_weights;
    float *G = reinterpret_cast<float *>(base + OFF_G);
    GSrc gs[2];
    for (int m = 0; m < 2; ++m) {
        gs[m].w0t = ctx->wt32[m][0]; gs[m].w1t = ctx->wt32[m][1];
        gs[m].b0 = ctx->b32[m][0]; gs[m].b1 = ctx->b32[m][1];
        gs[m].cin0 = ctx->cin[m][0];
    }
    build_g_kernel<<<dim3(512, 2), G_STRIDE + 1, 0, st>>>(gs[0], gs[1], G, reinterpret_cast<float *>(base + OFF_G0),
                                                         reinterpret_cast<float *>(base + OFF_QSTAR), reinterpret_cast<float *>(base + OFF_RSTAR));
    SURS_LAUNCH_CHECK(ctx, "build_g_kernel");

    PackDesc host[2 * BLOCKS_PER_MLP + 2 * 40 + 2 * XW_BLOCKS_PER_MLP];
    int n = 0;
    uint32_t off = 0;
    auto add_src = [&](const float *src, int cin, const float *extra, int row0, int nrows, int ntotal, int fblock, int k0, int c0) {
        PackDesc d;
        memset(&d, 0, sizeof(d));
        d.w = src;
        d.cin = cin;
        d.w_extra = extra;
        d.k0_extra = 128;
        d.row0 = row0; d.nrows = nrows; d.ntotal = ntotal; d.fblock = fblock; d.k0 = k0;
        d.c0 = c0;
        d.out_off = off;
        off += (uint32_t)ntotal * 128u;
        host[n++] = d;
    };
    auto add = [&](int m, int layer, int row0, int nrows, int ntotal, int fblock, int k0, bool extra) {
        add_src(w[m][layer], ctx->cin[m][layer], extra ? w[m][4] : nullptr, row0, nrows, ntotal, fblock, k0, m == 0 ? SURS_C0_LR : SURS_C0_HR);
    };
    for (int m = 0; m < 2; ++m) {                                   // query_col.cu main stream
        for (int kb = 0; kb < 16; ++kb) {
            add(m, 1, 0, 256, 256, -1, kb * 64, false);
            add(m, 1, 256, 256, 256, -1, kb * 64, false);
        }
        for (int kb = 0; kb < 8; ++kb) add(m, 2, 0, 256, 256, -1, kb * 64, false);
        for (int kb = 0; kb < 4; ++kb) add(m, 3, 0, 128, 128, -1, kb * 64, false);
    }
    if (off != OFF_TABLE) SURS_FAIL(ctx, "internal: column weight stream size mismatch");
    for (int m = 0; m < 2; ++m) {                                   // table stream (image-feature columns only)
        for (int c = 0; c < 4; ++c)
            for (int kb = 0; kb < 5; ++kb) add(m, 0, c * 256, 256, 256, kb, 0, false);
        for (int c = 0; c < 2; ++c)
            for (int kb = 0; kb < 5; ++kb) add_src(G + (size_t)m * 512 * G_STRIDE, G_STRIDE, nullptr, c * 256, 256, 256, kb, 0, SURS_C0_LR);
        for (int kb = 0; kb < 5; ++kb) add(m, 2, 0, 256, 256, kb, 512, false);
        for (int kb = 0; kb < 5; ++kb) add(m, 3, 0, 128, W3_ROWS, kb, 256, true);
    }
    if (off != OFF_GV) SURS_FAIL(ctx, "internal: table weight stream size mismatch");
    off = (uint32_t)OFF_XW;
    for (int m = 0; m < 2; ++m) {                                   // query_inc.cu stream: 128-row blocks
        for (int kb = 0; kb < 8; ++kb)
            for (int h = 0; h < 2; ++h) add(m, 2, h * 128, 128, 128, -1, kb * 64, false);
        for (int kb = 0; kb < 4; ++kb) add(m, 3, 0, 128, 128, -1, kb * 64, false);
    }
    if (off != OFF_XV) SURS_FAIL(ctx, "internal: incremental weight stream size mismatch");
    PackDesc *dev = nullptr;
    SURS_CUDA(ctx, cudaMalloc(&dev, sizeof(PackDesc) * n));
    SURS_CUDA(ctx, cudaMemcpyAsync(dev, host, sizeof(PackDesc) * n, cudaMemcpyHostToDevice, st));
    pack_weights_kernel<<<n, 256, 0, st>>>(dev, base);
    SURS_LAUNCH_CHECK(ctx, "pack_weights_kernel(col)");
    // split-operand streams (SURS_PREC_FP16X3).  Main stream: every block as (W_hi, W_lo) -- the A side is a (hi, lo)
    // slot pair and the kernel issues A_hi.W_hi, A_lo.W_hi, A_hi.W_lo; per MLP: layer 1 rows 0-255, layer 2
    // K blocks 0-3, layer 1 rows 256-511, layer 2 K blocks 4-7, layer 3.  Table stream: every block three times as
    // (W_hi, W_hi, W_lo) against F blocks (hi, lo, hi)
    {
        if (!ctx->col_weights_x3) SURS_CUDA(ctx, cudaMalloc(&ctx->col_weights_x3, X3_BYTES));
        const int n1 = 2 * BLOCKS_PER_MLP + 2 * 40;               // main + table descriptors come first in host[]
        PackDesc *h3 = new (std::nothrow) PackDesc[3 * n1];
        if (!h3) SURS_FAIL(ctx, "out of host memory");
        uint32_t o3 = 0;
        int n3 = 0;
        auto emit = [&](int i, int part) {
            PackDesc d = host[i];
            d.lo = part == 2;
            d.out_off = o3;
            o3 += (uint32_t)d.ntotal * 128u;
            h3[n3++] = d;
        };
        for (int m = 0; m < 2; ++m) {                           // main stream in the order of the P = 3 schedule
            const int i0 = m * BLOCKS_PER_MLP;
            for (int half = 0; half < 2; ++half) {
                for (int kb = 0; kb < 16; ++kb)
                    for (int part = 1; part < 3; ++part) emit(i0 + 2 * kb + half, part);
                for (int kb = 0; kb < 4; ++kb)
                    for (int part = 1; part < 3; ++part) emit(i0 + 32 + 4 * half + kb, part);
            }
            for (int kb = 0; kb < 4; ++kb)
                for (int part = 1; part < 3; ++part) emit(i0 + 40 + kb, part);
        }
        if (o3 != X3_TABLE_OFF) { delete[] h3; SURS_FAIL(ctx, "internal: split-operand main stream size mismatch"); }
        for (int i = 2 * BLOCKS_PER_MLP; i < n1; ++i)             // table stream
            for (int part = 0; part < 3; ++part) emit(i, part);
        PackDesc *dev3 = nullptr;
        cudaError_t e3 = o3 == X3_BYTES ? cudaMalloc(&dev3, sizeof(PackDesc) * n3) : cudaErrorInvalidValue;
        if (e3 == cudaSuccess) e3 = cudaMemcpyAsync(dev3, h3, sizeof(PackDesc) * n3, cudaMemcpyHostToDevice, st);
        if (e3 == cudaSuccess) {
            pack_weights_kernel<<<n3, 256, 0, st>>>(dev3, (uint8_t *)ctx->col_weights_x3);
            e3 = cudaStreamSynchronize(st);
        }
        delete[] h3;
        cudaFree(dev3);
        if (e3 != cudaSuccess) SURS_FAIL(ctx, "packing the split-operand weight streams failed: %s", cudaGetErrorString(e3));
        ctx->launches++;
    }
    GvSrc s[2];
    for (int m = 0; m < 2; ++m) {
        for (int l = 0; l < SURS_NUM_LAYERS; ++l) { s[m].w[l] = w[m][l]; s[m].cin[l] = ctx->cin[m][l]; }
        s[m].b1 = ctx->b32[m][1];
        s[m].has_pred = m;
    }
    build_gv_kernel<<<2, 256, 0, st>>>(s[0], s[1], reinterpret_cast<float *>(base + OFF_GV));
    SURS_LAUNCH_CHECK(ctx, "build_gv_kernel");
    build_xv_kernel<<<2, 256, 0, st>>>(s[0], s[1], reinterpret_cast<float *>(base + OFF_XV));
    SURS_LAUNCH_CHECK(ctx, "build_xv_kernel");
    build_w1h_kernel<<<2 * 1024 * 512 / 256, 256, 0, st>>>(ctx->wt32[0][1], ctx->wt32[1][1], reinterpret_cast<__half *>(base + OFF_W1H));
    SURS_LAUNCH_CHECK(ctx, "build_w1h_kernel");
    SURS_CUDA(ctx, cudaStreamSynchronize(st));
    SURS_CUDA(ctx, cudaFree(dev));
    return 0;
}

int surs_col_build_table(surs_ctx *ctx, const PointIO &io, int R1, int plane_lo, int64_t ncols, cudaStream_t st, int passes, int64_t point0)
{
    if (surs_ensure(ctx, (void **)&ctx->col_table, &ctx->col_table_cap, (size_t)ncols * CV_ROW_BYTES)) return 1;
    uint8_t *base = (uint8_t *)ctx->col_weights;
    TbParams tb;
    tb.weights = passes == 3 ? (const uint8_t *)ctx->col_weights_x3 + X3_TABLE_OFF : base + OFF_TABLE;
    for (int m = 0; m < 2; ++m)
        for (int l = 0; l < SURS_NUM_LAYERS; ++l) tb.bias[m][l] = ctx->b32[m][l];
    tb.g0 = reinterpret_cast<const float *>(base + OFF_G0);
    tb.fm.f_lr = ctx->f_lr16; tb.fm.f_hr = ctx->f_hr16;
    tb.fm.f_lr32 = ctx->f_lr32; tb.fm.f_hr32 = ctx->f_hr32;
    tb.fm.H_lr = ctx->H_lr; tb.fm.W_lr = ctx->W_lr; tb.fm.H_hr = ctx->H_hr; tb.fm.W_hr = ctx->W_hr;
    tb.table = (float *)ctx->col_table;
    tb.ncols = ncols; tb.R1 = R1; tb.plane_lo = plane_lo;
    tb.point0 = point0;
    tb.skip_c1 = getenv("SURS_COL_INC") == nullptr;
    const int64_t tb_tiles = (ncols + TILE_M - 1) / TILE_M;
    const int tb_grid = (int)(tb_tiles < ctx->sm_count ? tb_tiles : ctx->sm_count);
    if (passes == 3) {
        SURS_CUDA(ctx, cudaFuncSetAttribute(col_table_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, TbLayout<3>::SMEM_TOTAL));
        col_table_kernel<3><<<tb_grid, TB_THREADS, TbLayout<3>::SMEM_TOTAL, st>>>(io, tb);
    } else {
        SURS_CUDA(ctx, cudaFuncSetAttribute(col_table_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, TbLayout<1>::SMEM_TOTAL));
        col_table_kernel<1><<<tb_grid, TB_THREADS, TbLayout<1>::SMEM_TOTAL, st>>>(io, tb);
    }
    SURS_LAUNCH_CHECK(ctx, "col_table_kernel");
    return 0;
}

// Dense slab evaluation through the column-factored kernels.  io: grid mode, lin_base / n / out_* set
// for planes [plane_lo, plane_lo + nplanes) of a [R0, R1, R2] grid without transform.
int surs_launch_query_col(surs_ctx *ctx, const PointIO &io, int R1, int R2, int plane_lo, int nplanes, cudaStream_t st, int passes)
{
    const int64_t ncols = (int64_t)nplanes * R1;
    if (ncols <= 0) return 0;
    if (surs_col_build_table(ctx, io, R1, plane_lo, ncols, st, passes)) return 1;
    uint8_t *base = (uint8_t *)ctx->col_weights;

    ColParams prm;
    prm.weights = passes == 3 ? (const uint8_t *)ctx->col_weights_x3 : base + OFF_MAIN;
    prm.gv = reinterpret_cast<const float *>(base + OFF_GV);
    prm.table = (const float *)ctx->col_table;
    prm.nseg = (R2 + TILE_M - 1) / TILE_M;
    prm.ntiles = ncols * prm.nseg;
    prm.R1 = R1; prm.R2 = R2; prm.plane_lo = plane_lo;
    prm.n0 = 0; prm.n_end = 0; prm.col0 = 0;
    prm.ablate = getenv("SURS_COL_ABLATE") ? atoi(getenv("SURS_COL_ABLATE")) : 0;
    prm.nmlp = 2;
    const int grid = (int)(prm.ntiles < ctx->sm_count ? prm.ntiles : ctx->sm_count);
    static const bool profile = getenv("SURS_TC_PROFILE") != nullptr;
    if (passes == 3) {
        SURS_CUDA(ctx, cudaFuncSetAttribute(query_col_kernel<false, 0, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, Ring<3>::SMEM_TOTAL));
        query_col_kernel<false, 0, 3><<<grid, NTHREADS, Ring<3>::SMEM_TOTAL, st>>>(io, prm);
        SURS_LAUNCH_CHECK(ctx, "query_col_kernel<x3>");
        return 0;
    }
    // SURS_COL_PAIR=1 (opt-in, measured slower: DESIGN.md): CTA pairs (query_col2.cu) instead of one CTA per tile
    const char *pair_env = getenv("SURS_COL_PAIR");
    const bool pair = pair_env && atoi(pair_env) != 0;
    if (!profile && pair && prm.ntiles >= 2) return surs_launch_query_col_pair(ctx, io, prm, st);
    if (!profile) {
        SURS_CUDA(ctx, cudaFuncSetAttribute(query_col_kernel<false, 0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, Ring<1>::SMEM_TOTAL));
        query_col_kernel<false, 0, 1><<<grid, NTHREADS, Ring<1>::SMEM_TOTAL, st>>>(io, prm);
        SURS_LAUNCH_CHECK(ctx, "query_col_kernel");
        return 0;
    }
    unsigned long long zero[64] = {0}, h[64];
    SURS_CUDA(ctx, cudaMemcpyToSymbol(g_col_prof, zero, sizeof(zero)));
    SURS_CUDA(ctx, cudaFuncSetAttribute(query_col_kernel<true, 0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, Ring<1>::SMEM_TOTAL));
    query_col_kernel<true, 0, 1><<<grid, NTHREADS, Ring<1>::SMEM_TOTAL, st>>>(io, prm);
    SURS_LAUNCH_CHECK(ctx, "query_col_kernel<profile>");
    SURS_CUDA(ctx, cudaStreamSynchronize(st));
    SURS_CUDA(ctx, cudaMemcpyFromSymbol(h, g_col_prof, sizeof(h)));
    const double k = 1e-3 / (double)prm.ntiles;
    fprintf(stderr, "[surs col profile] tiles=%lld grid=%d kcycles/tile: total %.1f | epi warp0: wait_cv %.1f wait_a_free %.1f wait_acc_full(E1a %.1f E1b %.1f E2 %.1f E3 %.1f) "
                    "| mma: wait_w %.1f wait_a_ready %.1f wait_acc_free(%.1f %.1f %.1f %.1f) | loader wait_empty %.1f wait_cv_empty %.1f\n",
            (long long)prm.ntiles, grid, h[0] * k, h[11] * k, h[10] * k, h[20] * k, h[21] * k, h[22] * k, h[23] * k,
            h[30] * k, h[31] * k, h[33] * k, h[34] * k, h[35] * k, h[36] * k, h[40] * k, h[41] * k);
    fprintf(stderr, "[surs col profile] mma warp phases, kcycles/tile (both MLPs): wait acc_free + layer 1 %.1f | layer 2 %.1f | layer 3 %.1f\n", h[50] * k, h[51] * k, h[52] * k);
    if (h[27])
        fprintf(stderr, "[surs col profile] epilogue blocks of warp 0 (%.1f per tile), cycles per block: wait TMEM load %.0f | arithmetic + stores %.0f | fence + arrive %.0f\n",
                (double)h[27] / prm.ntiles, (double)h[24] / h[27], (double)h[25] / h[27], (double)h[26] / h[27]);
    return 0;
}

// Octree levels through the column table: io is in grid mode with idx_list / vol_* set (n selected nodes of a
// [R0, R1, R2] grid without transform); the table must cover all R0 x R1 columns (surs_col_build_table with
// plane_lo = 0), built once per reconstruction.
int surs_launch_query_col_indexed(surs_ctx *ctx, const PointIO &io, int R1, int R2, cudaStream_t st, int passes, int64_t col0, int nmlp)
{
    if (io.n <= 0) return 0;
    uint8_t *base = (uint8_t *)ctx->col_weights;
    ColParams prm;
    prm.weights = passes == 3 ? (const uint8_t *)ctx->col_weights_x3 : base + OFF_MAIN;
    prm.gv = reinterpret_cast<const float *>(base + OFF_GV);
    prm.table = (const float *)ctx->col_table;
    prm.nseg = 1;
    prm.ntiles = (io.n + TILE_M - 1) / TILE_M;
    prm.R1 = R1; prm.R2 = R2; prm.plane_lo = 0;
    prm.n0 = 0; prm.n_end = 0; prm.col0 = col0;
    prm.ablate = 0;
    prm.nmlp = nmlp;
    const int grid = (int)(prm.ntiles < ctx->sm_count ? prm.ntiles : ctx->sm_count);
    if (passes == 3) {
        SURS_CUDA(ctx, cudaFuncSetAttribute(query_col_kernel<false, 1, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, Ring<3>::SMEM_TOTAL));
        query_col_kernel<false, 1, 3><<<grid, NTHREADS, Ring<3>::SMEM_TOTAL, st>>>(io, prm);
    } else {
        SURS_CUDA(ctx, cudaFuncSetAttribute(query_col_kernel<false, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, Ring<1>::SMEM_TOTAL));
        query_col_kernel<false, 1, 1><<<grid, NTHREADS, Ring<1>::SMEM_TOTAL, st>>>(io, prm);
    }
    SURS_LAUNCH_CHECK(ctx, "query_col_kernel<indexed>");
    return 0;
}

// SURS_PREC_FP16X3 for point sources without a column structure (explicit points, transformed grids, sheared
// calibrations, octree lists of such grids): chunks of 128 K points; per chunk the table kernel computes the
// W.f products of every POINT (split operands) and query_col_kernel<MODE 2> runs layers 0-4 on them.  15.4 KB of
// table per point make this path HBM-heavy (30 KB per point written and read back), still ~10x the CUDA-core mode.
int surs_launch_query_generic_x3(surs_ctx *ctx, const PointIO &io, cudaStream_t st)
{
    if (io.n <= 0) return 0;
    // chunk = points per (table, main) kernel pair; the table takes 15.4 KB per point (512 K points: 8 GB).  Larger chunks
    // have fewer partially filled last waves (1 CTA per SM, 128 points per tile); SURS_X3_CHUNK overrides
    static const int64_t CH = getenv("SURS_X3_CHUNK") ? atoll(getenv("SURS_X3_CHUNK")) : 524288;
    uint8_t *base = (uint8_t *)ctx->col_weights;
    for (int64_t s0 = 0; s0 < io.n; s0 += CH) {
        const int64_t len = io.n - s0 < CH ? io.n - s0 : CH;
        if (surs_col_build_table(ctx, io, 1, 0, len, st, 3, s0)) return 1;
        ColParams prm;
        prm.weights = (const uint8_t *)ctx->col_weights_x3;
        prm.gv = reinterpret_cast<const float *>(base + OFF_GV);
        prm.table = (const float *)ctx->col_table;
        prm.nseg = 1;
        prm.ntiles = (len + TILE_M - 1) / TILE_M;
        prm.R1 = 1; prm.R2 = 1; prm.plane_lo = 0;
        prm.n0 = s0; prm.n_end = s0 + len; prm.col0 = 0;
        prm.ablate = 0;
    prm.nmlp = 2;
        const int grid = (int)(prm.ntiles < ctx->sm_count ? prm.ntiles : ctx->sm_count);
        SURS_CUDA(ctx, cudaFuncSetAttribute(query_col_kernel<false, 2, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, Ring<3>::SMEM_TOTAL));
        query_col_kernel<false, 2, 3><<<grid, NTHREADS, Ring<3>::SMEM_TOTAL, st>>>(io, prm);
        SURS_LAUNCH_CHECK(ctx, "query_col_kernel<points, x3>");
    }
    return 0;
}
