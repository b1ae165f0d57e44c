// Device helpers shared by the column-factored main kernels (query_col.cu: one CTA per tile; query_col2.cu: CTA
// pairs): the launch parameters and the CUDA-core arithmetic that turns per-column constants / accumulators into
// fp16 A-operand blocks.
#pragma once
#include "col_common.cuh"

namespace col {

struct ColParams {
    const uint8_t *weights;        // 2 x MLP_BYTES
    const float *gv;               // [2][GV_STRIDE]
    const float *table;            // [ncols][CV_ROW_FLOATS]
    int64_t ntiles;
    int nseg;                      // tiles per column = ceil(R2 / 128)
    int R1, R2, plane_lo;
    int64_t n0, n_end;             // MODE 2: the launch covers points [n0, n_end) of io
    int64_t col0;                  // MODE 1: the table starts at grid column col0 (a slab's table: plane_lo * R1)
    int ablate;                    // profiling only (SURS_COL_ABLATE): 1 = no weight traffic (results are garbage)
    int nmlp;                      // 2: both MLPs; 1: the LR MLP only (refinement of nodes only the LR surface depends on)
};

// 32 consecutive channels of one row -> fp16 -> A ring slot.  v = ((add + wz * zf) + wp * pred) + acc, in packed
// fp32 pairs (FFMA2 / FADD2: the same roundings as the scalar fmaf / + chain).
template <int P, bool HAS_ACC, bool HAS_Z, bool HAS_P>
__device__ __forceinline__ void finish32(const uint32_t *acc, const float *add, const float *wz, const float *wp,
                                         float zf, float pred, uint32_t dst, int row, int hsel, int part)
{
    const uint64_t zz = pk2(zf, zf), pp = pk2(pred, pred);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        uint64_t v[4];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const float4 a = *reinterpret_cast<const float4 *>(add + 8 * j + 4 * q);
            v[2 * q] = pk2(a.x, a.y); v[2 * q + 1] = pk2(a.z, a.w);
            if (HAS_Z) {
                const float4 z = *reinterpret_cast<const float4 *>(wz + 8 * j + 4 * q);
                v[2 * q] = ffma2(pk2(z.x, z.y), zz, v[2 * q]); v[2 * q + 1] = ffma2(pk2(z.z, z.w), zz, v[2 * q + 1]);
            }
            if (HAS_P) {
                const float4 p = *reinterpret_cast<const float4 *>(wp + 8 * j + 4 * q);
                v[2 * q] = ffma2(pk2(p.x, p.y), pp, v[2 * q]); v[2 * q + 1] = ffma2(pk2(p.z, p.w), pp, v[2 * q + 1]);
            }
        }
        if (HAS_ACC) {
#pragma unroll
            for (int i = 0; i < 4; ++i) v[i] = fadd2(v[i], pk2(__uint_as_float(acc[8 * j + 2 * i]), __uint_as_float(acc[8 * j + 2 * i + 1])));
        }
        float f[8];
#pragma unroll
        for (int i = 0; i < 4; ++i) up2(v[i], f[2 * i], f[2 * i + 1]);
        const uint4 o = make_uint4(act_h2<P>(f[0], f[1], part), act_h2<P>(f[2], f[3], part),
                                   act_h2<P>(f[4], f[5], part), act_h2<P>(f[6], f[7], part));
        st_shared_v4(dst + sw128_off(row, hsel * 4 + j), o);
    }
}

// Layer 0 for 8 channels (one 16-byte chunk) of 4 rows per lane: the per-channel constants are
// loaded once for the four rows.  y0 = leaky(C0 + w_z z (+ w_p pred_lr)), packed fp32 pairs.
template <int P, bool HAS_P>
__device__ __forceinline__ void produce8(const float *c0, const float *wz, const float *wp, const float (&zf)[4], const float (&pred)[4],
                                         uint32_t dst, int lane, int chunk, int part)
{
    uint64_t a[4], z[4], p[4];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const float4 av = *reinterpret_cast<const float4 *>(c0 + 4 * q), zv = *reinterpret_cast<const float4 *>(wz + 4 * q);
        a[2 * q] = pk2(av.x, av.y); a[2 * q + 1] = pk2(av.z, av.w);
        z[2 * q] = pk2(zv.x, zv.y); z[2 * q + 1] = pk2(zv.z, zv.w);
        if (HAS_P) {
            const float4 pv = *reinterpret_cast<const float4 *>(wp + 4 * q);
            p[2 * q] = pk2(pv.x, pv.y); p[2 * q + 1] = pk2(pv.z, pv.w);
        }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
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

}  // namespace col

// query_col2.cu: the one-pass dense kernel on CTA pairs (prm prepared by surs_launch_query_col)
int surs_launch_query_col_pair(surs_ctx *ctx, const PointIO &io, const col::ColParams &prm, cudaStream_t st);
