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

// 32 consecutive channels of one row -> fp16 -> A ring slot.  v = acc + add + wz * zf + wp * pred.
template <int P, bool HAS_ACC, bool HAS_Z, bool HAS_P>
__device__ __forceinline__ void finish32(const uint32_t *acc, const float *add, const float *wz, const float *wp,
                                         float zf, float pred, uint32_t dst, int row, int hsel, int part)
{
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float v[8];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const float4 a = *reinterpret_cast<const float4 *>(add + 8 * j + 4 * q);
            v[4 * q + 0] = a.x; v[4 * q + 1] = a.y; v[4 * q + 2] = a.z; v[4 * q + 3] = a.w;
            if (HAS_Z) {
                const float4 z = *reinterpret_cast<const float4 *>(wz + 8 * j + 4 * q);
                v[4 * q + 0] = fmaf(z.x, zf, v[4 * q + 0]); v[4 * q + 1] = fmaf(z.y, zf, v[4 * q + 1]);
                v[4 * q + 2] = fmaf(z.z, zf, v[4 * q + 2]); v[4 * q + 3] = fmaf(z.w, zf, v[4 * q + 3]);
            }
            if (HAS_P) {
                const float4 p = *reinterpret_cast<const float4 *>(wp + 8 * j + 4 * q);
                v[4 * q + 0] = fmaf(p.x, pred, v[4 * q + 0]); v[4 * q + 1] = fmaf(p.y, pred, v[4 * q + 1]);
                v[4 * q + 2] = fmaf(p.z, pred, v[4 * q + 2]); v[4 * q + 3] = fmaf(p.w, pred, v[4 * q + 3]);
            }
        }
        if (HAS_ACC) {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] += __uint_as_float(acc[8 * j + i]);
        }
        const uint4 o = make_uint4(act_h2<P>(v[0], v[1], part), act_h2<P>(v[2], v[3], part),
                                   act_h2<P>(v[4], v[5], part), act_h2<P>(v[6], v[7], part));
        st_shared_v4(dst + sw128_off(row, hsel * 4 + j), o);
    }
}

// Layer 0 for 8 channels (one 16-byte chunk) of 4 rows per lane: the per-channel constants are
// loaded once for the four rows.  y0 = leaky(C0 + w_z z (+ w_p pred_lr)).
template <int P, bool HAS_P>
__device__ __forceinline__ void produce8(const float *c0, const float *wz, const float *wp, const float (&zf)[4], const float (&pred)[4],
                                         uint32_t dst, int lane, int chunk, int part)
{
    float a[8], z[8], p[8];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const float4 av = *reinterpret_cast<const float4 *>(c0 + 4 * q), zv = *reinterpret_cast<const float4 *>(wz + 4 * q);
        a[4 * q] = av.x; a[4 * q + 1] = av.y; a[4 * q + 2] = av.z; a[4 * q + 3] = av.w;
        z[4 * q] = zv.x; z[4 * q + 1] = zv.y; z[4 * q + 2] = zv.z; z[4 * q + 3] = zv.w;
        if (HAS_P) {
            const float4 pv = *reinterpret_cast<const float4 *>(wp + 4 * q);
            p[4 * q] = pv.x; p[4 * q + 1] = pv.y; p[4 * q + 2] = pv.z; p[4 * q + 3] = pv.w;
        }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            v[i] = fmaf(z[i], zf[r], a[i]);
            if (HAS_P) v[i] = fmaf(p[i], pred[r], v[i]);
        }
        const uint4 o = make_uint4(act_h2<P>(v[0], v[1], part), act_h2<P>(v[2], v[3], part),
                                   act_h2<P>(v[4], v[5], part), act_h2<P>(v[6], v[7], part));
        st_shared_v4(dst + sw128_off(lane + 32 * r, chunk), o);
    }
}

}  // namespace col

// query_col2.cu: the one-pass dense kernel on CTA pairs (prm prepared by surs_launch_query_col)
int surs_launch_query_col_pair(surs_ctx *ctx, const PointIO &io, const col::ColParams &prm, cudaStream_t st);
