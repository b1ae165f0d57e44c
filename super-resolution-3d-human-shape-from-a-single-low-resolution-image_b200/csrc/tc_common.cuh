// Pieces shared by the two tensor-core query kernels (query_tc.cu: arbitrary point lists,
// query_col.cu: column-factored dense grids): swizzled K-major operand layout, the feature
// gather, the MMA issue helper and the weight-stream packer.
#pragma once
#include "common.cuh"
#include "ptx.cuh"

namespace tc {

constexpr int TILE_M = 128;
constexpr int A_BLK_BYTES = TILE_M * 128;      // 16 KB: 128 rows x 64 fp16
constexpr int W_BLK_BYTES = 256 * 128;         // 32 KB: 256 rows x 64 fp16
constexpr int W3_ROWS = 144;                   // 128 outputs + W4's skip row + padding to 16
constexpr int W3_BLK_BYTES = W3_ROWS * 128;

// byte offset of 16-byte chunk `chunk` of row `row` inside a [rows x 64] fp16 SWIZZLE_128B block
__host__ __device__ __forceinline__ uint32_t sw128_off(int row, int chunk)
{
    return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((chunk ^ (row & 7)) << 4));
}

__device__ __forceinline__ uint32_t pack_h2(float a, float b)
{
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&h);
}

__device__ __forceinline__ float leaky(float v) { return fmaxf(v, SURS_LEAKY * v); }

// leaky_relu of two fp32 values, rounded to fp16 first: one F2FP + HMUL2 + HMNMX2 for the pair instead of
// 2 FMUL + 2 FMNMX + F2FP.  The negative side is rounded twice (slope fp16(0.01) = 0.0100021): a relative
// 2e-4 on values that are 1 % of the scale -- far inside the fp16 operand rounding of the next layer.
__device__ __forceinline__ uint32_t leaky_h2(float a, float b)
{
    const __half2 h = __floats2half2_rn(a, b);
    const __half2 r = __hmax2(h, __hmul2(h, __float2half2_rn(SURS_LEAKY)));
    return *reinterpret_cast<const uint32_t *>(&r);
}

// Split-operand ("x3") mode: leaky_relu in fp32, then the value as fp16 hi (part 0 / 2) or as the fp16
// residual lo = fp16(y - hi) (part 1).  A.W is evaluated as A_hi.W_hi + A_lo.W_hi + A_hi.W_lo on the tensor
// cores (fp32 accumulate): the dropped A_lo.W_lo term and the rounding of the residuals are ~2^-22 relative.
__device__ __forceinline__ uint32_t split_h2(float a, float b, int part)
{
    const __half2 h = __floats2half2_rn(a, b);
    if (part != 1) return *reinterpret_cast<const uint32_t *>(&h);
    const float2 f = __half22float2(h);
    return pack_h2(a - f.x, b - f.y);
}
// P = 1: packed half2 leaky_relu; P = 3: fp32 leaky_relu + hi / lo split
template <int P>
__device__ __forceinline__ uint32_t act_h2(float a, float b, int part)
{
    if (P == 1) return leaky_h2(a, b);
    return split_h2(leaky(a), leaky(b), part);
}

// Packed fp32 pairs (sm_100: FFMA2 / FADD2): per element exactly fma.rn / add.rn, half the instructions on the FMA pipe
// that bounds the epilogue side of the column kernels.
__device__ __forceinline__ uint64_t pk2(float a, float b)
{
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ void up2(uint64_t r, float &a, float &b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(r)); }
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c)
{
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ uint64_t fmul2(uint64_t a, uint64_t b)
{
    uint64_t d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b)
{
    uint64_t d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

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

struct FeatMaps {
    const __half *f_lr, *f_hr;                 // channels-last fp16
    const float *f_lr32, *f_hr32;              // channels-last fp32 (split-operand mode)
    int H_lr, W_lr, H_hr, W_hr;
};

// Bilinear gather (lib/geometry.py:4-12) of NROWS consecutive rows of an F tile, starting at
// row0, by one warp.  Lane (l % NROWS) holds the projection of row row0 + l % NROWS.
// K blocks 0-3 = low-res channels, block 4 = high-res channels (UMMA SWIZZLE_128B layout).
template <int NROWS>
__device__ __forceinline__ void gather_rows(const FeatMaps &fm, const Projected &pr, int row0, int lane, uint32_t f_smem)
{
    const Taps tl = make_taps(pr.u, pr.v, fm.H_lr, fm.W_lr);
    const Taps th = make_taps(pr.u, pr.v, fm.H_hr, fm.W_hr);
    // low-res map: 256 channels = 32 lanes x 8 channels, one point per step
#pragma unroll 8
    for (int p = 0; p < NROWS; ++p) {
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int off = __shfl_sync(0xffffffffu, tl.off[q], p);
            const float w = __shfl_sync(0xffffffffu, tl.w[q], p);
            if (off >= 0) {
                const uint4 v = __ldg(reinterpret_cast<const uint4 *>(fm.f_lr + (size_t)off * SURS_C_LR) + lane);
                accum_tap(acc, v, w);
            }
        }
        const uint4 o = make_uint4(pack_h2(acc[0], acc[1]), pack_h2(acc[2], acc[3]), pack_h2(acc[4], acc[5]), pack_h2(acc[6], acc[7]));
        st_shared_v4(f_smem + (lane >> 3) * A_BLK_BYTES + sw128_off(row0 + p, lane & 7), o);
    }
    // high-res map: 64 channels = 8 lanes x 8 channels, four points per step
#pragma unroll
    for (int it = 0; it < NROWS / 4; ++it) {
        const int p = it * 4 + (lane >> 3);
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int off = __shfl_sync(0xffffffffu, th.off[q], p);
            const float w = __shfl_sync(0xffffffffu, th.w[q], p);
            if (off >= 0) {
                const uint4 v = __ldg(reinterpret_cast<const uint4 *>(fm.f_hr + (size_t)off * SURS_C_HR) + (lane & 7));
                accum_tap(acc, v, w);
            }
        }
        const uint4 o = make_uint4(pack_h2(acc[0], acc[1]), pack_h2(acc[2], acc[3]), pack_h2(acc[4], acc[5]), pack_h2(acc[6], acc[7]));
        st_shared_v4(f_smem + 4 * A_BLK_BYTES + sw128_off(row0 + p, lane & 7), o);
    }
}

// Split-operand variant: bilinear gather in fp32 from the fp32 maps (exactly lib/geometry.py:4-12), written
// as an fp16 hi tile (5 K blocks at f_smem) and an fp16 lo = fp16(f - hi) tile (5 K blocks behind it).
template <int NROWS>
__device__ __forceinline__ void gather_rows_x3(const FeatMaps &fm, const Projected &pr, int row0, int lane, uint32_t f_smem)
{
    const Taps tl = make_taps(pr.u, pr.v, fm.H_lr, fm.W_lr);
    const Taps th = make_taps(pr.u, pr.v, fm.H_hr, fm.W_hr);
    auto tap = [](float (&acc)[8], const float *src, float w) {
        const float4 a = __ldg(reinterpret_cast<const float4 *>(src)), b = __ldg(reinterpret_cast<const float4 *>(src) + 1);
        acc[0] = fmaf(w, a.x, acc[0]); acc[1] = fmaf(w, a.y, acc[1]); acc[2] = fmaf(w, a.z, acc[2]); acc[3] = fmaf(w, a.w, acc[3]);
        acc[4] = fmaf(w, b.x, acc[4]); acc[5] = fmaf(w, b.y, acc[5]); acc[6] = fmaf(w, b.z, acc[6]); acc[7] = fmaf(w, b.w, acc[7]);
    };
    auto store = [&](const float (&acc)[8], uint32_t dst) {
#pragma unroll
        for (int part = 0; part < 2; ++part) {
            const uint4 o = make_uint4(split_h2(acc[0], acc[1], part), split_h2(acc[2], acc[3], part),
                                       split_h2(acc[4], acc[5], part), split_h2(acc[6], acc[7], part));
            st_shared_v4(dst + part * 5 * A_BLK_BYTES, o);
        }
    };
#pragma unroll 4
    for (int p = 0; p < NROWS; ++p) {
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int off = __shfl_sync(0xffffffffu, tl.off[q], p);
            const float w = __shfl_sync(0xffffffffu, tl.w[q], p);
            if (off >= 0) tap(acc, fm.f_lr32 + (size_t)off * SURS_C_LR + lane * 8, w);
        }
        store(acc, f_smem + (lane >> 3) * A_BLK_BYTES + sw128_off(row0 + p, lane & 7));
    }
#pragma unroll
    for (int it = 0; it < NROWS / 4; ++it) {
        const int p = it * 4 + (lane >> 3);
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int off = __shfl_sync(0xffffffffu, th.off[q], p);
            const float w = __shfl_sync(0xffffffffu, th.w[q], p);
            if (off >= 0) tap(acc, fm.f_hr32 + (size_t)off * SURS_C_HR + (lane & 7) * 8, w);
        }
        store(acc, f_smem + 4 * A_BLK_BYTES + sw128_off(row0 + p, lane & 7));
    }
}

// K-major MMAs over one 64-wide (or 16-wide tail) K block: D[tmem_d] (+)= A[a_addr] . B[w_addr]^T
__device__ __forceinline__ void mma_block(uint32_t tmem_d, uint32_t a_addr, uint32_t w_addr, int ksteps, uint32_t idesc, bool zero_first)
{
    const uint64_t da = ptx::umma_desc_sw128(a_addr), db = ptx::umma_desc_sw128(w_addr);
#pragma unroll 1
    for (int k = 0; k < ksteps; ++k)
        ptx::umma_f16(tmem_d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (zero_first && k == 0) ? 0u : 1u);
}

// ------------------------------------------------------------------------------------------
// weight packing: fp32 [Cout, Cin] (+ bias) -> a stream of pre-swizzled operand blocks
// ------------------------------------------------------------------------------------------
struct PackDesc {
    const float *w;        // source layer, row-major [cout][cin]
    const float *w_extra;  // layer 4 (row 128 of the layer-3 skip blocks) or NULL
    const float *bias;     // bias of the rows (tail blocks only) or NULL
    const float *bias_extra;
    int cin;
    int row0, nrows;       // rows taken from w; the block has ntotal rows, the rest is zero
    int ntotal;
    int fblock;            // -1: plain columns k0 .. k0+63; else F-order block index 0..5
    int bias_only;         // tail block that carries nothing but the bias (layer 1)
    int k0, k0_extra;      // first column (plain) / start of the skip part (F-order)
    int c0;                // 321 / 322: width of the skip input
    int lo;                // split-operand streams: pack the residual w - fp16(w) instead of w
    uint32_t out_off;
};

// column of the skip input that F-tile position (fblock, kk) holds; -1 padding, -2 / -3 bias hi / lo
__device__ __forceinline__ int fmap(int fblock, int kk, int c0)
{
    if (fblock < 4) return fblock * 64 + kk;
    if (fblock == 4) return 256 + kk;
    if (kk < 2) return 320;                      // z_hi, z_lo
    if (kk < 4) return c0 > 321 ? 321 : -1;      // pred_hi, pred_lo (HR MLP only)
    if (kk == 4) return -2;
    if (kk == 5) return -3;
    return -1;
}

static __global__ void pack_weights_kernel(const PackDesc *descs, uint8_t *out)
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
                    if (col >= 0 && !d.bias_only) {
                        if (r < d.nrows) x = d.w[(size_t)(d.row0 + r) * d.cin + d.k0 + col];
                        else if (r == 128 && d.w_extra) x = d.w_extra[d.k0_extra + col];
                    } else if (col <= -2) {
                        float b = 0.0f;
                        if (r < d.nrows && d.bias) b = d.bias[d.row0 + r];
                        else if (r == 128 && d.bias_extra) b = d.bias_extra[0];
                        const float hi = __half2float(__float2half_rn(b));
                        x = col == -2 ? hi : b - hi;
                    }
                }
                v[j] = d.lo ? x - __half2float(__float2half_rn(x)) : x;
            }
            packed[i] = pack_h2(v[0], v[1]);
        }
        *reinterpret_cast<uint4 *>(out + d.out_off + sw128_off(r, c)) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
    }
}

}  // namespace tc
