// SURS_PREC_FP32: the fused point query on CUDA cores, fp32 FMA chains.
//
// One CTA = 16 points.  Projection (lib/geometry.py:15-31), in-image mask (SuRSNet.py:142),
// depth feature (DepthNormalizer.py:18), bilinear gather of both feature maps
// (lib/geometry.py:4-12) and the two SurfaceClassifier MLPs (SurfaceClassifier.py:45-81)
// run without leaving shared memory; nothing but the two predictions is written.
// This is the accuracy reference mode (agrees with the PyTorch fp32 path to ~1e-6); the
// throughput mode is query_tc.cu.
#include "common.cuh"

namespace {

constexpr int P = 16;          // points per CTA
constexpr int NT = 512;        // threads per CTA (16 warps: one per point in the last layer)

struct SimtParams {
    const float *wt[2][SURS_NUM_LAYERS];   // [Cin][Cout]
    const float *b[2][SURS_NUM_LAYERS];
    const float *f_lr, *f_hr;              // channels-last fp32
    int H_lr, W_lr, H_hr, W_hr;
};

__device__ __forceinline__ float leaky(float x) { return x >= 0.0f ? x : SURS_LEAKY * x; }

// out[c][p] = act(bias[c] + sum_k Wt[k][c] * in1[k][p] + sum_k Wt[K1+k][c] * in2[k][p])
__device__ __forceinline__ void dense_layer(const float *__restrict__ wt, const float *__restrict__ bias, int cout,
                                            const float *in1, int k1, const float *in2, int k2, float *out)
{
    for (int c = threadIdx.x; c < cout; c += NT) {
        float acc[P];
        const float b = __ldg(bias + c);
#pragma unroll
        for (int p = 0; p < P; ++p) acc[p] = b;
        const float *w = wt + c;
#pragma unroll 4
        for (int k = 0; k < k1; ++k) {
            const float wv = __ldg(w + (size_t)k * cout);
            const float4 *x = reinterpret_cast<const float4 *>(in1 + k * P);
#pragma unroll
            for (int q = 0; q < P / 4; ++q) {
                float4 xv = x[q];
                acc[4 * q + 0] = fmaf(wv, xv.x, acc[4 * q + 0]);
                acc[4 * q + 1] = fmaf(wv, xv.y, acc[4 * q + 1]);
                acc[4 * q + 2] = fmaf(wv, xv.z, acc[4 * q + 2]);
                acc[4 * q + 3] = fmaf(wv, xv.w, acc[4 * q + 3]);
            }
        }
        w += (size_t)k1 * cout;
#pragma unroll 4
        for (int k = 0; k < k2; ++k) {
            const float wv = __ldg(w + (size_t)k * cout);
            const float4 *x = reinterpret_cast<const float4 *>(in2 + k * P);
#pragma unroll
            for (int q = 0; q < P / 4; ++q) {
                float4 xv = x[q];
                acc[4 * q + 0] = fmaf(wv, xv.x, acc[4 * q + 0]);
                acc[4 * q + 1] = fmaf(wv, xv.y, acc[4 * q + 1]);
                acc[4 * q + 2] = fmaf(wv, xv.z, acc[4 * q + 2]);
                acc[4 * q + 3] = fmaf(wv, xv.w, acc[4 * q + 3]);
            }
        }
        float4 *o = reinterpret_cast<float4 *>(out + c * P);
#pragma unroll
        for (int q = 0; q < P / 4; ++q)
            o[q] = make_float4(leaky(acc[4 * q]), leaky(acc[4 * q + 1]), leaky(acc[4 * q + 2]), leaky(acc[4 * q + 3]));
    }
}

// last layer (Cout = 1) + sigmoid: warp w handles point w.
__device__ __forceinline__ float final_layer(const float *__restrict__ wt, const float *__restrict__ bias,
                                             const float *in1, int k1, const float *in2, int k2)
{
    const int lane = threadIdx.x & 31, p = threadIdx.x >> 5;
    float s = 0.0f;
    for (int k = lane; k < k1; k += 32) s = fmaf(__ldg(wt + k), in1[k * P + p], s);
    for (int k = lane; k < k2; k += 32) s = fmaf(__ldg(wt + k1 + k), in2[k * P + p], s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    s += __ldg(bias);
    return 1.0f / (1.0f + expf(-s));
}

__global__ void __launch_bounds__(NT, 1) query_simt_kernel(PointIO io, SimtParams prm)
{
    extern __shared__ __align__(16) float smem[];
    float *f = smem;                          // [322][P]
    float *bufA = f + SURS_C0_HR * P;         // [1024][P]
    float *bufB = bufA + 1024 * P;            // [512][P]
    __shared__ Taps taps_lr[P], taps_hr[P];
    __shared__ float s_mask[P];

    const int64_t base = (int64_t)blockIdx.x * P;
    const int tid = threadIdx.x;
    if (tid < P) {
        int64_t n = base + tid;
        float x = 0.f, y = 0.f, z = 0.f;
        if (n < io.n) pointio_load(io, n, x, y, z);
        Projected pr = project_point(io, x, y, z);
        taps_lr[tid] = make_taps(pr.u, pr.v, prm.H_lr, prm.W_lr);
        taps_hr[tid] = make_taps(pr.u, pr.v, prm.H_hr, prm.W_hr);
        s_mask[tid] = pr.mask;
        f[SURS_C_IMG * P + tid] = pr.zf;
        f[(SURS_C_IMG + 1) * P + tid] = 0.0f;
    }
    __syncthreads();
    for (int e = tid; e < SURS_C_IMG * P; e += NT) {
        int p = e / SURS_C_IMG, c = e - p * SURS_C_IMG;
        float v = 0.0f;
        if (c < SURS_C_LR) {
            const Taps &t = taps_lr[p];
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (t.off[q] >= 0) v = fmaf(t.w[q], __ldg(prm.f_lr + (size_t)t.off[q] * SURS_C_LR + c), v);
        } else {
            const Taps &t = taps_hr[p];
            const int ch = c - SURS_C_LR;
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (t.off[q] >= 0) v = fmaf(t.w[q], __ldg(prm.f_hr + (size_t)t.off[q] * SURS_C_HR + ch), v);
        }
        f[c * P + p] = v;
    }
    __syncthreads();

    float pred_lr = 0.0f;
#pragma unroll 1
    for (int m = 0; m < 2; ++m) {
        const int c0 = m == 0 ? SURS_C0_LR : SURS_C0_HR;
        dense_layer(prm.wt[m][0], prm.b[m][0], 1024, f, c0, nullptr, 0, bufA);
        __syncthreads();
        dense_layer(prm.wt[m][1], prm.b[m][1], 512, bufA, 1024, nullptr, 0, bufB);
        __syncthreads();
        dense_layer(prm.wt[m][2], prm.b[m][2], 256, bufB, 512, f, c0, bufA);
        __syncthreads();
        dense_layer(prm.wt[m][3], prm.b[m][3], 128, bufA, 256, f, c0, bufB);
        __syncthreads();
        float pred = final_layer(prm.wt[m][4], prm.b[m][4], bufB, 128, f, c0) * s_mask[tid >> 5];
        if (m == 0) {
            pred_lr = pred;
            if ((tid & 31) == 0) f[(SURS_C_IMG + 1) * P + (tid >> 5)] = pred;   // SuRSNet.py:180
            __syncthreads();
        } else if ((tid & 31) == 0) {
            int64_t n = base + (tid >> 5);
            if (n < io.n) pointio_store(io, n, pred, pred_lr);
        }
    }
}

// ---- multi-view (opt.num_views > 1) -----------------------------------------------------------------------------
// lib/model/SurfaceClassifier.py:70-76: layers 0..2 run per view; after layer 2 (and its leaky ReLU) the activations and
// the skip input are averaged over the views, layers 3, 4 and the sigmoid run once on the means.  Per view the
// prediction is the view's in-image mask times that value (lib/model/SuRSNet.py:156,183), and the HR MLP's 322nd
// input channel is the view's masked LR prediction (:180).
constexpr int MAX_VIEWS = 16;
struct MvParams {
    SimtParams base;
    const float *pts;          // [V][3][n]
    int64_t n;
    int V;
    float calib[MAX_VIEWS][12];
    float z_num, z_den;
    int persp, has_tf;
    float tf[6];
    size_t lr_stride, hr_stride;   // floats per view in the channels-last maps
    float *out_hr, *out_lr;    // [V][n]
};

__global__ void __launch_bounds__(NT, 1) query_simt_mv_kernel(const __grid_constant__ MvParams prm)
{
    extern __shared__ __align__(16) float smem[];
    float *f = smem;                          // [322][P]   this view's input
    float *bufA = f + SURS_C0_HR * P;         // [1024][P]
    float *bufB = bufA + 1024 * P;            // [512][P]
    float *ymean = bufB + 512 * P;            // [256][P]   mean over views of layer 2's output
    float *fmean = ymean + 256 * P;           // [322][P]   mean over views of the input
    __shared__ Taps taps_lr[P], taps_hr[P];
    __shared__ float s_mask[MAX_VIEWS][P], s_lr[MAX_VIEWS][P];

    const int64_t base = (int64_t)blockIdx.x * P;
    const int tid = threadIdx.x;
    const float inv_v = 1.0f / (float)prm.V;
#pragma unroll 1
    for (int m = 0; m < 2; ++m) {
        const int c0 = m == 0 ? SURS_C0_LR : SURS_C0_HR;
        for (int e = tid; e < (256 + SURS_C0_HR) * P; e += NT) ymean[e] = 0.0f;      // ymean and fmean are contiguous
#pragma unroll 1
        for (int v = 0; v < prm.V; ++v) {
            __syncthreads();
            if (tid < P) {
                const int64_t n = base + tid;
                float x = 0.f, y = 0.f, z = 0.f;
                if (n < prm.n) {
                    const float *pv = prm.pts + (size_t)v * 3 * prm.n;
                    x = pv[n]; y = pv[prm.n + n]; z = pv[2 * prm.n + n];
                }
                const Projected pr = project_core(prm.calib[v], prm.z_num, prm.z_den, prm.persp, prm.has_tf, prm.tf, x, y, z);
                taps_lr[tid] = make_taps(pr.u, pr.v, prm.base.H_lr, prm.base.W_lr);
                taps_hr[tid] = make_taps(pr.u, pr.v, prm.base.H_hr, prm.base.W_hr);
                s_mask[v][tid] = pr.mask;
                f[SURS_C_IMG * P + tid] = pr.zf;
                f[(SURS_C_IMG + 1) * P + tid] = m == 1 ? s_lr[v][tid] : 0.0f;
            }
            __syncthreads();
            const float *f_lr = prm.base.f_lr + (size_t)v * prm.lr_stride, *f_hr = prm.base.f_hr + (size_t)v * prm.hr_stride;
            for (int e = tid; e < SURS_C_IMG * P; e += NT) {
                const int p = e / SURS_C_IMG, c = e - p * SURS_C_IMG;
                float val = 0.0f;
                if (c < SURS_C_LR) {
                    const Taps &t = taps_lr[p];
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        if (t.off[q] >= 0) val = fmaf(t.w[q], __ldg(f_lr + (size_t)t.off[q] * SURS_C_LR + c), val);
                } else {
                    const Taps &t = taps_hr[p];
                    const int ch = c - SURS_C_LR;
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        if (t.off[q] >= 0) val = fmaf(t.w[q], __ldg(f_hr + (size_t)t.off[q] * SURS_C_HR + ch), val);
                }
                f[c * P + p] = val;
            }
            __syncthreads();
            dense_layer(prm.base.wt[m][0], prm.base.b[m][0], 1024, f, c0, nullptr, 0, bufA);
            __syncthreads();
            dense_layer(prm.base.wt[m][1], prm.base.b[m][1], 512, bufA, 1024, nullptr, 0, bufB);
            __syncthreads();
            dense_layer(prm.base.wt[m][2], prm.base.b[m][2], 256, bufB, 512, f, c0, bufA);
            __syncthreads();
            for (int e = tid; e < 256 * P; e += NT) ymean[e] += bufA[e];
            for (int e = tid; e < c0 * P; e += NT) fmean[e] += f[e];
        }
        __syncthreads();
        for (int e = tid; e < 256 * P; e += NT) ymean[e] *= inv_v;                    // .mean(dim=1)
        for (int e = tid; e < c0 * P; e += NT) fmean[e] *= inv_v;
        __syncthreads();
        dense_layer(prm.base.wt[m][3], prm.base.b[m][3], 128, ymean, 256, fmean, c0, bufB);
        __syncthreads();
        const float raw = final_layer(prm.base.wt[m][4], prm.base.b[m][4], bufB, 128, fmean, c0);
        const int p = tid >> 5;
        if ((tid & 31) == 0) {
            const int64_t n = base + p;
            for (int v = 0; v < prm.V; ++v) {
                const float pred = raw * s_mask[v][p];
                if (m == 0) s_lr[v][p] = pred;
                else if (n < prm.n) {
                    prm.out_hr[(size_t)v * prm.n + n] = pred;
                    prm.out_lr[(size_t)v * prm.n + n] = s_lr[v][p];
                }
            }
        }
        __syncthreads();
    }
}

}  // namespace

int surs_launch_query_simt_views(surs_ctx *ctx, const float *pts, int64_t n, const float *calibs, float z_num, float z_den,
                                 float *pred_hr, float *pred_lr, cudaStream_t st)
{
    if (n <= 0) return 0;
    MvParams prm;
    memset(&prm, 0, sizeof(prm));
    for (int m = 0; m < 2; ++m)
        for (int l = 0; l < SURS_NUM_LAYERS; ++l) {
            prm.base.wt[m][l] = ctx->wt32[m][l];
            prm.base.b[m][l] = ctx->b32[m][l];
        }
    prm.base.f_lr = ctx->mv_f_lr32; prm.base.f_hr = ctx->mv_f_hr32;
    prm.base.H_lr = ctx->mv_H_lr; prm.base.W_lr = ctx->mv_W_lr; prm.base.H_hr = ctx->mv_H_hr; prm.base.W_hr = ctx->mv_W_hr;
    prm.lr_stride = (size_t)ctx->mv_H_lr * ctx->mv_W_lr * SURS_C_LR;
    prm.hr_stride = (size_t)ctx->mv_H_hr * ctx->mv_W_hr * SURS_C_HR;
    prm.pts = pts; prm.n = n; prm.V = ctx->mv_views;
    memcpy(prm.calib, calibs, sizeof(float) * 12 * ctx->mv_views);
    prm.z_num = z_num; prm.z_den = z_den;
    prm.persp = ctx->persp; prm.has_tf = ctx->has_tf;
    memcpy(prm.tf, ctx->tf, sizeof(prm.tf));
    prm.out_hr = pred_hr; prm.out_lr = pred_lr;
    const size_t smem = (size_t)(SURS_C0_HR + 1024 + 512 + 256 + SURS_C0_HR) * P * sizeof(float);
    static bool attr_set = false;
    if (!attr_set) {
        SURS_CUDA(ctx, cudaFuncSetAttribute(query_simt_mv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    const int64_t blocks = (n + P - 1) / P;
    if (blocks > 0x7fffffffLL) SURS_FAIL(ctx, "too many points for one launch: %lld", (long long)n);
    query_simt_mv_kernel<<<(unsigned)blocks, NT, smem, st>>>(prm);
    SURS_LAUNCH_CHECK(ctx, "query_simt_mv_kernel");
    return 0;
}

int surs_launch_query_simt(surs_ctx *ctx, const PointIO &io, cudaStream_t st)
{
    if (io.n <= 0) return 0;
    SimtParams prm;
    for (int m = 0; m < 2; ++m)
        for (int l = 0; l < SURS_NUM_LAYERS; ++l) {
            prm.wt[m][l] = ctx->wt32[m][l];
            prm.b[m][l] = ctx->b32[m][l];
        }
    prm.f_lr = ctx->f_lr32; prm.f_hr = ctx->f_hr32;
    prm.H_lr = ctx->H_lr; prm.W_lr = ctx->W_lr; prm.H_hr = ctx->H_hr; prm.W_hr = ctx->W_hr;
    const size_t smem = (size_t)(SURS_C0_HR + 1024 + 512) * P * sizeof(float);
    static bool attr_set = false;
    if (!attr_set) {
        SURS_CUDA(ctx, cudaFuncSetAttribute(query_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    const int64_t blocks = (io.n + P - 1) / P;
    if (blocks > 0x7fffffffLL) SURS_FAIL(ctx, "too many points for one launch: %lld", (long long)io.n);
    query_simt_kernel<<<(unsigned)blocks, NT, smem, st>>>(io, prm);
    SURS_LAUNCH_CHECK(ctx, "query_simt_kernel");
    return 0;
}
