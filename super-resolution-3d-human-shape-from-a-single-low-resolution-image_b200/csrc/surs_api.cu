// C-ABI entry points of libsurs.so (see include/surs.h for the contract and the
// reference file:line each one replaces).
#include "common.cuh"
#include "col_common.cuh"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <new>
#include <stdlib.h>
#include <thread>
#include <vector>

static char g_create_err[512] = "";

extern "C" int surs_version(void) { return 100; }

int surs_ensure(surs_ctx *ctx, void **ptr, size_t *cap, size_t bytes)
{
    if (*cap >= bytes && *ptr) return 0;
    if (*ptr) SURS_CUDA(ctx, cudaFree(*ptr));
    *ptr = nullptr;
    *cap = 0;
    SURS_CUDA(ctx, cudaMalloc(ptr, bytes));
    *cap = bytes;
    return 0;
}

extern "C" int surs_create(surs_ctx **out, int device)
{
    if (!out) return 1;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || device < 0 || device >= count) {
        snprintf(g_create_err, sizeof(g_create_err), "surs_create: no CUDA device %d (%s)", device,
                 e == cudaSuccess ? "ordinal out of range" : cudaGetErrorString(e));
        return 1;
    }
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    if (prop.major != 10) {
        snprintf(g_create_err, sizeof(g_create_err),
                 "surs_create: device %d is sm_%d%d; libsurs is built for sm_100a (B200) only", device, prop.major, prop.minor);
        return 1;
    }
    surs_ctx *ctx = new (std::nothrow) surs_ctx();
    if (!ctx) return 1;
    memset(ctx, 0, sizeof(*ctx));
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    cudaSetDevice(device);
    if (cudaMalloc(&ctx->counter, 256) != cudaSuccess || surs_mc_init_tables(ctx) != 0) {
        snprintf(g_create_err, sizeof(g_create_err), "surs_create: device allocation failed: %s", ctx->err);
        delete ctx;
        return 1;
    }
    *out = ctx;
    return 0;
}

extern "C" void surs_destroy(surs_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    for (int m = 0; m < 2; ++m)
        for (int l = 0; l < SURS_NUM_LAYERS; ++l) {
            cudaFree(ctx->wt32[m][l]);
            cudaFree(ctx->b32[m][l]);
        }
    cudaFree(ctx->tc_weights);
    cudaFree(ctx->tc_scratch);
    cudaFree(ctx->col_weights);
    cudaFree(ctx->col_weights_x3);
    cudaFree(ctx->col_table);
    cudaFree(ctx->mv_f_lr32); cudaFree(ctx->mv_f_hr32);
    cudaFree(ctx->f_lr32); cudaFree(ctx->f_hr32); cudaFree(ctx->f_lr16); cudaFree(ctx->f_hr16); cudaFree(ctx->feat_stage);
    cudaFree(ctx->axis_dev); cudaFree(ctx->dirty); cudaFree(ctx->idx_list); cudaFree(ctx->counter);
    cudaFree(ctx->stage_pts); cudaFree(ctx->stage_out);
    for (int k = 0; k < 64; ++k)
        if (ctx->oct_ev[k]) cudaEventDestroy((cudaEvent_t)ctx->oct_ev[k]);
    cudaFree(ctx->mc_block_tot); cudaFree(ctx->mc_bits); cudaFree(ctx->mc_cell_tot); cudaFree(ctx->mc_cells); cudaFree(ctx->mc_vid); cudaFree(ctx->mc_tables);
    delete ctx;
}

extern "C" const char *surs_last_error(const surs_ctx *ctx) { return ctx ? ctx->err : g_create_err; }
extern "C" int64_t surs_launch_count(const surs_ctx *ctx) { return ctx ? ctx->launches : 0; }
extern "C" int64_t surs_refined_nodes(const surs_ctx *ctx) { return ctx ? ctx->refined_nodes : 0; }
extern "C" int surs_refine_stats(const surs_ctx *ctx, int64_t *nodes, int64_t *nodes_lr_only, float *max_diff, float *band, int *fell_back,
                                 int *attempts)
{
    if (ctx && attempts) *attempts = ctx->refine_attempts;
    if (!ctx) return 1;
    if (nodes) *nodes = ctx->refined_nodes;
    if (nodes_lr_only) *nodes_lr_only = ctx->refined_lr_only;
    if (max_diff) *max_diff = ctx->refine_maxdiff;
    if (band) *band = ctx->refine_band > 0.0f ? ctx->refine_band : SURS_REFINE_BAND;
    if (fell_back) *fell_back = ctx->refine_fallback;
    return 0;
}

// ------------------------------------------------------------------------------------
// peer arenas: device memory of one rank that the marching-cubes emit kernels of the other ranks write straight into
// over NVLink (CUDA IPC mapping), i.e. the mesh gather fused into the emission
// ------------------------------------------------------------------------------------
extern "C" int surs_arena_create(surs_ctx *ctx, int64_t bytes, void **dev_ptr, unsigned char handle[64])
{
    if (!ctx || !dev_ptr || !handle || bytes <= 0) return 1;
    SURS_CUDA(ctx, cudaSetDevice(ctx->device));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
    void *p = nullptr;
    SURS_CUDA(ctx, cudaMalloc(&p, (size_t)bytes));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        SURS_FAIL(ctx, "cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
    }
    memcpy(handle, &h, 64);
    *dev_ptr = p;
    return 0;
}

extern "C" int surs_arena_open(surs_ctx *ctx, const unsigned char handle[64], void **peer_ptr)
{
    if (!ctx || !handle || !peer_ptr) return 1;
    SURS_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    SURS_CUDA(ctx, cudaIpcOpenMemHandle(peer_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}

extern "C" int surs_arena_close(surs_ctx *ctx, void *peer_ptr)
{
    if (!ctx) return 1;
    SURS_CUDA(ctx, cudaSetDevice(ctx->device));
    SURS_CUDA(ctx, cudaIpcCloseMemHandle(peer_ptr));
    return 0;
}

extern "C" int surs_arena_destroy(surs_ctx *ctx, void *dev_ptr)
{
    if (!ctx) return 1;
    SURS_CUDA(ctx, cudaSetDevice(ctx->device));
    SURS_CUDA(ctx, cudaFree(dev_ptr));
    return 0;
}

// ------------------------------------------------------------------------------------
// parameters
// ------------------------------------------------------------------------------------
__global__ void transpose_kernel(const float *__restrict__ src, float *__restrict__ dst, int rows, int cols)
{
    __shared__ float tile[32][33];
    int c = blockIdx.x * 32 + threadIdx.x, r0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += 8)
        if (r0 + i < rows && c < cols) tile[i][threadIdx.x] = src[(size_t)(r0 + i) * cols + c];
    __syncthreads();
    int r = r0 + threadIdx.x, c0 = blockIdx.x * 32;
    for (int i = threadIdx.y; i < 32; i += 8)
        if (c0 + i < cols && r < rows) dst[(size_t)(c0 + i) * rows + r] = tile[threadIdx.x][i];
}

extern "C" int surs_set_weights(surs_ctx *ctx,
                                const float *const w_lr[SURS_NUM_LAYERS], const float *const b_lr[SURS_NUM_LAYERS],
                                const float *const w_hr[SURS_NUM_LAYERS], const float *const b_hr[SURS_NUM_LAYERS],
                                const int dims_lr[SURS_NUM_LAYERS + 1], const int dims_hr[SURS_NUM_LAYERS + 1],
                                const int *res_layers, int n_res, void *stream)
{
    SURS_NVTX("surs_set_weights");
    if (!ctx) return 1;
    cudaStream_t st = (cudaStream_t)stream;
    SURS_CUDA(ctx, cudaSetDevice(ctx->device));
    static const int want_lr[6] = {SURS_C0_LR, 1024, 512, 256, 128, 1};
    static const int want_hr[6] = {SURS_C0_HR, 1024, 512, 256, 128, 1};
    for (int i = 0; i < 6; ++i)
        if (dims_lr[i] != want_lr[i] || dims_hr[i] != want_hr[i])
            SURS_FAIL(ctx, "surs_set_weights: unsupported mlp_dim (kernels are instantiated for 321/322-1024-512-256-128-1)");
    if (n_res != 3 || res_layers[0] != 2 || res_layers[1] != 3 || res_layers[2] != 4)
        SURS_FAIL(ctx, "surs_set_weights: unsupported mlp_res_layers (kernels are instantiated for [2,3,4])");
    const float *const *w[2] = {w_lr, w_hr};
    const float *const *b[2] = {b_lr, b_hr};
    const int *dims[2] = {dims_lr, dims_hr};
    const float *wsrc[2][SURS_NUM_LAYERS];
    for (int m = 0; m < 2; ++m) {
        memcpy(ctx->dims[m], dims[m], sizeof(int) * 6);
        for (int l = 0; l < SURS_NUM_LAYERS; ++l) {
            const int cout = dims[m][l + 1];
            const int cin = dims[m][l] + ((l >= 2) ? dims[m][0] : 0);
            ctx->cin[m][l] = cin;
            if (!w[m][l] || !b[m][l]) SURS_FAIL(ctx, "surs_set_weights: null parameter pointer");
            if (!ctx->wt32[m][l]) SURS_CUDA(ctx, cudaMalloc(&ctx->wt32[m][l], sizeof(float) * (size_t)cin * cout));
            if (!ctx->b32[m][l]) SURS_CUDA(ctx, cudaMalloc(&ctx->b32[m][l], sizeof(float) * cout));
            dim3 grid((cin + 31) / 32, (cout + 31) / 32), block(32, 8);
            transpose_kernel<<<grid, block, 0, st>>>(w[m][l], ctx->wt32[m][l], cout, cin);
            SURS_LAUNCH_CHECK(ctx, "transpose_kernel");
            SURS_CUDA(ctx, cudaMemcpyAsync(ctx->b32[m][l], b[m][l], sizeof(float) * cout, cudaMemcpyDeviceToDevice, st));
            wsrc[m][l] = w[m][l];
        }
    }
    if (surs_tc_pack_weights(ctx, wsrc, st)) return 1;
    if (surs_col_pack_weights(ctx, wsrc, st)) return 1;
    SURS_CUDA(ctx, cudaStreamSynchronize(st));
    ctx->have_weights = 1;
    return 0;
}

// NCHW fp32 -> NHWC fp32 + fp16.  One block per (pixel tile of 32) x (channel tile of 32).
__global__ void repack_kernel(const float *__restrict__ src, float *__restrict__ dst32, __half *__restrict__ dst16,
                              int C, int HW)
{
    __shared__ float tile[32][33];
    int px = blockIdx.x * 32 + threadIdx.x, c0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += 8)
        if (c0 + i < C && px < HW) tile[i][threadIdx.x] = src[(size_t)(c0 + i) * HW + px];
    __syncthreads();
    int c = c0 + threadIdx.x, p0 = blockIdx.x * 32;
    for (int i = threadIdx.y; i < 32; i += 8)
        if (p0 + i < HW && c < C) {
            float v = tile[threadIdx.x][i];
            dst32[(size_t)(p0 + i) * C + c] = v;
            dst16[(size_t)(p0 + i) * C + c] = __float2half_rn(v);
        }
}

static int ensure_feature_maps(surs_ctx *ctx, int C_lr, int H_lr, int W_lr, int C_hr, int H_hr, int W_hr, bool have_ptrs)
{
    if (C_lr != SURS_C_LR || C_hr != SURS_C_HR)
        SURS_FAIL(ctx, "surs_set_features: expected %d + %d channels, got %d + %d", SURS_C_LR, SURS_C_HR, C_lr, C_hr);
    if (H_lr < 2 || W_lr < 2 || H_hr < 2 || W_hr < 2 || !have_ptrs)
        SURS_FAIL(ctx, "surs_set_features: bad feature map shape");
    const size_t n_lr = (size_t)H_lr * W_lr * C_lr, n_hr = (size_t)H_hr * W_hr * C_hr;
    if (n_lr > ctx->f_lr_cap) {
        cudaFree(ctx->f_lr32); cudaFree(ctx->f_lr16); ctx->f_lr32 = nullptr; ctx->f_lr16 = nullptr; ctx->f_lr_cap = 0;
        SURS_CUDA(ctx, cudaMalloc(&ctx->f_lr32, n_lr * sizeof(float)));
        SURS_CUDA(ctx, cudaMalloc(&ctx->f_lr16, n_lr * sizeof(__half)));
        ctx->f_lr_cap = n_lr;
    }
    if (n_hr > ctx->f_hr_cap) {
        cudaFree(ctx->f_hr32); cudaFree(ctx->f_hr16); ctx->f_hr32 = nullptr; ctx->f_hr16 = nullptr; ctx->f_hr_cap = 0;
        SURS_CUDA(ctx, cudaMalloc(&ctx->f_hr32, n_hr * sizeof(float)));
        SURS_CUDA(ctx, cudaMalloc(&ctx->f_hr16, n_hr * sizeof(__half)));
        ctx->f_hr_cap = n_hr;
    }
    return 0;
}

extern "C" int surs_set_features(surs_ctx *ctx, const float *f_lr, int C_lr, int H_lr, int W_lr,
                                 const float *f_hr, int C_hr, int H_hr, int W_hr, void *stream)
{
    SURS_NVTX("surs_set_features");
    if (!ctx) return 1;
    cudaStream_t st = (cudaStream_t)stream;
    SURS_CUDA(ctx, cudaSetDevice(ctx->device));
    if (ensure_feature_maps(ctx, C_lr, H_lr, W_lr, C_hr, H_hr, W_hr, f_lr && f_hr)) return 1;
    dim3 block(32, 8);
    repack_kernel<<<dim3((H_lr * W_lr + 31) / 32, C_lr / 32), block, 0, st>>>(f_lr, ctx->f_lr32, ctx->f_lr16, C_lr, H_lr * W_lr);
    SURS_LAUNCH_CHECK(ctx, "repack_kernel(lr)");
    repack_kernel<<<dim3((H_hr * W_hr + 31) / 32, C_hr / 32), block, 0, st>>>(f_hr, ctx->f_hr32, ctx->f_hr16, C_hr, H_hr * W_hr);
    SURS_LAUNCH_CHECK(ctx, "repack_kernel(hr)");
    ctx->H_lr = H_lr; ctx->W_lr = W_lr; ctx->H_hr = H_hr; ctx->W_hr = W_hr;
    ctx->feat_u_lo = -1.0f; ctx->feat_u_hi = 1.0f;
    ctx->have_features = 1;
    return 0;
}

// NCHW fp32 stripe [C][H][Ws] (pixel columns x0 .. x0 + Ws - 1 of a W-wide map) -> the same columns of the
// channels-last maps; every other pixel of the maps keeps whatever it held.
__global__ void repack_stripe_kernel(const float *__restrict__ src, float *__restrict__ dst32, __half *__restrict__ dst16,
                                     int C, int H, int W, int x0, int Ws)
{
    __shared__ float tile[32][33];
    const int HWs = H * Ws;
    int px = blockIdx.x * 32 + threadIdx.x, c0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += 8)
        if (c0 + i < C && px < HWs) tile[i][threadIdx.x] = src[(size_t)(c0 + i) * HWs + px];
    __syncthreads();
    int c = c0 + threadIdx.x, p0 = blockIdx.x * 32;
    for (int i = threadIdx.y; i < 32; i += 8) {
        const int p = p0 + i;
        if (p < HWs && c < C) {
            const int y = p / Ws, xs = p - y * Ws;
            const size_t o = ((size_t)y * W + x0 + xs) * C + c;
            const float v = tile[threadIdx.x][i];
            dst32[o] = v;
            dst16[o] = __float2half_rn(v);
        }
    }
}

// pixel columns touched by grid_sample(align_corners=True) for u in [u_lo, u_hi] (lib/geometry.py:11), one pixel of margin
static void stripe_columns(float u_lo, float u_hi, int W, int *x0, int *x1)
{
    const float a = (fminf(fmaxf(u_lo, -1.0f), 1.0f) + 1.0f) * 0.5f * (float)(W - 1);
    const float b = (fminf(fmaxf(u_hi, -1.0f), 1.0f) + 1.0f) * 0.5f * (float)(W - 1);
    int lo = (int)floorf(a) - 1, hi = (int)floorf(b) + 3;            // [lo, hi)
    *x0 = lo < 0 ? 0 : lo;
    *x1 = hi > W ? W : hi;
}

extern "C" int surs_set_features_host(surs_ctx *ctx, const float *f_lr, int C_lr, int H_lr, int W_lr,
                                      const float *f_hr, int C_hr, int H_hr, int W_hr, float u_lo, float u_hi, void *stream)
{
    SURS_NVTX("surs_set_features_host");
    if (!ctx) return 1;
    cudaStream_t st = (cudaStream_t)stream;
    SURS_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!(u_lo <= u_hi)) SURS_FAIL(ctx, "surs_set_features_host: empty u range");
    if (ensure_feature_maps(ctx, C_lr, H_lr, W_lr, C_hr, H_hr, W_hr, f_lr && f_hr)) return 1;
    const float *src[2] = {f_lr, f_hr};
    const int C[2] = {C_lr, C_hr}, H[2] = {H_lr, H_hr}, W[2] = {W_lr, W_hr};
    float *d32[2] = {ctx->f_lr32, ctx->f_hr32};
    __half *d16[2] = {ctx->f_lr16, ctx->f_hr16};
    size_t need = 0;
    for (int k = 0; k < 2; ++k) {
        int x0, x1;
        stripe_columns(u_lo, u_hi, W[k], &x0, &x1);
        const size_t bytes = x1 > x0 ? (size_t)C[k] * H[k] * (x1 - x0) * sizeof(float) : 0;
        need = bytes > need ? bytes : need;
    }
    if (surs_ensure(ctx, (void **)&ctx->feat_stage, &ctx->feat_stage_cap, need)) return 1;
    for (int k = 0; k < 2; ++k) {
        int x0, x1;
        stripe_columns(u_lo, u_hi, W[k], &x0, &x1);
        const int Ws = x1 - x0;
        if (Ws <= 0) continue;
        // rows of Ws floats out of rows of W floats: one strided copy for all C * H rows (pinned source: asynchronous)
        SURS_CUDA(ctx, cudaMemcpy2DAsync(ctx->feat_stage, (size_t)Ws * sizeof(float), src[k] + x0, (size_t)W[k] * sizeof(float),
                                         (size_t)Ws * sizeof(float), (size_t)C[k] * H[k], cudaMemcpyHostToDevice, st));
        repack_stripe_kernel<<<dim3((H[k] * Ws + 31) / 32, C[k] / 32), dim3(32, 8), 0, st>>>((const float *)ctx->feat_stage, d32[k], d16[k], C[k], H[k], W[k], x0, Ws);
        SURS_LAUNCH_CHECK(ctx, "repack_stripe_kernel");
    }
    ctx->H_lr = H_lr; ctx->W_lr = W_lr; ctx->H_hr = H_hr; ctx->W_hr = W_hr;
    ctx->feat_u_lo = u_lo <= -1.0f ? -1.0f : u_lo;
    ctx->feat_u_hi = u_hi >= 1.0f ? 1.0f : u_hi;
    ctx->have_features = 1;
    return 0;
}

// ------------------------------------------------------------------------------------
// query
// ------------------------------------------------------------------------------------
// u_lo / u_hi: the range of image coordinates u the call will sample; the resident features must cover it
// (surs_set_features_host may have uploaded a stripe of pixel columns only)
static int check_ready(surs_ctx *ctx, int precision, float u_lo = -1.0f, float u_hi = 1.0f)
{
    if (!ctx->have_weights) SURS_FAIL(ctx, "surs_set_weights has not been called");
    if (!ctx->have_features) SURS_FAIL(ctx, "surs_set_features has not been called");
    // (the stripe carries a pixel of margin, > 1e-3 in u for any map below 2000 pixels: 1e-4 of slack is safe)
    if (fmaxf(u_lo, -1.0f) < ctx->feat_u_lo - 1e-4f || fminf(u_hi, 1.0f) > ctx->feat_u_hi + 1e-4f)
        SURS_FAIL(ctx, "the resident feature maps cover u in [%g, %g] only (surs_set_features_host stripe); this call samples [%g, %g]",
                  ctx->feat_u_lo, ctx->feat_u_hi, u_lo, u_hi);
    if (precision != SURS_PREC_FP32 && precision != SURS_PREC_FP16 && precision != SURS_PREC_FP16X3 && precision != SURS_PREC_FP16R) SURS_FAIL(ctx, "unknown precision %d", precision);
    return 0;
}

static int run_query(surs_ctx *ctx, const PointIO &io, int precision, cudaStream_t st)
{
    // SURS_PREC_FP16X3 without a column structure: per-point tables through the column kernels (query_col.cu);
    // SURS_PREC_FP16R only differs from it on dense column-factored slabs (surs_eval_grid)
    if (precision == SURS_PREC_FP16X3 || precision == SURS_PREC_FP16R) return surs_launch_query_generic_x3(ctx, io, st);
    return precision == SURS_PREC_FP16 ? surs_launch_query_tc(ctx, io, st) : surs_launch_query_simt(ctx, io, st);
}

static void fill_proj(const surs_ctx *ctx, PointIO &io, const float calib[12], float z_num, float z_den)
{
    memcpy(io.calib, calib, sizeof(float) * 12);
    io.z_num = z_num;
    io.z_den = z_den;
    io.persp = ctx->persp;
    io.has_tf = ctx->has_tf;
    memcpy(io.tf, ctx->tf, sizeof(io.tf));
}

extern "C" int surs_set_projection(surs_ctx *ctx, int perspective, const float *transform)
{
    if (!ctx) return 1;
    ctx->persp = perspective ? 1 : 0;
    ctx->has_tf = transform != nullptr;
    if (transform) memcpy(ctx->tf, transform, sizeof(ctx->tf));
    return 0;
}

extern "C" int surs_query(surs_ctx *ctx, const float *pts, int64_t n, const float calib[12],
                          float z_num, float z_den, int precision, float *pred_hr, float *pred_lr, void *stream)
{
    SURS_NVTX("surs_query");
    if (!ctx) return 1;
    SURS_CUDA(ctx, cudaSetDevice(ctx->device));
    if (check_ready(ctx, precision)) return 1;
    if (n < 0 || (n > 0 && (!pts || !pred_hr || !pred_lr))) SURS_FAIL(ctx, "surs_query: bad arguments");
    PointIO io;
    memset(&io, 0, sizeof(io));
    io.pts = pts; io.n = n; io.out_hr = pred_hr; io.out_lr = pred_lr;
    fill_proj(ctx, io, calib, z_num, z_den);
    return run_query(ctx, io, precision, (cudaStream_t)stream);
}

// NCHW fp32 -> NHWC fp32 only (multi-view maps feed the fp32 kernel)
__global__ void repack32_kernel(const float *__restrict__ src, float *__restrict__ dst32, int C, int HW)
{
    __shared__ float tile[32][33];
    int px = blockIdx.x * 32 + threadIdx.x, c0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += 8)
        if (c0 + i < C && px < HW) tile[i][threadIdx.x] = src[(size_t)(c0 + i) * HW + px];
    __syncthreads();
    int c = c0 + threadIdx.x, p0 = blockIdx.x * 32;
    for (int i = threadIdx.y; i < 32; i += 8)
        if (p0 + i < HW && c < C) dst32[(size_t)(p0 + i) * C + c] = tile[threadIdx.x][i];
}

extern "C" int surs_set_features_views(surs_ctx *ctx, int n_views, const float *f_lr, int C_lr, int H_lr, int W_lr,
                                       const float *f_hr, int C_hr, int H_hr, int W_hr, void *stream)
{
    SURS_NVTX("surs_set_features_views");
    if (!ctx) return 1;
    cudaStream_t st = (cudaStream_t)stream;
    SURS_CUDA(ctx, cudaSetDevice(ctx->device));
    if (n_views < 1 || n_views > 16) SURS_FAIL(ctx, "surs_set_features_views: 1..16 views");
    if (C_lr != SURS_C_LR || C_hr != SURS_C_HR || !f_lr || !f_hr || H_lr < 2 || W_lr < 2 || H_hr < 2 || W_hr < 2)
        SURS_FAIL(ctx, "surs_set_features_views: expected [V,%d,H,W] and [V,%d,H,W] maps", SURS_C_LR, SURS_C_HR);
    const size_t n_lr = (size_t)H_lr * W_lr * C_lr, n_hr = (size_t)H_hr * W_hr * C_hr;
    if (surs_ensure(ctx, (void **)&ctx->mv_f_lr32, &ctx->mv_lr_cap, n_lr * n_views * sizeof(float))) return 1;
    if (surs_ensure(ctx, (void **)&ctx->mv_f_hr32, &ctx->mv_hr_cap, n_hr * n_views * sizeof(float))) return 1;
    for (int v = 0; v < n_views; ++v) {
        repack32_kernel<<<dim3((H_lr * W_lr + 31) / 32, C_lr / 32), dim3(32, 8), 0, st>>>(f_lr + v * n_lr, ctx->mv_f_lr32 + v * n_lr, C_lr, H_lr * W_lr);
        SURS_LAUNCH_CHECK(ctx, "repack32_kernel(lr)");
        repack32_kernel<<<dim3((H_hr * W_hr + 31) / 32, C_hr / 32), dim3(32, 8), 0, st>>>(f_hr + v * n_hr, ctx->mv_f_hr32 + v * n_hr, C_hr, H_hr * W_hr);
        SURS_LAUNCH_CHECK(ctx, "repack32_kernel(hr)");
    }
    ctx->mv_views = n_views;
    ctx->mv_H_lr = H_lr; ctx->mv_W_lr = W_lr; ctx->mv_H_hr = H_hr; ctx->mv_W_hr = W_hr;
    return 0;
}

extern "C" int surs_query_views(surs_ctx *ctx, const float *pts, int64_t n, const float *calibs, float z_num, float z_den,
                                float *pred_hr, float *pred_lr, void *stream)
{
    SURS_NVTX("surs_query_views");
    if (!ctx) return 1;
    SURS_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!ctx->have_weights) SURS_FAIL(ctx, "surs_set_weights has not been called");
    if (ctx->mv_views < 1) SURS_FAIL(ctx, "surs_set_features_views has not been called");
    if (n < 0 || (n > 0 && (!pts || !calibs || !pred_hr || !pred_lr))) SURS_FAIL(ctx, "surs_query_views: bad arguments");
    return surs_launch_query_simt_views(ctx, pts, n, calibs, z_num, z_den, pred_hr, pred_lr, (cudaStream_t)stream);
}

extern "C" int surs_query_host(surs_ctx *ctx, const float *pts_host, int64_t n, const float calib[12],
                               float z_num, float z_den, int precision,
                               float *pred_hr_host, float *pred_lr_host, void *stream)
{
    if (!ctx) return 1;
    cudaStream_t st = (cudaStream_t)stream;
    SURS_CUDA(ctx, cudaSetDevice(ctx->device));
    if (n <= 0) return n < 0;
    size_t need = sizeof(float) * 3 * (size_t)n;
    if (need > ctx->stage_cap) {
        cudaFree(ctx->stage_pts); cudaFree(ctx->stage_out); ctx->stage_pts = ctx->stage_out = nullptr; ctx->stage_cap = 0;
        SURS_CUDA(ctx, cudaMalloc(&ctx->stage_pts, need));
        SURS_CUDA(ctx, cudaMalloc(&ctx->stage_out, sizeof(float) * 2 * (size_t)n));
        ctx->stage_cap = need;
    }
    SURS_CUDA(ctx, cudaMemcpyAsync(ctx->stage_pts, pts_host, need, cudaMemcpyHostToDevice, st));
    if (surs_query(ctx, ctx->stage_pts, n, calib, z_num, z_den, precision, ctx->stage_out, ctx->stage_out + n, stream)) return 1;
    SURS_CUDA(ctx, cudaMemcpyAsync(pred_hr_host, ctx->stage_out, sizeof(float) * n, cudaMemcpyDeviceToHost, st));
    SURS_CUDA(ctx, cudaMemcpyAsync(pred_lr_host, ctx->stage_out + n, sizeof(float) * n, cudaMemcpyDeviceToHost, st));
    SURS_CUDA(ctx, cudaStreamSynchronize(st));
    return 0;
}

// ------------------------------------------------------------------------------------
// grid evaluation
// ------------------------------------------------------------------------------------
// lib/sdf.py:17-24: coords = diag(len/res) @ idx + b_min in float64 (product rounded, then sum rounded).
static int setup_grid(surs_ctx *ctx, PointIO &io, const int res[3], const double b_min[3], const double b_max[3],
                      const double *transform, cudaStream_t st)
{
    if (res[0] < 1 || res[1] < 1 || res[2] < 1) SURS_FAIL(ctx, "bad grid resolution");
    const size_t total = (size_t)res[0] + res[1] + res[2];
    if (surs_ensure(ctx, (void **)&ctx->axis_dev, &ctx->axis_cap, total * sizeof(double))) return 1;
    double *host = (double *)malloc(total * sizeof(double));
    if (!host) SURS_FAIL(ctx, "out of host memory");
    size_t o = 0;
    for (int a = 0; a < 3; ++a) {
        volatile double step = (b_max[a] - b_min[a]) / (double)res[a];
        io.axis[a] = ctx->axis_dev + o;
        for (int i = 0; i < res[a]; ++i) {
            volatile double prod = step * (double)i;     // volatile: no FMA contraction, as numpy
            host[o + i] = prod + b_min[a];
        }
        o += res[a];
    }
    cudaError_t e = cudaMemcpyAsync(ctx->axis_dev, host, total * sizeof(double), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);   // host buffer is pageable and freed below
    free(host);
    if (e != cudaSuccess) SURS_FAIL(ctx, "axis table upload failed: %s", cudaGetErrorString(e));
    io.grid = 1;
    io.R1 = res[1];
    io.R2 = res[2];
    io.has_T = transform != nullptr;
    if (transform) memcpy(io.T, transform, sizeof(double) * 12);
    return 0;
}

extern "C" int surs_eval_grid(surs_ctx *ctx, const int res[3], const double b_min[3], const double b_max[3],
                              const double *transform, const float calib[12], float z_num, float z_den,
                              int precision, int plane_lo, int plane_hi, float *sdf_hr, float *sdf_lr, void *stream)
{
    SURS_NVTX("surs_eval_grid");
    if (!ctx) return 1;
    cudaStream_t st = (cudaStream_t)stream;
    SURS_CUDA(ctx, cudaSetDevice(ctx->device));
    if (plane_lo < 0 || plane_hi > res[0] || plane_lo >= plane_hi) SURS_FAIL(ctx, "surs_eval_grid: bad slab [%d,%d)", plane_lo, plane_hi);
    // image-coordinate range of the slab (extremes of an affine map over a box are at its corners): a stripe of the
    // feature maps (surs_set_features_host) is enough when it covers this range
    float u_lo = -1.0f, u_hi = 1.0f;
    if (!transform && !ctx->persp && !ctx->has_tf) {
        double lo[3], hi[3];
        for (int a = 0; a < 3; ++a) {
            const double step = (b_max[a] - b_min[a]) / res[a];
            lo[a] = b_min[a] + step * (a == 0 ? plane_lo : 0);
            hi[a] = b_min[a] + step * ((a == 0 ? plane_hi : res[a]) - 1);
        }
        double mn = 1e300, mx = -1e300;
        for (int c = 0; c < 8; ++c) {
            const double u = calib[0] * (c & 1 ? hi[0] : lo[0]) + calib[1] * (c & 2 ? hi[1] : lo[1]) + calib[2] * (c & 4 ? hi[2] : lo[2]) + calib[3];
            mn = u < mn ? u : mn;
            mx = u > mx ? u : mx;
        }
        u_lo = (float)(mn - 1e-5);
        u_hi = (float)(mx + 1e-5);
    }
    if (check_ready(ctx, precision, u_lo, u_hi)) return 1;
    PointIO io;
    memset(&io, 0, sizeof(io));
    if (setup_grid(ctx, io, res, b_min, b_max, transform, st)) return 1;
    fill_proj(ctx, io, calib, z_num, z_den);
    const int64_t plane = (int64_t)res[1] * res[2];
    io.lin_base = plane * plane_lo;
    io.n = plane * (plane_hi - plane_lo);
    // outputs are slab-relative: out[n]
    io.out_hr = sdf_hr;
    io.out_lr = sdf_lr;
    // column-factored path: (u,v) must not depend on the grid's last axis (see query_col.cu)
    static const bool no_column = getenv("SURS_NO_COLUMN") != nullptr;
    // SURS_COL_INC=1: layer 1 by incremental updates along the column (query_inc.cu; exact but, as measured,
    // slower than the GEMM of query_col.cu -- experiments/README.md).  Read per call so that tests can toggle it.
    const bool col_inc = getenv("SURS_COL_INC") != nullptr;
    if (precision != SURS_PREC_FP32 && !transform && !ctx->persp && calib[2] == 0.0f && calib[6] == 0.0f && res[2] >= 64 && !no_column) {
        if (precision == SURS_PREC_FP16X3) return surs_launch_query_col(ctx, io, res[1], res[2], plane_lo, plane_hi - plane_lo, st, 3);
        if (precision == SURS_PREC_FP16R) {
            // one pass everywhere, then split operands on the nodes the 0.5 iso-surface can depend on
            const int np = plane_hi - plane_lo;
            if (surs_launch_query_col(ctx, io, res[1], res[2], plane_lo, np, st, 1)) return 1;
            if (surs_ensure(ctx, (void **)&ctx->idx_list, &ctx->idx_cap, 2 * (size_t)io.n * sizeof(int64_t))) return 1;
            int64_t *idx_both = ctx->idx_list, *idx_lr = ctx->idx_list + io.n;
            // (SURS_REFINE_BAND overrides the band, SURS_REFINE_RETRIES the number of widened retries: the tests use a tiny
            // band to drive the retry and the fall-back below)
            float band = getenv("SURS_REFINE_BAND") ? (float)atof(getenv("SURS_REFINE_BAND")) : SURS_REFINE_BAND;
            const int retries = getenv("SURS_REFINE_RETRIES") ? atoi(getenv("SURS_REFINE_RETRIES")) : 1;
            unsigned *maxdiff = reinterpret_cast<unsigned *>(ctx->counter + 4);
            SURS_CUDA(ctx, cudaMemsetAsync(maxdiff, 0, sizeof(unsigned), st));
            ctx->refined_nodes = ctx->refined_lr_only = 0;
            ctx->refine_maxdiff = 0.0f;
            ctx->refine_fallback = 0;
            ctx->refine_attempts = 0;
            bool have_table = false;
            for (int attempt = 0;; ++attempt) {
                int64_t n_both = 0, n_lr = 0;
                ctx->refine_band = band;
                ctx->refine_attempts = attempt + 1;
                if (surs_refine_select_impl(ctx, sdf_hr, sdf_lr, np, res[1], res[2], io.lin_base, SURS_REFINE_LEVEL, band,
                                            idx_both, idx_lr, &n_both, &n_lr, st)) return 1;
                ctx->refined_nodes = n_both + n_lr;          // the last (widest) selection; earlier ones are subsets up to noise
                ctx->refined_lr_only = n_lr;
                if (n_both + n_lr == 0) return 0;
                if (!have_table && surs_col_build_table(ctx, io, res[1], plane_lo, (int64_t)np * res[1], st, 3)) return 1;
                have_table = true;
                PointIO part = io;
                part.out_hr = part.out_lr = nullptr;
                part.vol32_hr = sdf_hr; part.vol32_lr = sdf_lr; part.vol32_base = io.lin_base;
                part.refine_maxdiff = maxdiff;
                // nodes the HR surface depends on: both MLPs; nodes only the LR surface depends on: the LR MLP alone
                part.idx_list = idx_both;
                part.n = n_both;
                if (surs_launch_query_col_indexed(ctx, part, res[1], res[2], st, 3, (int64_t)plane_lo * res[1], 2)) return 1;
                part.idx_list = idx_lr;
                part.n = n_lr;
                if (surs_launch_query_col_indexed(ctx, part, res[1], res[2], st, 3, (int64_t)plane_lo * res[1], 1)) return 1;
                // verify the band on this very input: every re-evaluated value is a sample of the one-pass error (the
                // maximum runs over all attempts; values refined before compare equal to themselves)
                unsigned bits = 0;
                SURS_CUDA(ctx, cudaMemcpyAsync(&bits, maxdiff, sizeof(bits), cudaMemcpyDeviceToHost, st));
                SURS_CUDA(ctx, cudaStreamSynchronize(st));
                memcpy(&ctx->refine_maxdiff, &bits, sizeof(float));
                if (ctx->refine_maxdiff < SURS_REFINE_SAFETY * band) return 0;
                static bool warned = false;
                if (attempt < retries && ctx->refine_maxdiff < 0.15f) {
                    // the one-pass error on this input is larger than the band assumed: widen the band so that the measured
                    // maximum sits at 0.6 of it and select again (what was refined already is re-evaluated to the same value)
                    if (!warned)
                        fprintf(stderr, "libsurs: SURS_PREC_FP16R band %g too narrow for this input (max |one-pass - split| = %g): retrying with %g\n",
                                band, ctx->refine_maxdiff, ctx->refine_maxdiff / 0.6f);
                    warned = true;
                    band = ctx->refine_maxdiff / 0.6f;
                    continue;
                }
                ctx->refine_fallback = 1;
                if (!warned)
                    fprintf(stderr, "libsurs: SURS_PREC_FP16R band check failed (max |one-pass - split| = %g >= %g): "
                                    "re-evaluating the slab with split operands (SURS_PREC_FP16X3)\n",
                            ctx->refine_maxdiff, SURS_REFINE_SAFETY * band);
                warned = true;
                return surs_launch_query_col(ctx, io, res[1], res[2], plane_lo, np, st, 3);
            }
        }
        return col_inc ? surs_launch_query_inc(ctx, io, res[1], res[2], plane_lo, plane_hi - plane_lo, st)
                       : surs_launch_query_col(ctx, io, res[1], res[2], plane_lo, plane_hi - plane_lo, st);
    }
    // one launch handles < 2^31 CTAs; split very large slabs
    const int64_t chunk = (int64_t)1 << 30;
    for (int64_t s = 0; s < io.n; s += chunk) {
        PointIO part = io;
        part.lin_base = io.lin_base + s;
        part.n = (io.n - s < chunk) ? io.n - s : chunk;
        part.out_hr = sdf_hr + s;
        part.out_lr = sdf_lr + s;
        if (run_query(ctx, part, precision, st)) return 1;
    }
    return 0;
}

// octree kernels live in grid.cu
int surs_octree_select_impl(surs_ctx *ctx, const int res[3], int reso, uint8_t *dirty, int64_t *idx,
                            int64_t *n_selected, cudaStream_t st);
int surs_octree_cells_impl(surs_ctx *ctx, const int res[3], int reso, double threshold, double *sdf_hr,
                           double *sdf_lr, uint8_t *dirty, cudaStream_t st);

extern "C" int surs_octree_select(surs_ctx *ctx, const int res[3], int reso, uint8_t *dirty, int64_t *idx,
                                  int64_t *n_selected, void *stream)
{
    if (!ctx) return 1;
    SURS_CUDA(ctx, cudaSetDevice(ctx->device));
    return surs_octree_select_impl(ctx, res, reso, dirty, idx, n_selected, (cudaStream_t)stream);
}

extern "C" int surs_octree_cells(surs_ctx *ctx, const int res[3], int reso, double threshold, double *sdf_hr,
                                 double *sdf_lr, uint8_t *dirty, void *stream)
{
    if (!ctx) return 1;
    SURS_CUDA(ctx, cudaSetDevice(ctx->device));
    return surs_octree_cells_impl(ctx, res, reso, threshold, sdf_hr, sdf_lr, dirty, (cudaStream_t)stream);
}

// phase timing of the octree: CUDA event pairs on the call's stream, summed by surs_octree_stats
namespace {
enum { OCT_INIT = 0, OCT_TABLE = 1, OCT_SELECT = 2, OCT_QUERY = 3, OCT_CELLS = 4, OCT_NPHASE = 5 };
struct OctPhase {
    surs_ctx *ctx;
    cudaStream_t st;
    int pair;
    OctPhase(surs_ctx *c, cudaStream_t s, int phase) : ctx(c), st(s), pair(-1)
    {
        if (ctx->oct_npairs >= 32) return;
        pair = ctx->oct_npairs++;
        for (int k = 0; k < 2; ++k)
            if (!ctx->oct_ev[2 * pair + k]) cudaEventCreate((cudaEvent_t *)&ctx->oct_ev[2 * pair + k]);
        ctx->oct_phase[pair] = phase;
        cudaEventRecord((cudaEvent_t)ctx->oct_ev[2 * pair], st);
    }
    ~OctPhase() { if (pair >= 0) cudaEventRecord((cudaEvent_t)ctx->oct_ev[2 * pair + 1], st); }
};
}  // namespace

extern "C" int surs_octree_stats(surs_ctx *ctx, float ms[5], int64_t counts[6])
{
    if (!ctx) return 1;
    SURS_CUDA(ctx, cudaSetDevice(ctx->device));
    float acc[OCT_NPHASE] = {0, 0, 0, 0, 0};
    for (int p = 0; p < ctx->oct_npairs; ++p) {
        SURS_CUDA(ctx, cudaEventSynchronize((cudaEvent_t)ctx->oct_ev[2 * p + 1]));
        float t = 0.0f;
        SURS_CUDA(ctx, cudaEventElapsedTime(&t, (cudaEvent_t)ctx->oct_ev[2 * p], (cudaEvent_t)ctx->oct_ev[2 * p + 1]));
        acc[ctx->oct_phase[p]] += t;
    }
    if (ms) memcpy(ms, acc, sizeof(acc));
    if (counts) {
        unsigned long long dev[3] = {0, 0, 0};
        SURS_CUDA(ctx, cudaMemcpy(dev, ctx->counter + 16, sizeof(dev), cudaMemcpyDeviceToHost));
        counts[0] = ctx->oct_candidates; counts[1] = ctx->oct_evaluated; counts[2] = ctx->oct_cells;
        counts[3] = (int64_t)dev[0]; counts[4] = (int64_t)dev[1]; counts[5] = (int64_t)dev[2];
    }
    return 0;
}

extern "C" int surs_eval_grid_octree(surs_ctx *ctx, const int res[3], const double b_min[3], const double b_max[3],
                                     const double *transform, const float calib[12], float z_num, float z_den,
                                     int precision, int init_resolution, double threshold,
                                     double *sdf_hr, double *sdf_lr, int64_t *n_evaluated, void *stream)
{
    SURS_NVTX("surs_eval_grid_octree");
    if (!ctx) return 1;
    cudaStream_t st = (cudaStream_t)stream;
    SURS_CUDA(ctx, cudaSetDevice(ctx->device));
    if (check_ready(ctx, precision)) return 1;
    if (init_resolution < 1) SURS_FAIL(ctx, "surs_eval_grid_octree: bad init_resolution");
    const size_t nnode = (size_t)res[0] * res[1] * res[2];
    ctx->oct_npairs = 0;
    ctx->oct_candidates = ctx->oct_evaluated = ctx->oct_cells = 0;
    memcpy(ctx->oct_res, res, sizeof(int) * 3);
    if (n_evaluated) *n_evaluated = 0;
    int reso = res[0] / init_resolution;                        // lib/sdf.py:66
    {
        OctPhase ph(ctx, st, OCT_INIT);
        // lib/sdf.py:60-64: zeros, dirty = ones
        SURS_CUDA(ctx, cudaMemsetAsync(sdf_hr, 0, nnode * sizeof(double), st));
        SURS_CUDA(ctx, cudaMemsetAsync(sdf_lr, 0, nnode * sizeof(double), st));
        SURS_CUDA(ctx, cudaMemsetAsync(ctx->counter + 16, 0, 3 * sizeof(unsigned long long), st));
        if (reso <= 0) return 0;                                // the reference returns zeros
        if (surs_ensure(ctx, (void **)&ctx->dirty, &ctx->dirty_cap, nnode)) return 1;
        SURS_CUDA(ctx, cudaMemsetAsync(ctx->dirty, 1, nnode, st));
    }
    PointIO io;
    memset(&io, 0, sizeof(io));
    if (setup_grid(ctx, io, res, b_min, b_max, transform, st)) return 1;
    fill_proj(ctx, io, calib, z_num, z_den);
    io.vol_hr = sdf_hr;
    io.vol_lr = sdf_lr;
    // Column-table path (same preconditions as the dense column kernels): every W.f product once per column,
    // for all levels; the levels then run the indexed variant of query_col_kernel.  SURS_NO_COLUMN=1 disables it.
    if (precision == SURS_PREC_FP16R) precision = SURS_PREC_FP16X3;   // the octree already evaluates near the surface only
    const int passes = precision == SURS_PREC_FP16X3 ? 3 : 1;
    const bool use_table = precision != SURS_PREC_FP32 && !transform && !ctx->persp && calib[2] == 0.0f && calib[6] == 0.0f && getenv("SURS_NO_COLUMN") == nullptr;
    if (use_table) {
        OctPhase ph(ctx, st, OCT_TABLE);
        if (surs_col_build_table(ctx, io, res[1], 0, (int64_t)res[0] * res[1], st, passes)) return 1;
    }
    while (reso > 0) {
        const size_t cand = (size_t)((res[0] + reso - 1) / reso) * ((res[1] + reso - 1) / reso) * ((res[2] + reso - 1) / reso);
        if (surs_ensure(ctx, (void **)&ctx->idx_list, &ctx->idx_cap, cand * sizeof(int64_t))) return 1;
        int64_t nsel = 0;
        {
            OctPhase ph(ctx, st, OCT_SELECT);
            if (surs_octree_select_impl(ctx, res, reso, ctx->dirty, ctx->idx_list, &nsel, st)) return 1;
        }
        ctx->oct_candidates += (int64_t)cand;
        ctx->oct_evaluated += nsel;
        if (n_evaluated) *n_evaluated += nsel;
        const int64_t chunk = (int64_t)1 << 30;
        {
            OctPhase ph(ctx, st, OCT_QUERY);
            for (int64_t s = 0; s < nsel; s += chunk) {
                PointIO part = io;
                part.idx_list = ctx->idx_list + s;
                part.n = (nsel - s < chunk) ? nsel - s : chunk;
                if (use_table ? surs_launch_query_col_indexed(ctx, part, res[1], res[2], st, passes) : run_query(ctx, part, precision, st)) return 1;
            }
        }
        if (reso <= 1) break;                                   // lib/sdf.py:79
        {
            OctPhase ph(ctx, st, OCT_CELLS);
            if (surs_octree_cells_impl(ctx, res, reso, threshold, sdf_hr, sdf_lr, ctx->dirty, st)) return 1;
        }
        int64_t ncell = 1;
        for (int a = 0; a < 3; ++a) ncell *= res[a] > reso ? (res[a] - 1) / reso : 0;
        ctx->oct_cells += ncell;
        reso /= 2;
    }
    return 0;
}

__global__ void cast_kernel(const double *__restrict__ src, float *__restrict__ dst, int64_t n)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) dst[i] = (float)src[i];
}

extern "C" int surs_cast_f64_f32(surs_ctx *ctx, const double *src, float *dst, int64_t n, void *stream)
{
    if (!ctx) return 1;
    SURS_CUDA(ctx, cudaSetDevice(ctx->device));
    if (n <= 0) return 0;
    int64_t blocks = (n + 255) / 256;
    if (blocks > ctx->sm_count * 32) blocks = ctx->sm_count * 32;
    cast_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(src, dst, n);
    SURS_LAUNCH_CHECK(ctx, "cast_kernel");
    return 0;
}

// ------------------------------------------------------------------------------------
// OBJ writer (host): lib/mesh_util.py:53-61
// ------------------------------------------------------------------------------------
// The text is formatted by all host threads (chunks of 64 K lines, written in order): a 512^3 octree mesh is
// 12 M vertices + 23 M faces, and one thread's snprintf would dominate the whole gen_mesh call.
namespace {
struct ObjChunk {
    std::vector<char> buf;
    size_t used = 0;
    char *room(size_t n)
    {
        if (buf.size() - used < n) buf.resize(buf.size() * 2 + n);
        return buf.data() + used;
    }
};
// "%.4f" without printf for the common case.  printf rounds the EXACT binary value half-to-even at the fourth
// decimal; s = |x| * 1e4 carries a relative rounding error of 2^-53, so for |x| < 1e5 (error < 6e-8 in s) the
// rounding direction is certain unless frac(s) lies within 1e-6 of 0.5 -- those (and non-finite / large values)
// go through snprintf, which keeps the file byte-identical to the reference's '%.4f' (lib/mesh_util.py:56).
inline size_t put_uint(char *dst, uint64_t v)
{
    char tmp[24];
    int n = 0;
    do { tmp[n++] = (char)('0' + v % 10); v /= 10; } while (v);
    for (int i = 0; i < n; ++i) dst[i] = tmp[n - 1 - i];
    return (size_t)n;
}
inline size_t put_fixed4(char *dst, double x)
{
    const double ax = fabs(x);
    if (!(ax < 1e5)) return (size_t)snprintf(dst, 400, "%.4f", x);
    const double sc = ax * 10000.0, fl = floor(sc), frac = sc - fl;
    if (fabs(frac - 0.5) < 1e-6) return (size_t)snprintf(dst, 400, "%.4f", x);
    const uint64_t n = (uint64_t)fl + (frac > 0.5 ? 1u : 0u);
    size_t k = 0;
    if (std::signbit(x)) dst[k++] = '-';
    k += put_uint(dst + k, n / 10000);
    const unsigned q = (unsigned)(n % 10000);
    dst[k++] = '.';
    dst[k++] = (char)('0' + q / 1000); dst[k++] = (char)('0' + q / 100 % 10);
    dst[k++] = (char)('0' + q / 10 % 10); dst[k++] = (char)('0' + q % 10);
    return k;
}
inline size_t put_int(char *dst, int64_t v)
{
    if (v < 0) { dst[0] = '-'; return 1 + put_uint(dst + 1, (uint64_t)(-v)); }
    return put_uint(dst, (uint64_t)v);
}

void obj_format_chunk(ObjChunk &c, const double *verts, int64_t n_verts, const int32_t *faces, int64_t lo, int64_t hi)
{
    c.used = 0;
    if (c.buf.empty()) c.buf.resize((size_t)(hi - lo) * 40 + 1024);
    for (int64_t i = lo; i < hi; ++i) {
        char *dst = c.room(1400);
        size_t k = 0;
        if (i < n_verts) {
            dst[k++] = 'v';
            for (int a = 0; a < 3; ++a) { dst[k++] = ' '; k += put_fixed4(dst + k, verts[3 * i + a]); }
        } else {
            const int64_t t = i - n_verts;                          // lib/mesh_util.py:58-60: 1-based, winding flipped
            dst[k++] = 'f';
            dst[k++] = ' '; k += put_int(dst + k, (int64_t)faces[3 * t] + 1);
            dst[k++] = ' '; k += put_int(dst + k, (int64_t)faces[3 * t + 2] + 1);
            dst[k++] = ' '; k += put_int(dst + k, (int64_t)faces[3 * t + 1] + 1);
        }
        dst[k++] = '\n';
        c.used += k;
    }
}
}  // namespace

extern "C" int surs_save_obj_mesh(const char *path, const double *verts, int64_t n_verts,
                                  const int32_t *faces, int64_t n_faces)
{
    FILE *f = fopen(path, "w");
    if (!f) return 1;
    const int64_t total = n_verts + n_faces, chunk = 1 << 16;
    const int64_t nchunks = (total + chunk - 1) / chunk;
    unsigned hw = std::thread::hardware_concurrency();
    const int nthreads = (int)std::max<int64_t>(1, std::min<int64_t>(std::min<unsigned>(hw ? hw : 1, 32), nchunks));
    bool ok = true;
    const bool timing = getenv("SURS_TIMING") != nullptr;
    double t_fmt = 0.0, t_io = 0.0;
    auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    try {
        std::vector<ObjChunk> chunks(nthreads);
        for (int64_t c0 = 0; c0 < nchunks && ok; c0 += nthreads) {
            const int n = (int)std::min<int64_t>(nthreads, nchunks - c0);
            auto work = [&](int t) {
                const int64_t lo = (c0 + t) * chunk, hi = std::min(total, lo + chunk);
                obj_format_chunk(chunks[t], verts, n_verts, faces, lo, hi);
            };
            const double t0 = now();
            std::vector<std::thread> pool;
            for (int t = 1; t < n; ++t) pool.emplace_back(work, t);
            work(0);
            for (auto &th : pool) th.join();
            const double t1 = now();
            for (int t = 0; t < n && ok; ++t) ok = fwrite(chunks[t].buf.data(), 1, chunks[t].used, f) == chunks[t].used;
            t_fmt += t1 - t0;
            t_io += now() - t1;
        }
    } catch (...) {
        ok = false;
    }
    if (timing) fprintf(stderr, "[surs timing] obj writer: %d threads, format %.1f ms, fwrite %.1f ms\n", nthreads, t_fmt * 1e3, t_io * 1e3);
    return (fclose(f) != 0 || !ok) ? 1 : 0;
}
