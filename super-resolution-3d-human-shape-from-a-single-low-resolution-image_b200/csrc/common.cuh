// Shared declarations of libsurs.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <nvtx3/nvToolsExt.h>

#include "../../include/surs.h"

// NVTX range around every C-ABI entry point (tracing).  Header-only NVTX v3: a no-op unless a tool
// (nsys, ncu --nvtx) is attached.
struct SursRange {
    explicit SursRange(const char *name) { nvtxRangePushA(name); }
    ~SursRange() { nvtxRangePop(); }
};
#define SURS_NVTX(name) SursRange surs_nvtx_range_(name)

#define SURS_C_LR 256        // channels of the 'low_res' hourglass feature (lib/model/SuRSNet.py:59)
#define SURS_C_HR 64         // channels of the 'high_res' feature            (lib/model/SuRSNet.py:61)
#define SURS_C_IMG 320       // gathered image channels (lr then hr, SuRSNet.py:151-153)
#define SURS_C0_LR 321       // + z_feat
#define SURS_C0_HR 322       // + masked pred_lr (SuRSNet.py:180)
#define SURS_REFINE_LEVEL 0.5f   // the iso level of lib/mesh_util.py:40,45
// SURS_PREC_FP16R: nodes within BAND of the level are re-evaluated with split operands.  The band is not trusted, it is
// VERIFIED on every call: the refinement measures max |one-pass - split| over all re-evaluated nodes (the nodes next to
// the iso-surface, where the sigmoid is steepest and the one-pass error largest; 10 M samples at 512^3) and, if that
// exceeds SAFETY x BAND, the whole slab is re-evaluated with split operands (surs_refine_stats reports it).
// Measured one-pass error on the synthetic saturating weights: max 1.49e-2 (profiles/r1_parity_report.json).
#define SURS_REFINE_BAND 0.025f
#define SURS_REFINE_SAFETY 0.8f
#define SURS_LEAKY 0.01f     // F.leaky_relu default slope (SurfaceClassifier.py:66)

// Where the points of a launch come from (explicit list, dense grid slab, or octree index list)
// and where their two predictions go.
struct PointIO {
    // explicit points [3,n]
    const float *pts;
    // grid modes: node index -> coordinate.  lin = linear C-order index into [R0,R1,R2]
    const int64_t *idx_list;   // octree: the launch's n-th point is node idx_list[n]; NULL: node = lin_base + n
    int64_t lin_base;
    int R1, R2;
    const double *axis[3];     // per-axis float64 node coordinates (lib/sdf.py:17-24)
    double T[12];              // optional float64 affine (lib/sdf.py:25-27), row-major 3x4
    int has_T;
    int grid;                  // 0: explicit points, 1: grid nodes
    // projection / depth
    float calib[12];
    float z_num, z_den;
    int persp;                 // 1: lib/geometry.py:34-48 `perspective` (xy / z), 0: :15-31 `orthogonal`
    int has_tf;                // image-space affine `transforms` (lib/geometry.py:27-30,43-46): (u, v) <- tf[:, :2] (u, v) + tf[:, 2]
    float tf[6];
    // outputs
    float *out_hr, *out_lr;    // fp32, index n (explicit / dense slab)
    double *vol_hr, *vol_lr;   // float64 volumes, index = node (octree scatter); used when non-NULL
    float *vol32_hr, *vol32_lr; // fp32 slab volumes, index = node - vol32_base (refinement scatter); used when non-NULL
    int64_t vol32_base;
    unsigned *refine_maxdiff;  // refinement scatter: bits of max |old - new| over the overwritten nodes (float >= 0 orders as uint)
    int64_t n;
};

__device__ __forceinline__ void pointio_load(const PointIO &io, int64_t n, float &x, float &y, float &z)
{
    if (!io.grid) {
        x = io.pts[n];
        y = io.pts[io.n + n];
        z = io.pts[2 * io.n + n];
        return;
    }
    int64_t lin = io.idx_list ? io.idx_list[n] : io.lin_base + n;
    int k = (int)(lin % io.R2);
    int64_t t = lin / io.R2;
    int j = (int)(t % io.R1);
    int i = (int)(t / io.R1);
    double px = io.axis[0][i], py = io.axis[1][j], pz = io.axis[2][k];
    if (io.has_T) {
        double qx = io.T[0] * px + io.T[1] * py + io.T[2] * pz + io.T[3];
        double qy = io.T[4] * px + io.T[5] * py + io.T[6] * pz + io.T[7];
        double qz = io.T[8] * px + io.T[9] * py + io.T[10] * pz + io.T[11];
        px = qx; py = qy; pz = qz;
    }
    x = (float)px; y = (float)py; z = (float)pz;   // the .float() of lib/mesh_util.py:24
}

__device__ __forceinline__ void pointio_store(const PointIO &io, int64_t n, float hr, float lr)
{
    if (io.vol_hr) {
        int64_t lin = io.idx_list ? io.idx_list[n] : io.lin_base + n;
        io.vol_hr[lin] = (double)hr;
        io.vol_lr[lin] = (double)lr;
    } else if (io.vol32_hr) {
        int64_t lin = (io.idx_list ? io.idx_list[n] : io.lin_base + n) - io.vol32_base;
        if (io.refine_maxdiff) {
            const unsigned b = __float_as_uint(fmaxf(fabsf(io.vol32_hr[lin] - hr), fabsf(io.vol32_lr[lin] - lr)));
            if (b > __ldcg(io.refine_maxdiff)) atomicMax(io.refine_maxdiff, b);   // almost never taken after the first few tiles
        }
        io.vol32_hr[lin] = hr;
        io.vol32_lr[lin] = lr;
    } else {
        io.out_hr[n] = hr;
        io.out_lr[n] = lr;
    }
}

// refinement of a node that only the LR surface depends on: the LR volume alone is rewritten
__device__ __forceinline__ void pointio_store_lr(const PointIO &io, int64_t n, float lr)
{
    if (!io.vol32_lr) return;
    const int64_t lin = (io.idx_list ? io.idx_list[n] : io.lin_base + n) - io.vol32_base;
    if (io.refine_maxdiff) {
        const unsigned b = __float_as_uint(fabsf(io.vol32_lr[lin] - lr));
        if (b > __ldcg(io.refine_maxdiff)) atomicMax(io.refine_maxdiff, b);
    }
    io.vol32_lr[lin] = lr;
}

// lib/geometry.py:15-31 + lib/model/SuRSNet.py:142 + lib/model/DepthNormalizer.py:18
struct Projected {
    float u, v, zf, mask;
};
__device__ __forceinline__ Projected project_core(const float *c, float z_num, float z_den, int persp, int has_tf, const float *tf,
                                                  float x, float y, float z)
{
    Projected p;
    p.u = fmaf(c[2], z, fmaf(c[1], y, fmaf(c[0], x, c[3])));
    p.v = fmaf(c[6], z, fmaf(c[5], y, fmaf(c[4], x, c[7])));
    float zz = fmaf(c[10], z, fmaf(c[9], y, fmaf(c[8], x, c[11])));
    if (persp) {                                       // lib/geometry.py:41: xy = homo[:, :2] / homo[:, 2:3]
        p.u = __fdiv_rn(p.u, zz);
        p.v = __fdiv_rn(p.v, zz);
    }
    if (has_tf) {                                      // lib/geometry.py:27-30 / 43-46: baddbmm(shift, scale, xy)
        const float u = p.u, v = p.v;
        p.u = fmaf(tf[1], v, fmaf(tf[0], u, tf[2]));
        p.v = fmaf(tf[4], v, fmaf(tf[3], u, tf[5]));
    }
    p.mask = (p.u >= -1.0f && p.u <= 1.0f && p.v >= -1.0f && p.v <= 1.0f) ? 1.0f : 0.0f;
    p.zf = __fdiv_rn(__fmul_rn(zz, z_num), z_den);
    return p;
}
__device__ __forceinline__ Projected project_point(const PointIO &io, float x, float y, float z)
{
    return project_core(io.calib, io.z_num, io.z_den, io.persp, io.has_tf, io.tf, x, y, z);
}

// grid_sample(align_corners=True, bilinear, zeros) tap set-up (lib/geometry.py:11)
struct Taps {
    int off[4];      // pixel index y*W+x of each corner, or -1 when outside
    float w[4];
};
__device__ __forceinline__ Taps make_taps(float u, float v, int H, int W)
{
    float ix = (u + 1.0f) * 0.5f * (float)(W - 1);
    float iy = (v + 1.0f) * 0.5f * (float)(H - 1);
    float fx = floorf(ix), fy = floorf(iy);
    // clamp before the int conversion so far-away points cannot overflow
    fx = fminf(fmaxf(fx, -2.0f), (float)W + 1.0f);
    fy = fminf(fmaxf(fy, -2.0f), (float)H + 1.0f);
    int x0 = (int)fx, y0 = (int)fy;
    float ax = ix - fx, ay = iy - fy;          // (ix - x0), (iy - y0)
    float bx = (fx + 1.0f) - ix, by = (fy + 1.0f) - iy;
    Taps t;
    const int xs[4] = {x0, x0 + 1, x0, x0 + 1};
    const int ys[4] = {y0, y0, y0 + 1, y0 + 1};
    t.w[0] = bx * by; t.w[1] = ax * by; t.w[2] = bx * ay; t.w[3] = ax * ay;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        bool ok = xs[q] >= 0 && xs[q] < W && ys[q] >= 0 && ys[q] < H;
        t.off[q] = ok ? ys[q] * W + xs[q] : -1;
        if (!ok) t.w[q] = 0.0f;
    }
    return t;
}

struct surs_ctx {
    int device;
    char err[512];
    int64_t launches;
    int64_t refined_nodes;                 // SURS_PREC_FP16R: nodes re-evaluated by the last surs_eval_grid
    int64_t refined_lr_only;               // ... of which only the LR surface depends on (LR MLP alone)
    float refine_band;                     // ... the band it used
    float refine_maxdiff;                  // ... max |one-pass - split| over them
    int refine_attempts;                   // ... selections it took (> 1: the band had to be widened)
    int refine_fallback;                   // ... 1: the band check failed and the whole slab was re-evaluated with split operands
    void *mc_count_stream;                 // stream of the last surs_mc_count (surs_mc_interior_stats)
    void *mc_faces_stream;                 // stream of the last surs_mc_emit_faces (surs_mc_seam_violations reads its counter)
    int sm_count;
    // ---- MLP parameters -------------------------------------------------------
    int have_weights;
    int dims[2][SURS_NUM_LAYERS + 1];      // [0] = lr, [1] = hr
    int cin[2][SURS_NUM_LAYERS];           // input width of each conv (incl. skip)
    float *wt32[2][SURS_NUM_LAYERS];       // fp32, transposed: [Cin][Cout]
    float *b32[2][SURS_NUM_LAYERS];
    void *tc_weights;                      // packed fp16 operand images (query_tc.cu)
    size_t tc_weights_bytes;
    void *tc_scratch;                      // 64 KB per CTA: first half of layer 1 between passes
    size_t tc_scratch_cap;
    void *col_weights;                     // column-factored dense path: weight streams + constant vectors
    void *col_weights_x3;                  // split-operand (hi, hi, lo) copy of the main + table weight streams
    void *col_table;                       // per-column vectors (query_col.cu)
    size_t col_table_cap;
    float tc_w4y[2][128];                  // W4[0, 0:128] of both MLPs (host copy, passed as kernel parameter)
    // ---- projection variant used by the following query / grid calls (surs_set_projection) ----
    int persp, has_tf;
    float tf[6];
    // ---- multi-view features (surs_set_features_views): [V][H][W][C] fp32, channels-last ----
    int mv_views;
    float *mv_f_lr32, *mv_f_hr32;
    size_t mv_lr_cap, mv_hr_cap;
    int mv_H_lr, mv_W_lr, mv_H_hr, mv_W_hr;
    // ---- features, channels-last ------------------------------------------------
    int have_features;
    int H_lr, W_lr, H_hr, W_hr;
    float *f_lr32, *f_hr32;                // [H][W][C] fp32
    __half *f_lr16, *f_hr16;               // [H][W][C] fp16
    size_t f_lr_cap, f_hr_cap;
    float feat_u_lo, feat_u_hi;            // image-coordinate range covered by the resident maps ([-1, 1] = whole maps)
    void *feat_stage;                      // NCHW stripe staging of surs_set_features_host
    size_t feat_stage_cap;
    // ---- grid scratch -----------------------------------------------------------
    double *axis_dev;                      // 3 per-axis coordinate tables
    size_t axis_cap;
    uint8_t *dirty;                        // octree dirty flags
    size_t dirty_cap;
    int64_t *idx_list;                     // octree compaction output
    size_t idx_cap;
    unsigned long long *counter;           // device scalars: [0..7] select / marching cubes, [4] refinement max diff, [8] seam
                                           // violations, [16..18] octree statistics
    // phase timing of the last surs_eval_grid_octree (surs_octree_stats): event pairs, phase id per pair
    void *oct_ev[64];
    int oct_phase[32];
    int oct_npairs;
    int64_t oct_candidates, oct_evaluated, oct_cells;
    int oct_res[3];
    float *stage_pts, *stage_out;          // staging for surs_query_host
    size_t stage_cap;
    // ---- marching cubes state (between count and emit) ---------------------------
    const float *mc_vol;
    int mc_res[3];
    float mc_level;
    int mc_flags;
    int64_t mc_nv, mc_nf, mc_id_offset;
    int mc_fast;                           // R2 % 4 == 0: quad classification + compact active-cell list
    int64_t mc_nact;
    float mc_vmin, mc_vmax;                // value range of the last counted volume (surs_mc_value_range)
    unsigned mc_range_host[2];
    int64_t mc_cap_hint;                   // active cells of the previous volume: estimate of the next list length
    void *mc_cells;                        // active cells in scan order (CellRec, mc.cu)
    size_t mc_cells_cap;
    void *mc_block_tot;                    // per-block totals, then exclusive prefix (uint2 / uint4)
    size_t mc_block_cap;
    void *mc_bits;                         // 1 bit per node: value > level (fast path)
    size_t mc_bits_cap;
    void *mc_cell_tot;                     // per-block vertex / face totals of the active-cell list, then exclusive prefix
    size_t mc_cell_tot_cap;
    int32_t *mc_vid;                       // edge -> vertex id map, 3 per node
    size_t mc_vid_cap;
    void *mc_tables;                       // device copy of the case tables
};

#define SURS_FAIL(ctx, ...)                                  \
    do {                                                     \
        snprintf((ctx)->err, sizeof((ctx)->err), __VA_ARGS__); \
        return 1;                                            \
    } while (0)

#define SURS_CUDA(ctx, call)                                                            \
    do {                                                                                \
        cudaError_t e_ = (call);                                                        \
        if (e_ != cudaSuccess)                                                          \
            SURS_FAIL(ctx, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

#define SURS_LAUNCH_CHECK(ctx, name)                                                    \
    do {                                                                                \
        cudaError_t e_ = cudaGetLastError();                                            \
        if (e_ != cudaSuccess)                                                          \
            SURS_FAIL(ctx, "launch of %s failed: %s", name, cudaGetErrorString(e_));    \
        (ctx)->launches++;                                                              \
    } while (0)

int surs_ensure(surs_ctx *ctx, void **ptr, size_t *cap, size_t bytes);

// query_simt.cu
int surs_launch_query_simt(surs_ctx *ctx, const PointIO &io, cudaStream_t st);
int surs_launch_query_simt_views(surs_ctx *ctx, const float *pts, int64_t n, const float *calibs, float z_num, float z_den,
                                 float *pred_hr, float *pred_lr, cudaStream_t st);
// query_tc.cu
int surs_tc_pack_weights(surs_ctx *ctx, const float *const w[2][SURS_NUM_LAYERS], cudaStream_t st);
int surs_launch_query_tc(surs_ctx *ctx, const PointIO &io, cudaStream_t st);
// query_col.cu
int surs_col_pack_weights(surs_ctx *ctx, const float *const w[2][SURS_NUM_LAYERS], cudaStream_t st);
int surs_launch_query_col(surs_ctx *ctx, const PointIO &io, int R1, int R2, int plane_lo, int nplanes, cudaStream_t st, int passes = 1);
int surs_launch_query_inc(surs_ctx *ctx, const PointIO &io, int R1, int R2, int plane_lo, int nplanes, cudaStream_t st);
// grid.cu
int surs_refine_select_impl(surs_ctx *ctx, const float *hr, const float *lr, int np, int R1, int R2, int64_t lin_base,
                            float level, float band, int64_t *idx_both, int64_t *idx_lr_only, int64_t *n_both, int64_t *n_lr_only, cudaStream_t st);
// mc.cu
int surs_mc_init_tables(surs_ctx *ctx);
