/* surs.h -- C ABI of libsurs.so, the B200 (sm_100a) implementation of the SuRS
 * reconstruction hot path: point query (projection + bilinear feature indexing + the two
 * SurfaceClassifier MLPs), dense / octree grid evaluation and marching cubes.
 *
 * The reference (marcopesavento/Super-resolution-3D-Human-Shape-from-a-Single-Low-Resolution-Image)
 * is pure Python and has no FFI; the seam this library plugs into is the set of Python
 * callables used by lib/train_util.py:53-85 `gen_mesh` and lib/mesh_util.py:8-49
 * `reconstruction`.  Each entry point below names the reference code it replaces
 * (file:line relative to the reference root).  INTEGRATION.md shows the ctypes stub a
 * maintainer of the reference would add.
 *
 * Conventions
 *   - plain C types only; no torch / CUDA types in the signatures.  `stream` is a
 *     cudaStream_t passed as void* (NULL = legacy default stream).
 *   - pointers marked [dev] are device pointers valid on the context's device, [host] are
 *     host pointers.  Inputs are borrowed, outputs are caller allocated.
 *   - every function returns 0 on success, non-zero on failure; surs_last_error() then
 *     holds a message.  Nothing here falls back to a CPU path.
 *   - a context is bound to one device and used by one host thread at a time.
 *   - all calls are asynchronous on `stream` unless documented otherwise.
 */
#ifndef SURS_H
#define SURS_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct surs_ctx surs_ctx;

/* arithmetic used by the MLP chain */
enum {
    SURS_PREC_FP32 = 0,   /* CUDA-core fp32 FMA chain; agrees with the reference to ~1e-6     */
    SURS_PREC_FP16 = 1,   /* tcgen05 tensor cores, fp16 operands / fp32 accumulate, ONE pass: up to 1.5e-2 from the
                           * reference on sharp weights -- explicit opt-in, not a parity mode                     */
    SURS_PREC_FP16X3 = 2  /* tensor cores with split operands: every product as A_hi.W_hi + A_lo.W_hi + A_hi.W_lo
                           * (fp16 hi/lo pairs, fp32 accumulate), < 1e-4 from the reference's fp32 result.  Column-
                           * factored grids (surs_eval_grid / surs_eval_grid_octree without a transform and with
                           * calib[0][2] == calib[1][2] == 0) run at a third of the FP16 rate; every other point
                           * source (surs_query, transformed grids, sheared calibrations) goes through per-point
                           * tables on the same kernels (~45 M queries/s, 9x the SURS_PREC_FP32 kernel). */,
    SURS_PREC_FP16R = 3   /* "refined": SURS_PREC_FP16 on every node of a dense slab, then SURS_PREC_FP16X3 on the
                           * nodes the 0.5 iso-surface can depend on (the node or a 6-neighbour within the band of 0.5,
                           * or an inside / outside change to a neighbour) -- every inside / outside bit and every value
                           * marching cubes interpolates equals the FP16X3 result, at ~1.2x the FP16 time.  The band is
                           * verified on every call (surs_refine_stats).  Dense column-factored slabs only
                           * (surs_eval_grid); elsewhere identical to SURS_PREC_FP16X3.  The host layers' default. */
};

#define SURS_NUM_LAYERS 5

int surs_version(void);
/* nodes re-evaluated with split operands by the last surs_eval_grid(SURS_PREC_FP16R) of this context */
int64_t surs_refined_nodes(const surs_ctx *ctx);
/* ... how many of them only the LR surface depends on (those go through the LR MLP alone; the HR MLP needs the LR
 * prediction as an input, so nodes of the HR surface take both), and the run-time check of the band: max |one-pass -
 * split| over the re-evaluated values and the band.  When the maximum is not below 0.8 band the band is widened once
 * (to max_diff / 0.6; `attempts` = 2) and the selection repeated; if the check still fails the whole slab is
 * re-evaluated with SURS_PREC_FP16X3 (`fell_back`).  Any pointer may be NULL. */
int surs_refine_stats(const surs_ctx *ctx, int64_t *nodes, int64_t *nodes_lr_only, float *max_diff, float *band, int *fell_back,
                      int *attempts);

/* Lifetime.  `device` is a CUDA ordinal. */
int surs_create(surs_ctx **out, int device);
void surs_destroy(surs_ctx *ctx);
/* ctx may be NULL: returns the last error of a failed surs_create. */
const char *surs_last_error(const surs_ctx *ctx);

/* Replaces: the parameters of lib/model/SurfaceClassifier.py:7-43 as held by
 * SuRSNet.mlp_lr / SuRSNet.mlp_hr (lib/model/SuRSNet.py:67-78) after load_state_dict.
 * w_*[l] [dev]: conv{l}.weight as fp32 row-major [Cout, Cin] (Cin includes the skip input
 * dims[0] on residual layers, appended AFTER the previous activations, SurfaceClassifier.py:63-64);
 * b_*[l] [dev]: conv{l}.bias [Cout].  dims_* = opt.mlp_dim_{lr,hr} (6 ints),
 * res_layers = opt.mlp_res_layers_* (same list for both MLPs).
 * Supported: dims_lr = {321,1024,512,256,128,1}, dims_hr = {322,...}, res_layers = {2,3,4};
 * anything else returns an error (the Python shim then uses its torch path, never silently).
 * Synchronous with respect to the host (packs operand images for both precisions). */
int surs_set_weights(surs_ctx *ctx,
                     const float *const w_lr[SURS_NUM_LAYERS], const float *const b_lr[SURS_NUM_LAYERS],
                     const float *const w_hr[SURS_NUM_LAYERS], const float *const b_hr[SURS_NUM_LAYERS],
                     const int dims_lr[SURS_NUM_LAYERS + 1], const int dims_hr[SURS_NUM_LAYERS + 1],
                     const int *res_layers, int n_res, void *stream);

/* Replaces: SuRSNet.im_feat_list_lr[-1] / im_feat_list_hr[0] as consumed by lib/geometry.py:4-12
 * `index` (lib/model/SuRSNet.py:151,176).  f_lr [dev] NCHW fp32 [C_lr,H_lr,W_lr] (C_lr = 256),
 * f_hr [dev] [C_hr,H_hr,W_hr] (C_hr = 64), batch/view size 1.  Repacked to channels-last. */
int surs_set_features(surs_ctx *ctx, const float *f_lr, int C_lr, int H_lr, int W_lr,
                      const float *f_hr, int C_hr, int H_hr, int W_hr, void *stream);

/* Same maps from HOST memory (NCHW fp32; pinned memory makes the copies asynchronous), restricted to what a
 * caller that only samples image coordinates u in [u_lo, u_hi] needs: the pixel columns grid_sample touches for
 * that range (+ one pixel) are uploaded with one strided copy per map and repacked; the rest of the resident maps
 * is left as it was.  A slab-parallel rank (axis 0 = world x, u = calib[0][0] x + ...) uploads 1/N of the
 * 80-320 MiB this way.  Calls that would sample outside the covered range return an error (surs_query and
 * surs_eval_grid_octree need whole maps; surs_eval_grid checks its slab's range).  u_lo <= -1 and u_hi >= 1: whole maps. */
int surs_set_features_host(surs_ctx *ctx, const float *f_lr, int C_lr, int H_lr, int W_lr,
                           const float *f_hr, int C_hr, int H_hr, int W_hr, float u_lo, float u_hi, void *stream);

/* Projection variant used by every following surs_query* / surs_eval_grid* call of this context.
 * Replaces: the `projection_mode` of lib/model/SuRSNet.py:44-57 -- perspective = 0: lib/geometry.py:15-31 `orthogonal`
 * (default), 1: :34-48 `perspective` ((u, v) = (R p + t).xy / (R p + t).z) -- and the optional image-space
 * `transforms` argument of query_mr / query_sr (lib/geometry.py:27-30,43-46): transform [host] = 2x3 row-major
 * [scale | shift], (u, v) <- scale (u, v) + shift, applied before the in-image test; NULL = none.
 * Column-factored grids need the orthogonal projection (any transform is fine); perspective grids take the
 * generic kernels. */
int surs_set_projection(surs_ctx *ctx, int perspective, const float *transform);

/* Multi-view (opt.num_views > 1).  Replaces: lib/model/SurfaceClassifier.py:70-76 -- after layer 2 the activations
 * and the skip input are averaged over the views of a subject -- together with the per-view projection / indexing of
 * lib/model/SuRSNet.py:131-187.  f_lr [dev] NCHW fp32 [V,256,H,W], f_hr [dev] [V,64,H,W]; pts [dev] fp32 [V,3,N]
 * (the reference repeats the same points per view, lib/mesh_util.py:22), calibs [host] V x (upper 3x4, row-major).
 * pred_hr / pred_lr [dev] fp32 [V,N]: mask_v x sigmoid(...), as the reference's [V,1,N] predictions.
 * Runs the fp32 CUDA-core kernel (exact mode) whatever the precision of the single-view calls. */
int surs_set_features_views(surs_ctx *ctx, int n_views, const float *f_lr, int C_lr, int H_lr, int W_lr,
                            const float *f_hr, int C_hr, int H_hr, int W_hr, void *stream);
int surs_query_views(surs_ctx *ctx, const float *pts, int64_t n, const float *calibs, float z_num, float z_den,
                     float *pred_hr, float *pred_lr, void *stream);

/* Replaces: SuRSNet.query_mr + query_sr + get_preds (lib/model/SuRSNet.py:131-187,
 * lib/model/BaseSuRSNet.py:80-85) with lib/geometry.py:15-31 `orthogonal`, :4-12 `index`,
 * lib/model/DepthNormalizer.py:18 and lib/model/SurfaceClassifier.py:45-81 fused.
 * pts [dev] fp32 [3,N] (row 0 = x, ...), calib [host] the upper 3x4 of the calibration
 * matrix, row-major; z_feat = z * z_num / z_den with z_num = opt.loadSize // 2,
 * z_den = opt.z_size.  pred_hr / pred_lr [dev] fp32 [N]. */
int surs_query(surs_ctx *ctx, const float *pts, int64_t n, const float calib[12],
               float z_num, float z_den, int precision,
               float *pred_hr, float *pred_lr, void *stream);

/* Same with HOST buffers: copies pts in, runs, copies both predictions out and
 * synchronises the stream (what lib/mesh_util.py:20-28 `eval_func` does per chunk). */
int surs_query_host(surs_ctx *ctx, const float *pts_host, int64_t n, const float calib[12],
                    float z_num, float z_den, int precision,
                    float *pred_hr_host, float *pred_lr_host, void *stream);

/* Replaces: lib/sdf.py:4-29 `create_grid` + :48-52 `eval_grid` (+ :32-45 `batch_eval`) with
 * the eval_func of lib/mesh_util.py:20-28.  Grid nodes are generated on the device:
 * node (i,j,k) = b_min + (i,j,k) * (b_max - b_min) / res (float64, optional 4x4 row-major
 * `transform` [host] applied in float64), then cast to fp32 as lib/mesh_util.py:24 does.
 * Evaluates the planes [plane_lo, plane_hi) of array axis 0 (a slab; 0,res[0] = everything);
 * sdf_hr / sdf_lr [dev] fp32 [(plane_hi - plane_lo), res[1], res[2]] C-order. */
int surs_eval_grid(surs_ctx *ctx, const int res[3], const double b_min[3], const double b_max[3],
                   const double *transform, const float calib[12], float z_num, float z_den,
                   int precision, int plane_lo, int plane_hi,
                   float *sdf_hr, float *sdf_lr, void *stream);

/* Replaces: lib/sdf.py:55-120 `eval_grid_octree` (same eval_func).  Reproduces the
 * reference semantics exactly: shared `dirty` mask, range threshold, origin-node overwrite,
 * skipped last cell row, zero holes.  sdf_hr / sdf_lr [dev] float64 [res0,res1,res2]
 * (float64 because filled mid-range values are float64 in the reference and are read back
 * as corners at finer levels).  n_evaluated [host, may be NULL] receives the number of
 * network evaluations.  Synchronises the stream once per level. */
int surs_eval_grid_octree(surs_ctx *ctx, const int res[3], const double b_min[3], const double b_max[3],
                          const double *transform, const float calib[12], float z_num, float z_den,
                          int precision, int init_resolution, double threshold,
                          double *sdf_hr, double *sdf_lr, int64_t *n_evaluated, void *stream);

/* Phase statistics of the last surs_eval_grid_octree on this context (synchronises on its events):
 *   ms[5]     = device time of {volume / dirty initialisation, per-column table, select, network evaluation, cell pass}
 *   counts[6] = {lattice candidates read by select, nodes evaluated, cells visited by the cell pass,
 *                cells whose 16 corners were read, nodes block-filled in HR, nodes block-filled in LR}
 * -- the terms of the octree bookkeeping's HBM roofline (bench.py's `roofline_octree`). */
int surs_octree_stats(surs_ctx *ctx, float ms[5], int64_t counts[6]);

/* Building blocks of the octree for callers that bring their own eval_func (the generic
 * lib/sdf.py path): `surs_octree_select` marks grid_mask & dirty nodes of level `reso`
 * (lib/sdf.py:70-74) and writes their linear indices (C-order) to idx [dev, capacity
 * res0*res1*res2 / reso^3], count to n_selected [host] (synchronises), and clears their dirty
 * flag (:77); `surs_octree_cells` is the interpolation loop lib/sdf.py:81-117 for one level.
 * dirty [dev] uint8 [res0,res1,res2] (1 = dirty), caller initialises it to 1. */
int surs_octree_select(surs_ctx *ctx, const int res[3], int reso, uint8_t *dirty,
                       int64_t *idx, int64_t *n_selected, void *stream);
int surs_octree_cells(surs_ctx *ctx, const int res[3], int reso, double threshold,
                      double *sdf_hr, double *sdf_lr, uint8_t *dirty, void *stream);

/* Replaces: skimage.measure.marching_cubes_lewiner(volume, level) as called at
 * lib/mesh_util.py:40,45 (scikit-image 0.17.2, defaults), followed by the world transform
 * lib/mesh_util.py:42-43.  Two phases because the output size is data dependent:
 *   surs_mc_count  classifies the cells of vol [dev] fp32 [res0,res1,res2] and returns the
 *                  vertex / face counts [host] (synchronises the stream);
 *   surs_mc_emit   writes verts [dev] fp32 [V,3] (index coordinates, (axis0,axis1,axis2)),
 *                  faces [dev] int32 [F,3], normals [dev] fp32 [V,3], values [dev] fp32 [V]
 *                  for the volume of the preceding surs_mc_count on this context; any of
 *                  normals / values / verts_world may be NULL.  If mat [host] (upper 3x4 of
 *                  the 4x4 grid matrix, row-major, float64) is given, verts_world [dev]
 *                  float64 [V,3] receives mat[:3,:3] @ v + mat[:3,3].
 *                  = surs_mc_emit_verts(offset 0, no seam) + surs_mc_emit_faces.
 * Vertex order = order of first use scanning cells with axis 0 outermost; faces in scan
 * order; right-handed normals point towards increasing values.
 *
 * Multi-GPU (the volume is a slab of array axis 0 with one halo plane on top): a mesh that
 * is identical to the single-GPU one needs the vertices on the plane shared by two slabs to
 * exist once.  They belong to the LOWER slab (first use in scan order).  The upper slab is
 * counted with SURS_MC_LOWER_FOREIGN: edges lying in its plane 0 are not emitted and its
 * faces reference them through seam_in [dev] int32 [2,res1,res2] (global ids; axis-1 edges
 * then axis-2 edges), which is the lower slab's seam_out [dev] int32 [2,res1,res2] (ids of
 * the vertices lying in its LAST plane, -1 where there is none).  vert_id_offset (the sum of
 * the vertex counts of all lower slabs) is added to every id written to faces / seam_out;
 * plane_offset (the slab's first plane in the full grid) is added to the axis-0 coordinate
 * before interpolation, so positions are bit identical to the single-volume result.
 * So: count everywhere -> exchange counts -> emit_verts everywhere -> pass seam_out up ->
 * emit_faces everywhere -> concatenate in rank order.
 * n_ambiguous [host, may be NULL]: number of cells with an ambiguous face (the part of the
 * Lewiner algorithm whose parity with scikit-image is unpinned).
 * surs_mc_interior_stats: cells of the last surs_mc_count whose face decisions left an INTERIOR ambiguity (two
 * same-sign regions that the trilinear interpolant may join by a tunnel: Lewiner's cases 4, 6, 7, 10, 12, 13) and
 * how many of them took the tunnel triangulation (csrc/gen_mc_tables.py).  Synchronises. */
enum { SURS_MC_LOWER_FOREIGN = 1 };
int surs_mc_count(surs_ctx *ctx, const float *vol, const int res[3], float level, int flags,
                  int64_t *n_verts, int64_t *n_faces, int64_t *n_ambiguous, void *stream);
/* surs_mc_count_f64: the same for a float64 volume (lib/sdf.py keeps the octree volumes in float64; skimage casts its
 * input to float32): the pass that takes the inside / outside bits also writes the float32 copy into vol32 [dev, caller
 * allocated, borrowed until the emit calls ran] -- no separate cast pass.
 * surs_mc_value_range: minimum and maximum of the last counted volume, a by-product of that same pass; it is what
 * skimage compares the level with ("Surface level must be within volume data range") -- no separate reduction pass. */
int surs_mc_count_f64(surs_ctx *ctx, const double *vol64, float *vol32, const int res[3], float level, int flags,
                      int64_t *n_verts, int64_t *n_faces, int64_t *n_ambiguous, void *stream);
int surs_mc_value_range(surs_ctx *ctx, float *vmin, float *vmax);
int surs_mc_interior_stats(surs_ctx *ctx, int64_t *n_interior_ambiguous, int64_t *n_tunnels);
int surs_mc_emit(surs_ctx *ctx, const double *mat, float *verts, double *verts_world,
                 int32_t *faces, float *normals, float *values, void *stream);
int surs_mc_emit_verts(surs_ctx *ctx, const double *mat, float *verts, double *verts_world,
                       float *normals, float *values, int64_t vert_id_offset, int plane_offset,
                       int32_t *seam_out, void *stream);
int surs_mc_emit_faces(surs_ctx *ctx, int32_t *faces, const int32_t *seam_in, void *stream);
/* Seam edges for which the last surs_mc_emit_faces found seam_in == -1 (the two slabs disagree about an inside /
 * outside bit of the shared plane; the face then holds -1).  0 on every consistent input.  Synchronises. */
int64_t surs_mc_seam_violations(surs_ctx *ctx);

/* Peer arenas (multi-GPU, one process per GPU on one NVLink / NVSwitch box): `surs_arena_create` allocates `bytes` of
 * device memory on this context's device and returns its CUDA IPC handle; another process passes the 64 handle bytes to
 * `surs_arena_open` and gets a pointer that is valid on ITS device and backed by the owner's HBM (peer access over
 * NVLink).  Every output pointer of surs_mc_emit_verts / surs_mc_emit_faces may point into such a mapping: the
 * emission then writes each rank's part of the mesh directly at its offset in the gathering rank's buffers -- the gather
 * of the reference-free multi-GPU path fused into the kernels, no collective for the payload.
 * `verts` may be NULL in surs_mc_emit_verts when verts_world (with mat) is requested. */
int surs_arena_create(surs_ctx *ctx, int64_t bytes, void **dev_ptr, unsigned char handle[64]);
int surs_arena_open(surs_ctx *ctx, const unsigned char handle[64], void **peer_ptr);
int surs_arena_close(surs_ctx *ctx, void *peer_ptr);
int surs_arena_destroy(surs_ctx *ctx, void *dev_ptr);

/* float64 -> float32 cast of a volume (what skimage does to its input); n elements. */
int surs_cast_f64_f32(surs_ctx *ctx, const double *src, float *dst, int64_t n, void *stream);

/* Replaces: lib/mesh_util.py:53-61 `save_obj_mesh` (byte-identical text: 'v %.4f %.4f %.4f',
 * then 1-based 'f a c b').  verts [host] float64 [V,3], faces [host] int32 [F,3]. */
int surs_save_obj_mesh(const char *path, const double *verts, int64_t n_verts,
                       const int32_t *faces, int64_t n_faces);

/* Debug / unit test: D[128,N] = A[128,K] . B[N,K]^T (fp32 in/out on the device, operands
 * rounded to fp16) through the same UMMA descriptors, swizzled layouts and TMEM loads the
 * query kernel uses.  tail16 != 0: only the first 16 columns of the last 64-wide K block
 * take part.  N multiple of 16 in [16,256], K <= 192. */
int surs_selftest_umma(surs_ctx *ctx, const float *A, const float *B, int N, int K, int tail16,
                       float *D, void *stream);

/* Same for the CTA-pair path (tcgen05.mma.cta_group::2, one cluster of two CTAs):
 * D[256,N] = A[256,K] . B[N,K]^T, N multiple of 32 in [32,256], K <= 192. */
int surs_selftest_umma2(surs_ctx *ctx, const float *A, const float *B, int N, int K, float *D, void *stream);

/* Tensor-pipe rate probe (scripts/umma_rate.py): every issuing thread of a `grid`-CTA launch times
 * reps x 4 MMAs of 128 x 256 x 16 (pair = 0, cta_group::1) or 256 x 256 x 16 (pair = 1, clusters of two CTAs,
 * each holding half of the B operand); cycles[b] (device, `grid` entries) receives the clock cycles of CTA b's
 * issuer (pair = 1: leaders only).  Measured on B200: 128.0 cycles per MMA in both modes. */
int surs_selftest_umma_rate(surs_ctx *ctx, int pair, int grid, int reps, unsigned long long *cycles, void *stream);

/* Counters for bench.py's "gpu_launches": kernels launched by this context so far. */
int64_t surs_launch_count(const surs_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* SURS_H */
