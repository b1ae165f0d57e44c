// Coarse-to-fine octree bookkeeping of lib/sdf.py:55-120 on the device.
//
//   select : grid_mask & dirty -> compacted node list (warp-ballot compaction), dirty cleared
//            (lib/sdf.py:70-77)
//   cells  : the interpolation loop (lib/sdf.py:81-117) as two data-parallel passes.
//            The sequential reference loop is order independent inside one level because the
//            only on-grid node a block fill overwrites is the cell's own origin and every
//            other cell that reads it has a smaller origin; so pass A takes
//            all decisions from the unmodified volumes and parks the fill value at the cell's
//            centre node (which only that cell touches), pass B broadcasts it over the block.
//
// `dirty` encoding: 1 = dirty; 0 = clean; 2|4 at a cell centre = "HR|LR block filled at some
// level" (clean).  Volumes are float64, as in the reference, so that mid-range fill values
// read back as corners at finer levels are bit identical.
#include "common.cuh"

namespace {

constexpr int SEL_THREADS = 256;
constexpr int SEL_PER_THREAD = 4;

__global__ void __launch_bounds__(SEL_THREADS) octree_select_kernel(uint8_t *__restrict__ dirty, int64_t *__restrict__ idx,
                                                                    unsigned long long *counter, int R1, int R2,
                                                                    int n1, int n2, int64_t ncand, int reso)
{
    __shared__ unsigned warp_cnt[SEL_THREADS / 32];
    __shared__ unsigned long long block_base;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t first = ((int64_t)blockIdx.x * SEL_THREADS * SEL_PER_THREAD) + (int64_t)warp * 32 * SEL_PER_THREAD;
    int64_t lin[SEL_PER_THREAD];
    unsigned ballots[SEL_PER_THREAD];
    unsigned mine = 0;
#pragma unroll
    for (int r = 0; r < SEL_PER_THREAD; ++r) {
        const int64_t c = first + r * 32 + lane;      // candidate id in C-order of the coarse lattice
        bool take = false;
        lin[r] = 0;
        if (c < ncand) {
            const int ck = (int)(c % n2);
            const int64_t t = c / n2;
            const int cj = (int)(t % n1);
            const int ci = (int)(t / n1);
            lin[r] = ((int64_t)ci * reso * R1 + (int64_t)cj * reso) * R2 + (int64_t)ck * reso;
            take = dirty[lin[r]] == 1;
        }
        ballots[r] = __ballot_sync(0xffffffffu, take);
        mine += __popc(ballots[r]);
    }
    if (lane == 0) warp_cnt[warp] = mine;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned tot = 0;
        for (int w = 0; w < SEL_THREADS / 32; ++w) { unsigned c = warp_cnt[w]; warp_cnt[w] = tot; tot += c; }
        block_base = tot ? atomicAdd(counter, (unsigned long long)tot) : 0ull;
    }
    __syncthreads();
    unsigned long long o = block_base + warp_cnt[warp];
#pragma unroll
    for (int r = 0; r < SEL_PER_THREAD; ++r) {
        if ((ballots[r] >> lane) & 1u) {
            idx[o + __popc(ballots[r] & ((1u << lane) - 1u))] = lin[r];
            dirty[lin[r]] = 0;                         // lib/sdf.py:77
        }
        o += __popc(ballots[r]);
    }
}

// select at the finest level (reso = 1: every node is a candidate) for volumes whose node count is a multiple of 4:
// one 32-bit load brings a thread's four consecutive dirty flags (the per-candidate byte loads of the generic kernel
// reach a fraction of the memory bandwidth), the list stays in ascending node order
__global__ void __launch_bounds__(SEL_THREADS) octree_select1_kernel(uint8_t *__restrict__ dirty, int64_t *__restrict__ idx,
                                                                     unsigned long long *counter, int64_t nquads)
{
    __shared__ unsigned warp_cnt[SEL_THREADS / 32];
    __shared__ unsigned long long block_base;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t q = (int64_t)blockIdx.x * SEL_THREADS + threadIdx.x;
    uint32_t flags = 0;
    if (q < nquads) flags = reinterpret_cast<const uint32_t *>(dirty)[q];
    unsigned take = 0;
#pragma unroll
    for (int b = 0; b < 4; ++b) take |= (((flags >> (8 * b)) & 0xffu) == 1u) ? (1u << b) : 0u;
    const unsigned mine = __popc(take);
    unsigned inc = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned x = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += x;
    }
    if (lane == 31) warp_cnt[warp] = inc;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned tot = 0;
        for (int w = 0; w < SEL_THREADS / 32; ++w) { unsigned c = warp_cnt[w]; warp_cnt[w] = tot; tot += c; }
        block_base = tot ? atomicAdd(counter, (unsigned long long)tot) : 0ull;
    }
    __syncthreads();
    if (!take) return;
    unsigned long long o = block_base + warp_cnt[warp] + inc - mine;
    uint32_t cleared = flags;
#pragma unroll
    for (int b = 0; b < 4; ++b)
        if ((take >> b) & 1u) {
            idx[o++] = 4 * q + b;
            cleared &= ~(0xffu << (8 * b));                            // lib/sdf.py:77: dirty[test] = False
        }
    reinterpret_cast<uint32_t *>(dirty)[q] = cleared;
}

// SURS_PREC_FP16R: nodes of a dense slab [np, R1, R2] whose one-pass value must be recomputed with split operands --
// every node that marching cubes at `level` can interpolate from after the refinement: the node or one of its six
// neighbours lies within `band` of the level (its inside / outside bit may still change), or its bit differs from a
// neighbour's.  The criterion is evaluated per volume: nodes the HR surface depends on go to `idx_both` (the HR MLP
// takes the LR prediction as an input, so both MLPs are re-evaluated), nodes only the LR surface depends on go to
// `idx_lr` (the LR MLP alone).  The lists hold global linear node indices (slab-relative index + lin_base).
__global__ void __launch_bounds__(SEL_THREADS) refine_select_kernel(const float *__restrict__ hr, const float *__restrict__ lr,
                                                                    int64_t *__restrict__ idx_both, int64_t *__restrict__ idx_lr,
                                                                    unsigned long long *counter, int np, int R1, int R2, int64_t lin_base,
                                                                    float level, float band)
{
    __shared__ unsigned warp_cnt[2][SEL_THREADS / 32];
    __shared__ unsigned long long block_base[2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t n = (int64_t)np * R1 * R2;
    const int64_t first = ((int64_t)blockIdx.x * SEL_THREADS * SEL_PER_THREAD) + (int64_t)warp * 32 * SEL_PER_THREAD;
    unsigned ballots[2][SEL_PER_THREAD];
    unsigned mine[2] = {0, 0};
    const int64_t s1 = R2, s0 = (int64_t)R1 * R2;
#pragma unroll
    for (int r = 0; r < SEL_PER_THREAD; ++r) {
        const int64_t c = first + r * 32 + lane;
        bool take[2] = {false, false};
        if (c < n) {
            const int k = (int)(c % R2);
            const int64_t t = c / R2;
            const int j = (int)(t % R1), i = (int)(t / R1);
            const int64_t nb[6] = {k > 0 ? c - 1 : c, k + 1 < R2 ? c + 1 : c, j > 0 ? c - s1 : c, j + 1 < R1 ? c + s1 : c,
                                   i > 0 ? c - s0 : c, i + 1 < np ? c + s0 : c};
#pragma unroll
            for (int v = 0; v < 2; ++v) {
                const float *vol = v ? lr : hr;
                const float x = vol[c];
                bool tk = fabsf(x - level) < band;
#pragma unroll
                for (int q = 0; q < 6; ++q) {
                    const float y = vol[nb[q]];
                    tk = tk || fabsf(y - level) < band || ((x > level) != (y > level));
                }
                take[v] = tk;
            }
        }
        ballots[0][r] = __ballot_sync(0xffffffffu, take[0]);
        ballots[1][r] = __ballot_sync(0xffffffffu, take[1] && !take[0]);
        mine[0] += __popc(ballots[0][r]);
        mine[1] += __popc(ballots[1][r]);
    }
    if (lane == 0) { warp_cnt[0][warp] = mine[0]; warp_cnt[1][warp] = mine[1]; }
    __syncthreads();
    if (threadIdx.x < 2) {
        unsigned tot = 0;
        for (int w = 0; w < SEL_THREADS / 32; ++w) { unsigned c = warp_cnt[threadIdx.x][w]; warp_cnt[threadIdx.x][w] = tot; tot += c; }
        block_base[threadIdx.x] = tot ? atomicAdd(counter + threadIdx.x, (unsigned long long)tot) : 0ull;
    }
    __syncthreads();
#pragma unroll
    for (int l = 0; l < 2; ++l) {
        int64_t *idx = l ? idx_lr : idx_both;
        unsigned long long o = block_base[l] + warp_cnt[l][warp];
#pragma unroll
        for (int r = 0; r < SEL_PER_THREAD; ++r) {
            if ((ballots[l][r] >> lane) & 1u) idx[o + __popc(ballots[l][r] & ((1u << lane) - 1u))] = lin_base + first + r * 32 + lane;
            o += __popc(ballots[l][r]);
        }
    }
}

struct Range { double lo, hi; };
__device__ __forceinline__ Range corner_range(const double *__restrict__ v, int64_t o, int64_t s0, int64_t s1, int64_t s2)
{
    Range r;
    r.lo = r.hi = v[o];
#pragma unroll
    for (int c = 1; c < 8; ++c) {
        double x = v[o + ((c & 4) ? s0 : 0) + ((c & 2) ? s1 : 0) + ((c & 1) ? s2 : 0)];
        r.lo = fmin(r.lo, x);
        r.hi = fmax(r.hi, x);
    }
    return r;
}

// pass A: one thread per cell (origins in range(0, R - reso, reso), lib/sdf.py:81-83)
// stats[0] += cells whose 16 corners were read, stats[1] / stats[2] += nodes the fill pass will write in HR / LR
__global__ void octree_decide_kernel(double *__restrict__ hr, double *__restrict__ lr, uint8_t *__restrict__ dirty,
                                     int R1, int R2, int m1, int m2, int64_t ncell, int reso, double threshold,
                                     unsigned long long *stats)
{
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = c < ncell && [&] {
        const int cz = (int)(c % m2);
        const int64_t t = c / m2;
        const int64_t o = ((int64_t)(t / m1) * reso * R1 + (int64_t)(t % m1) * reso) * R2 + (int64_t)cz * reso;
        return dirty[o + (int64_t)(reso / 2) * ((int64_t)R1 * R2 + R2 + 1)] == 1;
    }();
    const int nlive = __syncthreads_count(live);
    if (threadIdx.x == 0 && nlive) atomicAdd(stats, (unsigned long long)nlive);
    if (c >= ncell) return;
    const int cz = (int)(c % m2);
    const int64_t t = c / m2;
    const int cy = (int)(t % m1);
    const int cx = (int)(t / m1);
    const int64_t s2 = reso, s1 = (int64_t)reso * R2, s0 = (int64_t)reso * R1 * R2;
    const int64_t o = (int64_t)cx * s0 + (int64_t)cy * s1 + (int64_t)cz * s2;
    const int64_t centre = o + (int64_t)(reso / 2) * ((int64_t)R1 * R2 + R2 + 1);   // x + reso // 2, ...
    if (dirty[centre] != 1) return;                                  // lib/sdf.py:85
    const Range a = corner_range(hr, o, s0, s1, s2);
    const Range b = corner_range(lr, o, s0, s1, s2);
    unsigned code = 0;
    if (__dsub_rn(a.hi, a.lo) < threshold) { hr[centre] = __dadd_rn(a.hi, a.lo) / 2.0; code |= 2; }   // :97-101
    if (__dsub_rn(b.hi, b.lo) < threshold) { lr[centre] = __dadd_rn(b.hi, b.lo) / 2.0; code |= 4; }   // :113-117
    if (code) dirty[centre] = (uint8_t)code;
    const unsigned long long vox = (unsigned long long)reso * reso * reso - 1;
    if (code & 2) atomicAdd(stats + 1, vox);
    if (code & 4) atomicAdd(stats + 2, vox);
}

// pass B: one thread per node, coalesced along the last axis
__global__ void octree_fill_kernel(double *__restrict__ hr, double *__restrict__ lr, uint8_t *__restrict__ dirty,
                                   int R0, int R1, int R2, int m0, int m1, int m2, int reso)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y, i = blockIdx.z;
    if (k >= R2) return;
    const int cx = i / reso, cy = j / reso, cz = k / reso;
    if (cx >= m0 || cy >= m1 || cz >= m2) return;                    // outside every cell: last rows
    const int h = reso / 2;
    const int64_t centre = ((int64_t)(cx * reso + h) * R1 + (cy * reso + h)) * R2 + (cz * reso + h);
    const unsigned code = dirty[centre];
    if (code < 2) return;
    const int64_t lin = ((int64_t)i * R1 + j) * R2 + k;
    if (lin == centre) return;                                       // holds the value and the code
    if (code & 2) hr[lin] = hr[centre];
    if (code & 4) lr[lin] = lr[centre];
    dirty[lin] = 0;
}

// pass B, R2 % 4 == 0 and reso = 2 or a multiple of 4: one thread per FOUR consecutive nodes of a row (two cells for reso = 2, one cell
// for reso >= 4): a quarter of the threads, one centre-flag read per cell instead of per node, 32-byte stores
__global__ void __launch_bounds__(128) octree_fill4_kernel(double *__restrict__ hr, double *__restrict__ lr, uint8_t *__restrict__ dirty,
                                                           int R0, int R1, int R2, int m0, int m1, int m2, int reso)
{
    const int k0 = 4 * (blockIdx.x * blockDim.x + threadIdx.x);
    const int j = blockIdx.y, i = blockIdx.z;
    if (k0 >= R2) return;
    const int cx = i / reso, cy = j / reso;
    if (cx >= m0 || cy >= m1) return;
    const int h = reso / 2;
    const int64_t row = ((int64_t)i * R1 + j) * R2;
    const int64_t crow = ((int64_t)(cx * reso + h) * R1 + (cy * reso + h)) * R2;
    const int ncell = reso >= 4 ? 1 : 2;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        if (c >= ncell) break;
        const int k = k0 + 2 * c, cz = k / reso;                     // reso = 2: nodes (k0, k0+1) and (k0+2, k0+3); else all four
        if (cz >= m2) continue;
        const int64_t centre = crow + (cz * reso + h);
        const unsigned code = dirty[centre];
        if (code < 2) continue;
        const int n = reso >= 4 ? 4 : 2;
        const double vh = (code & 2) ? hr[centre] : 0.0, vl = (code & 4) ? lr[centre] : 0.0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (q >= n) break;
            const int64_t lin = row + k + q;
            if (lin == centre) continue;                             // holds the value and the code
            if (code & 2) hr[lin] = vh;
            if (code & 4) lr[lin] = vl;
            dirty[lin] = 0;
        }
    }
}

}  // namespace

int surs_octree_select_impl(surs_ctx *ctx, const int res[3], int reso, uint8_t *dirty, int64_t *idx,
                            int64_t *n_selected, cudaStream_t st)
{
    if (reso < 1) SURS_FAIL(ctx, "octree: bad level");
    const int n0 = (res[0] + reso - 1) / reso, n1 = (res[1] + reso - 1) / reso, n2 = (res[2] + reso - 1) / reso;
    const int64_t ncand = (int64_t)n0 * n1 * n2;
    SURS_CUDA(ctx, cudaMemsetAsync(ctx->counter, 0, sizeof(unsigned long long), st));
    if (reso == 1 && ncand % 4 == 0 && ((uintptr_t)dirty & 3) == 0) {
        const int64_t nquads = ncand / 4;
        octree_select1_kernel<<<(unsigned)((nquads + SEL_THREADS - 1) / SEL_THREADS), SEL_THREADS, 0, st>>>(dirty, idx, ctx->counter, nquads);
    } else {
        const int64_t per_block = SEL_THREADS * SEL_PER_THREAD;
        const int64_t blocks = (ncand + per_block - 1) / per_block;
        octree_select_kernel<<<(unsigned)blocks, SEL_THREADS, 0, st>>>(dirty, idx, ctx->counter, res[1], res[2], n1, n2, ncand, reso);
    }
    SURS_LAUNCH_CHECK(ctx, "octree_select_kernel");
    unsigned long long n = 0;
    SURS_CUDA(ctx, cudaMemcpyAsync(&n, ctx->counter, sizeof(n), cudaMemcpyDeviceToHost, st));
    SURS_CUDA(ctx, cudaStreamSynchronize(st));
    *n_selected = (int64_t)n;
    return 0;
}

int surs_refine_select_impl(surs_ctx *ctx, const float *hr, const float *lr, int np, int R1, int R2, int64_t lin_base,
                            float level, float band, int64_t *idx_both, int64_t *idx_lr_only, int64_t *n_both, int64_t *n_lr_only, cudaStream_t st)
{
    const int64_t n = (int64_t)np * R1 * R2;
    SURS_CUDA(ctx, cudaMemsetAsync(ctx->counter, 0, 2 * sizeof(unsigned long long), st));
    const int64_t per_block = SEL_THREADS * SEL_PER_THREAD;
    refine_select_kernel<<<(unsigned)((n + per_block - 1) / per_block), SEL_THREADS, 0, st>>>(hr, lr, idx_both, idx_lr_only, ctx->counter, np, R1, R2,
                                                                                               lin_base, level, band);
    SURS_LAUNCH_CHECK(ctx, "refine_select_kernel");
    unsigned long long c[2] = {0, 0};
    SURS_CUDA(ctx, cudaMemcpyAsync(c, ctx->counter, sizeof(c), cudaMemcpyDeviceToHost, st));
    SURS_CUDA(ctx, cudaStreamSynchronize(st));
    *n_both = (int64_t)c[0];
    *n_lr_only = (int64_t)c[1];
    return 0;
}

int surs_octree_cells_impl(surs_ctx *ctx, const int res[3], int reso, double threshold, double *sdf_hr,
                           double *sdf_lr, uint8_t *dirty, cudaStream_t st)
{
    if (reso < 2) return 0;
    // number of origins in range(0, R - reso, reso)
    int m[3];
    for (int a = 0; a < 3; ++a) m[a] = res[a] > reso ? (res[a] - reso + reso - 1) / reso : 0;
    const int64_t ncell = (int64_t)m[0] * m[1] * m[2];
    if (ncell == 0) return 0;
    octree_decide_kernel<<<(unsigned)((ncell + 127) / 128), 128, 0, st>>>(sdf_hr, sdf_lr, dirty, res[1], res[2],
                                                                         m[1], m[2], ncell, reso, threshold, ctx->counter + 16);
    SURS_LAUNCH_CHECK(ctx, "octree_decide_kernel");
    if (res[1] > 65535 || res[0] > 65535) SURS_FAIL(ctx, "octree: resolution too large");
    if (res[2] % 4 == 0 && (reso == 2 || reso % 4 == 0)) {
        dim3 grid((res[2] / 4 + 127) / 128, res[1], res[0]);
        octree_fill4_kernel<<<grid, 128, 0, st>>>(sdf_hr, sdf_lr, dirty, res[0], res[1], res[2], m[0], m[1], m[2], reso);
    } else {
        dim3 grid((res[2] + 127) / 128, res[1], res[0]);
        octree_fill_kernel<<<grid, 128, 0, st>>>(sdf_hr, sdf_lr, dirty, res[0], res[1], res[2], m[0], m[1], m[2], reso);
    }
    SURS_LAUNCH_CHECK(ctx, "octree_fill_kernel");
    return 0;
}
