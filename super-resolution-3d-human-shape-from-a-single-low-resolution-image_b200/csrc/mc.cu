// Marching cubes on the device: replaces skimage.measure.marching_cubes_lewiner(volume, 0.5)
// as called at lib/mesh_util.py:40,45 plus the world transform of lib/mesh_util.py:42-43.
//
// Output order is the sequential algorithm's (vertices numbered by first use while scanning
// cells with axis 0 outermost and, inside a cell, its triangle list; faces in scan order), but
// derived without any serial dependency:
//   * every cube edge is OWNED by the first cell in scan order that contains it, i.e. the cell
//     for which the edge sits at offset 1 in both transverse axes (or at offset 0 where the cell
//     index is 0); a cell's new vertices are its owned, sign-changing edges (+ its centre vertex),
//   * inside the cell they are ordered by first appearance in the cell's triangle list
//     (mc_edge_rank, generated with the case tables),
//   * prefix sums over cells in scan order give the global numbering.
// Kernels: count (per-block totals) -> scan of block totals -> emit vertices (+ edge->id map)
// -> emit faces.  Case tables: csrc/mc_tables.h (derivation in oracle/gen_mc_tables.py).
#include "common.cuh"
#include "mc_tables.h"

#include <float.h>
#include <stdlib.h>

namespace {

constexpr int MC_THREADS = 256;

struct McTables {
    const uint8_t *amb_mask;     // [256]
    const uint16_t *var_base;    // [256]
    const uint8_t *ntri;         // [E]
    const uint16_t *tri_off;     // [E]
    const uint8_t *tri_edges;    // [3*T]
    const uint8_t *edge_rank;    // [E][13]
};

__constant__ uint8_t c_corner_off[24];
__constant__ uint8_t c_edge_corner[24];
__constant__ uint8_t c_edge_axis[12];
__constant__ uint8_t c_edge_base[36];
__constant__ uint8_t c_face_corner[24];

struct McParams {
    const float *vol;
    int R0, R1, R2;
    double level;
    int lower_foreign;
    McTables tb;
};

struct Cell {
    int entry;          // index into the (case, decider bits) tables; -1: no cell / no surface
    unsigned owned;     // 13-bit mask of vertex slots this cell creates
    int nv, nt;
    int ambiguous;
    double d[8];
};

__device__ __forceinline__ int64_t node_lin(const McParams &p, int i, int j, int k) { return ((int64_t)i * p.R1 + j) * p.R2 + k; }

__device__ __forceinline__ void classify_vals(const McParams &p, int i, int j, int k, const float (&val)[8], Cell &c);

__device__ __forceinline__ void classify(const McParams &p, int i, int j, int k, Cell &c)
{
    c.entry = -1; c.nv = 0; c.nt = 0; c.owned = 0; c.ambiguous = 0;
    if (i >= p.R0 - 1 || j >= p.R1 - 1 || k >= p.R2 - 1) return;
    float val[8];
#pragma unroll
    for (int q = 0; q < 8; ++q)
        val[q] = __ldg(p.vol + node_lin(p, i + c_corner_off[3 * q], j + c_corner_off[3 * q + 1], k + c_corner_off[3 * q + 2]));
    classify_vals(p, i, j, k, val, c);
}

// val[q] = volume value at corner q of cell (i,j,k) (which must be a valid cell)
__device__ __forceinline__ void classify_vals(const McParams &p, int i, int j, int k, const float (&val)[8], Cell &c)
{
    c.entry = -1; c.nv = 0; c.nt = 0; c.owned = 0; c.ambiguous = 0;
    unsigned cas = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        c.d[q] = __dsub_rn((double)val[q], p.level);
        if (c.d[q] > 0.0) cas |= 1u << q;
    }
    if (cas == 0 || cas == 255) return;
    const unsigned amb = p.tb.amb_mask[cas];
    unsigned var = 0, nb = 0;
    if (amb) {
        c.ambiguous = 1;
        for (int f = 0; f < 6; ++f) {
            if (!((amb >> f) & 1u)) continue;
            const uint8_t *fc = &c_face_corner[4 * f];
            const double p02 = __dmul_rn(c.d[fc[0]], c.d[fc[2]]);
            const double p13 = __dmul_rn(c.d[fc[1]], c.d[fc[3]]);
            const unsigned connect = (c.d[fc[0]] > 0.0) ? (p02 > p13) : (p13 > p02);
            var |= connect << nb;
            ++nb;
        }
    }
    c.entry = (int)p.tb.var_base[cas] + (int)var;
    c.nt = p.tb.ntri[c.entry];
    // ownership: offset 1 in both transverse axes, or offset 0 where the cell index is 0
    const int idx[3] = {i, j, k};
    const uint8_t *rank = p.tb.edge_rank + 13 * c.entry;
    unsigned owned = 0;
#pragma unroll
    for (int e = 0; e < 12; ++e) {
        if (rank[e] == 255) continue;
        const int a = c_edge_axis[e];
        bool own = true;
#pragma unroll
        for (int dd = 0; dd < 3; ++dd)
            if (dd != a) own = own && (c_edge_base[3 * e + dd] == 1 || idx[dd] == 0);
        if (p.lower_foreign && a != 0 && c_edge_base[3 * e] == 0 && i == 0) own = false;
        if (own) owned |= 1u << e;
    }
    if (rank[12] != 255) owned |= 1u << 12;
    c.owned = owned;
    c.nv = __popc(owned);
}

// exclusive scan of (a, b) over the block in thread order; returns totals through ta / tb
__device__ __forceinline__ void block_scan2(unsigned a, unsigned b, unsigned &ea, unsigned &eb, unsigned &ta, unsigned &tb)
{
    __shared__ unsigned wa[MC_THREADS / 32], wb[MC_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned ia = a, ib = b;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        unsigned xa = __shfl_up_sync(0xffffffffu, ia, o), xb = __shfl_up_sync(0xffffffffu, ib, o);
        if (lane >= o) { ia += xa; ib += xb; }
    }
    if (lane == 31) { wa[warp] = ia; wb[warp] = ib; }
    __syncthreads();
    unsigned oa = 0, ob = 0, sa = 0, sb = 0;
#pragma unroll
    for (int w = 0; w < MC_THREADS / 32; ++w) {
        if (w == warp) { oa = sa; ob = sb; }
        sa += wa[w]; sb += wb[w];
    }
    ea = oa + ia - a; eb = ob + ib - b; ta = sa; tb = sb;
    __syncthreads();
}

__device__ __forceinline__ void cell_coords(const McParams &p, int64_t lin, int &i, int &j, int &k)
{
    k = (int)(lin % p.R2);
    const int64_t t = lin / p.R2;
    j = (int)(t % p.R1);
    i = (int)(t / p.R1);
}

__global__ void __launch_bounds__(MC_THREADS) mc_count_kernel(McParams p, int64_t nnode, uint2 *block_tot, unsigned long long *n_amb)
{
    const int64_t lin = (int64_t)blockIdx.x * MC_THREADS + threadIdx.x;
    Cell c;
    c.nv = c.nt = 0; c.ambiguous = 0;
    if (lin < nnode) {
        int i, j, k;
        cell_coords(p, lin, i, j, k);
        classify(p, i, j, k, c);
    }
    unsigned ea, eb, ta, tb;
    block_scan2((unsigned)c.nv, (unsigned)c.nt, ea, eb, ta, tb);
    const unsigned amb = __syncthreads_count(c.ambiguous);
    if (threadIdx.x == 0) {
        block_tot[blockIdx.x] = make_uint2(ta, tb);
        if (amb) atomicAdd(n_amb, (unsigned long long)amb);
    }
}

// single-CTA exclusive scan of the per-block totals (64-bit running sums kept in the totals only)
__global__ void __launch_bounds__(1024) mc_scan_blocks_kernel(uint2 *block_tot, int64_t nblocks, unsigned long long *totals)
{
    __shared__ unsigned long long wsum_a[32], wsum_b[32];
    __shared__ unsigned long long carry_a, carry_b;
    if (threadIdx.x == 0) { carry_a = 0; carry_b = 0; }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t base = 0; base < nblocks; base += 1024) {
        const int64_t idx = base + threadIdx.x;
        uint2 v = idx < nblocks ? block_tot[idx] : make_uint2(0, 0);
        unsigned long long a = v.x, b = v.y;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned long long xa = __shfl_up_sync(0xffffffffu, a, o), xb = __shfl_up_sync(0xffffffffu, b, o);
            if (lane >= o) { a += xa; b += xb; }
        }
        if (lane == 31) { wsum_a[warp] = a; wsum_b[warp] = b; }
        __syncthreads();
        unsigned long long oa = carry_a, ob = carry_b;
        for (int w = 0; w < warp; ++w) { oa += wsum_a[w]; ob += wsum_b[w]; }
        if (idx < nblocks) block_tot[idx] = make_uint2((unsigned)(oa + a - v.x), (unsigned)(ob + b - v.y));
        __syncthreads();
        if (threadIdx.x == 1023) { carry_a = oa + a; carry_b = ob + b; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { totals[0] = carry_a; totals[1] = carry_b; }
}

// position along an edge whose lower end sits at coordinate `base` (oracle/mc_oracle.c edge_point)
__device__ __forceinline__ double edge_point(double base, double da, double db)
{
    const double wa = __ddiv_rn(1.0, __dadd_rn((double)FLT_EPSILON, fabs(da)));
    const double wb = __ddiv_rn(1.0, __dadd_rn((double)FLT_EPSILON, fabs(db)));
    return __ddiv_rn(__dadd_rn(__dmul_rn(base, wa), __dmul_rn(__dadd_rn(base, 1.0), wb)), __dadd_rn(wa, wb));
}

__device__ __forceinline__ double node_grad(const McParams &p, int i, int j, int k, int axis)
{
    const int n = axis == 0 ? p.R0 : (axis == 1 ? p.R1 : p.R2);
    const int q = axis == 0 ? i : (axis == 1 ? j : k);
    const int lo = q > 0 ? q - 1 : q, hi = q < n - 1 ? q + 1 : q;
    int il = i, jl = j, kl = k, ih = i, jh = j, kh = k;
    if (axis == 0) { il = lo; ih = hi; } else if (axis == 1) { jl = lo; jh = hi; } else { kl = lo; kh = hi; }
    const double dv = (double)__ldg(p.vol + node_lin(p, ih, jh, kh)) - (double)__ldg(p.vol + node_lin(p, il, jl, kl));
    return (hi - lo) == 2 ? dv * 0.5 : dv;
}

struct McOut {
    float *verts;
    double *verts_world;
    float *normals;
    float *values;
    double mat[12];
    int has_mat;
    int32_t *vid;          // [nnode][3] edge -> global vertex id
    int64_t id_offset;
    int plane_offset;      // index of the volume's plane 0 along axis 0 in the full grid (slabs)
};

__device__ __forceinline__ void write_vertex(const McOut &o, int64_t v, const double pos[3], const double g[3], double value)
{
    const float fx = (float)pos[0], fy = (float)pos[1], fz = (float)pos[2];
    o.verts[3 * v] = fx; o.verts[3 * v + 1] = fy; o.verts[3 * v + 2] = fz;
    if (o.verts_world && o.has_mat) {
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            double acc = __dmul_rn(o.mat[4 * r], (double)fx);
            acc = __fma_rn(o.mat[4 * r + 1], (double)fy, acc);
            acc = __fma_rn(o.mat[4 * r + 2], (double)fz, acc);
            o.verts_world[3 * v + r] = __dadd_rn(acc, o.mat[4 * r + 3]);
        }
    }
    if (o.normals) {
        const double nn = sqrt(g[0] * g[0] + g[1] * g[1] + g[2] * g[2]);
#pragma unroll
        for (int a = 0; a < 3; ++a) o.normals[3 * v + a] = nn > 0.0 ? (float)(g[a] / nn) : 0.0f;
    }
    if (o.values) o.values[v] = (float)value;
}

// all vertices created by cell (i,j,k); vbase = index of its first vertex in this volume's list
__device__ __forceinline__ void emit_cell_verts(const McParams &p, const McOut &o, int i, int j, int k, const Cell &c, int64_t vbase)
{
    const uint8_t *rank = p.tb.edge_rank + 13 * c.entry;
#pragma unroll 1
    for (int e = 0; e < 12; ++e) {
        if (!((c.owned >> e) & 1u)) continue;
        int r = 0;
        for (int e2 = 0; e2 < 13; ++e2)
            if (((c.owned >> e2) & 1u) && rank[e2] < rank[e]) ++r;
        const int64_t v = vbase + r;
        const int axis = c_edge_axis[e];
        const int bi = i + c_edge_base[3 * e], bj = j + c_edge_base[3 * e + 1], bk = k + c_edge_base[3 * e + 2];
        const int ei = bi + (axis == 0), ej = bj + (axis == 1), ek = bk + (axis == 2);
        const int ca = c_edge_corner[2 * e], cb = c_edge_corner[2 * e + 1];
        const int lo = c_corner_off[3 * ca + axis] == 0 ? ca : cb, hi = lo == ca ? cb : ca;
        double pos[3] = {(double)(bi + o.plane_offset), (double)bj, (double)bk};
        const double base = pos[axis];
        const double x = edge_point(base, c.d[lo], c.d[hi]);
        pos[axis] = x;
        double g[3] = {0.0, 0.0, 0.0};
        if (o.normals) {
            const double tt = x - base;
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const double ga = node_grad(p, bi, bj, bk, a), gb = node_grad(p, ei, ej, ek, a);
                g[a] = ga + tt * (gb - ga);
            }
        }
        const double va = (double)__ldg(p.vol + node_lin(p, bi, bj, bk)), vb = (double)__ldg(p.vol + node_lin(p, ei, ej, ek));
        write_vertex(o, v, pos, g, va > vb ? va : vb);
        o.vid[3 * node_lin(p, bi, bj, bk) + axis] = (int32_t)(v + o.id_offset);
    }
    if ((c.owned >> 12) & 1u) {                  // centre vertex: mean of the cell's edge vertices
        int r = 0;
        for (int e2 = 0; e2 < 12; ++e2)
            if (((c.owned >> e2) & 1u) && rank[e2] < rank[12]) ++r;
        double s[3] = {0.0, 0.0, 0.0};
        int n = 0;
        for (int e = 0; e < 12; ++e) {
            const int ca = c_edge_corner[2 * e], cb = c_edge_corner[2 * e + 1];
            if ((c.d[ca] > 0.0) == (c.d[cb] > 0.0)) continue;
            const int axis = c_edge_axis[e];
            const int lo = c_corner_off[3 * ca + axis] == 0 ? ca : cb, hi = lo == ca ? cb : ca;
            double q[3] = {(double)(i + o.plane_offset + c_edge_base[3 * e]), (double)(j + c_edge_base[3 * e + 1]), (double)(k + c_edge_base[3 * e + 2])};
            q[axis] = edge_point(q[axis], c.d[lo], c.d[hi]);
            s[0] = __dadd_rn(s[0], q[0]); s[1] = __dadd_rn(s[1], q[1]); s[2] = __dadd_rn(s[2], q[2]);
            ++n;
        }
        double pos[3] = {__ddiv_rn(s[0], (double)n), __ddiv_rn(s[1], (double)n), __ddiv_rn(s[2], (double)n)};
        double g[3] = {0.0, 0.0, 0.0}, vmax = -INFINITY;
        for (int q = 0; q < 8; ++q) {
            const int ci = i + c_corner_off[3 * q], cj = j + c_corner_off[3 * q + 1], ck = k + c_corner_off[3 * q + 2];
            if (o.normals)
                for (int a = 0; a < 3; ++a) g[a] += node_grad(p, ci, cj, ck, a);
            vmax = fmax(vmax, (double)__ldg(p.vol + node_lin(p, ci, cj, ck)));
        }
        write_vertex(o, vbase + r, pos, g, vmax);
    }
}

__global__ void __launch_bounds__(MC_THREADS) mc_emit_verts_kernel(McParams p, int64_t nnode, const uint2 *block_prefix, McOut o)
{
    const int64_t lin = (int64_t)blockIdx.x * MC_THREADS + threadIdx.x;
    Cell c;
    c.nv = c.nt = 0; c.entry = -1;
    int i = 0, j = 0, k = 0;
    if (lin < nnode) {
        cell_coords(p, lin, i, j, k);
        classify(p, i, j, k, c);
    }
    unsigned ea, eb, ta, tb;
    block_scan2((unsigned)c.nv, 0u, ea, eb, ta, tb);
    if (c.nv == 0) return;
    emit_cell_verts(p, o, i, j, k, c, (int64_t)block_prefix[blockIdx.x].x + ea);
}

// all faces of cell (i,j,k); vbase / fbase = its first vertex / face in this volume's lists
__device__ __forceinline__ void emit_cell_faces(const McParams &p, int i, int j, int k, const Cell &c, int64_t vbase, int64_t fbase,
                                                const int32_t *__restrict__ vid, const int32_t *__restrict__ seam_in,
                                                int64_t id_offset, int32_t *__restrict__ faces)
{
    const uint8_t *rank = p.tb.edge_rank + 13 * c.entry;
    const uint8_t *te = p.tb.tri_edges + 3 * (int)p.tb.tri_off[c.entry];
    int32_t centre_id = -1;
    if ((c.owned >> 12) & 1u) {
        int r = 0;
        for (int e2 = 0; e2 < 12; ++e2)
            if (((c.owned >> e2) & 1u) && rank[e2] < rank[12]) ++r;
        centre_id = (int32_t)(vbase + r + id_offset);
    }
    for (int t = 0; t < c.nt; ++t) {
        int32_t tri[3];
#pragma unroll
        for (int s = 0; s < 3; ++s) {
            const int e = te[3 * t + s];
            if (e == MC_CENTRE) { tri[s] = centre_id; continue; }
            const int axis = c_edge_axis[e];
            const int bi = i + c_edge_base[3 * e], bj = j + c_edge_base[3 * e + 1], bk = k + c_edge_base[3 * e + 2];
            if (p.lower_foreign && bi == 0 && axis != 0)
                tri[s] = seam_in[((int64_t)(axis - 1) * p.R1 + bj) * p.R2 + bk];
            else
                tri[s] = vid[3 * node_lin(p, bi, bj, bk) + axis];
        }
        faces[3 * (fbase + t)] = tri[0];
        faces[3 * (fbase + t) + 1] = tri[1];
        faces[3 * (fbase + t) + 2] = tri[2];
    }
}

__global__ void __launch_bounds__(MC_THREADS) mc_emit_faces_kernel(McParams p, int64_t nnode, const uint2 *block_prefix,
                                                                   const int32_t *__restrict__ vid, const int32_t *__restrict__ seam_in,
                                                                   int64_t id_offset, int32_t *__restrict__ faces)
{
    const int64_t lin = (int64_t)blockIdx.x * MC_THREADS + threadIdx.x;
    Cell c;
    c.nv = c.nt = 0; c.entry = -1;
    int i = 0, j = 0, k = 0;
    if (lin < nnode) {
        cell_coords(p, lin, i, j, k);
        classify(p, i, j, k, c);
    }
    unsigned ea, eb, ta, tb;
    block_scan2((unsigned)c.nv, (unsigned)c.nt, ea, eb, ta, tb);
    if (c.nt == 0) return;
    const uint2 pre = block_prefix[blockIdx.x];
    emit_cell_faces(p, i, j, k, c, (int64_t)pre.x + ea, (int64_t)pre.y + eb, vid, seam_in, id_offset, faces);
}

// ids of the vertices lying in the last plane of axis 0 (for the slab above): [2][R1][R2]
__global__ void mc_seam_export_kernel(McParams p, const int32_t *__restrict__ vid, int32_t *__restrict__ seam_out)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
    if (k >= p.R2) return;
    const int i = p.R0 - 1;
    const double d0 = (double)p.vol[node_lin(p, i, j, k)] - p.level;
    int32_t a1 = -1, a2 = -1;
    if (j + 1 < p.R1) {
        const double d1 = (double)p.vol[node_lin(p, i, j + 1, k)] - p.level;
        if ((d0 > 0.0) != (d1 > 0.0)) a1 = vid[3 * node_lin(p, i, j, k) + 1];
    }
    if (k + 1 < p.R2) {
        const double d2 = (double)p.vol[node_lin(p, i, j, k + 1)] - p.level;
        if ((d0 > 0.0) != (d2 > 0.0)) a2 = vid[3 * node_lin(p, i, j, k) + 2];
    }
    seam_out[(int64_t)j * p.R2 + k] = a1;
    seam_out[((int64_t)p.R1 + j) * p.R2 + k] = a2;
}

// ------------------------------------------------------------------------------------------
// Fast path (R2 % 4 == 0): every thread looks at 4 consecutive cells of a k-row with 4 aligned
// float4 loads (+ 4 scalars), rejects the all-inside / all-outside case on 20 sign bits, and only
// surface cells go through classify_vals.  A first pass counts, a second pass writes the compact,
// scan-ordered list of active cells (with their vertex / face offsets); vertices and faces are
// then emitted by one thread per ACTIVE cell, so the volume is read twice and never again.
// ------------------------------------------------------------------------------------------
struct CellRec { uint32_t lin, vbase, fbase, pad; };

struct Quad {
    int i, j, k0;
    int ncell;              // number of valid cells among k0 .. k0+3
    unsigned bits;          // sign bits: 5 per row, rows (0,0) (0,1) (1,0) (1,1)
    float v[4][5];
};

__device__ __forceinline__ bool load_quad(const McParams &p, float level, uint32_t q, uint32_t nquad, Quad &Q)
{
    Q.ncell = 0;
    if (q >= nquad) return false;
    const uint32_t qrow = (uint32_t)p.R2 >> 2;
    Q.k0 = (int)(q % qrow) * 4;
    const uint32_t t = q / qrow;
    Q.j = (int)(t % (uint32_t)p.R1);
    Q.i = (int)(t / (uint32_t)p.R1);
    if (Q.i >= p.R0 - 1 || Q.j >= p.R1 - 1) return false;
    Q.ncell = min(4, p.R2 - 1 - Q.k0);
    unsigned bits = 0;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const float *row = p.vol + node_lin(p, Q.i + (r >> 1), Q.j + (r & 1), Q.k0);
        const float4 a = __ldg(reinterpret_cast<const float4 *>(row));
        Q.v[r][0] = a.x; Q.v[r][1] = a.y; Q.v[r][2] = a.z; Q.v[r][3] = a.w;
        Q.v[r][4] = Q.k0 + 4 < p.R2 ? __ldg(row + 4) : a.w;
#pragma unroll
        for (int c = 0; c < 5; ++c) bits |= (Q.v[r][c] > level ? 1u : 0u) << (5 * r + c);
    }
    Q.bits = bits;
    return bits != 0u && bits != 0xFFFFFu;
}

// cell t of the quad: corner values in Lewiner order; returns false when the cell has no surface
__device__ __forceinline__ bool quad_cell(const Quad &Q, int t, float (&val)[8])
{
    const unsigned b = Q.bits >> t;
    const unsigned cas = (b & 1u) | ((b >> 1) & 1u) << 1 | ((b >> 6) & 1u) << 2 | ((b >> 5) & 1u) << 3 |
                         ((b >> 10) & 1u) << 4 | ((b >> 11) & 1u) << 5 | ((b >> 16) & 1u) << 6 | ((b >> 15) & 1u) << 7;
    if (cas == 0u || cas == 255u) return false;
    val[0] = Q.v[0][t]; val[1] = Q.v[0][t + 1]; val[2] = Q.v[1][t + 1]; val[3] = Q.v[1][t];
    val[4] = Q.v[2][t]; val[5] = Q.v[2][t + 1]; val[6] = Q.v[3][t + 1]; val[7] = Q.v[3][t];
    return true;
}

__device__ __forceinline__ void block_scan3(uint3 a, uint3 &excl, uint3 &tot)
{
    __shared__ uint3 wsum[MC_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint3 inc = a;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned x = __shfl_up_sync(0xffffffffu, inc.x, o), y = __shfl_up_sync(0xffffffffu, inc.y, o), z = __shfl_up_sync(0xffffffffu, inc.z, o);
        if (lane >= o) { inc.x += x; inc.y += y; inc.z += z; }
    }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    uint3 off = make_uint3(0, 0, 0), sum = make_uint3(0, 0, 0);
#pragma unroll
    for (int w = 0; w < MC_THREADS / 32; ++w) {
        if (w == warp) off = sum;
        sum.x += wsum[w].x; sum.y += wsum[w].y; sum.z += wsum[w].z;
    }
    excl = make_uint3(off.x + inc.x - a.x, off.y + inc.y - a.y, off.z + inc.z - a.z);
    tot = sum;
    __syncthreads();
}

template <bool WRITE>
__global__ void __launch_bounds__(MC_THREADS) mc_quad_kernel(McParams p, float level, uint32_t nquad, uint4 *block_tot,
                                                             unsigned long long *n_amb, CellRec *cells)
{
    const uint32_t q = blockIdx.x * MC_THREADS + threadIdx.x;
    Quad Q;
    uint3 mine = make_uint3(0, 0, 0);
    unsigned amb = 0;
    unsigned nv[4] = {0, 0, 0, 0}, nt[4] = {0, 0, 0, 0};
    const bool any = load_quad(p, level, q, nquad, Q);
    if (any) {
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            float val[8];
            if (t < Q.ncell && quad_cell(Q, t, val)) {
                Cell c;
                classify_vals(p, Q.i, Q.j, Q.k0 + t, val, c);
                nv[t] = (unsigned)c.nv; nt[t] = (unsigned)c.nt;
                if (c.nt | c.nv) { mine.x += 1; mine.y += nv[t]; mine.z += nt[t]; }
                amb += c.ambiguous;
            }
        }
    }
    uint3 excl, tot;
    block_scan3(mine, excl, tot);
    if (!WRITE) {
        if (amb) atomicAdd(n_amb, (unsigned long long)amb);
        if (threadIdx.x == 0) block_tot[blockIdx.x] = make_uint4(tot.x, tot.y, tot.z, 0);
    } else if (mine.x) {
        const uint4 pre = block_tot[blockIdx.x];
        uint32_t a = pre.x + excl.x, vb = pre.y + excl.y, fb = pre.z + excl.z;
        const uint32_t lin0 = q * 4;
#pragma unroll
        for (int t = 0; t < 4; ++t)
            if (nv[t] | nt[t]) {
                CellRec r;
                r.lin = lin0 + t; r.vbase = vb; r.fbase = fb; r.pad = 0;
                cells[a++] = r;
                vb += nv[t]; fb += nt[t];
            }
    }
}

__global__ void __launch_bounds__(1024) mc_scan_blocks3_kernel(uint4 *block_tot, int64_t nblocks, unsigned long long *totals)
{
    __shared__ unsigned long long wsum[3][32];
    __shared__ unsigned long long carry[3];
    if (threadIdx.x < 3) carry[threadIdx.x] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t base = 0; base < nblocks; base += 1024) {
        const int64_t idx = base + threadIdx.x;
        const uint4 v = idx < nblocks ? block_tot[idx] : make_uint4(0, 0, 0, 0);
        unsigned long long a[3] = {v.x, v.y, v.z};
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const unsigned long long x = __shfl_up_sync(0xffffffffu, a[c], o);
                if (lane >= o) a[c] += x;
            }
        if (lane == 31)
            for (int c = 0; c < 3; ++c) wsum[c][warp] = a[c];
        __syncthreads();
        unsigned long long off[3] = {carry[0], carry[1], carry[2]};
        for (int w = 0; w < warp; ++w)
            for (int c = 0; c < 3; ++c) off[c] += wsum[c][w];
        if (idx < nblocks)
            block_tot[idx] = make_uint4((unsigned)(off[0] + a[0] - v.x), (unsigned)(off[1] + a[1] - v.y), (unsigned)(off[2] + a[2] - v.z), 0);
        __syncthreads();
        if (threadIdx.x == 1023)
            for (int c = 0; c < 3; ++c) carry[c] = off[c] + a[c];
        __syncthreads();
    }
    if (threadIdx.x == 0) { totals[0] = carry[1]; totals[1] = carry[2]; totals[2] = carry[0]; }
}

__global__ void __launch_bounds__(128) mc_list_verts_kernel(McParams p, const CellRec *__restrict__ cells, uint32_t nact, McOut o)
{
    const uint32_t a = blockIdx.x * 128 + threadIdx.x;
    if (a >= nact) return;
    const CellRec r = cells[a];
    int i, j, k;
    cell_coords(p, (int64_t)r.lin, i, j, k);
    Cell c;
    classify(p, i, j, k, c);
    if (c.nv) emit_cell_verts(p, o, i, j, k, c, (int64_t)r.vbase);
}

__global__ void __launch_bounds__(128) mc_list_faces_kernel(McParams p, const CellRec *__restrict__ cells, uint32_t nact,
                                                            const int32_t *__restrict__ vid, const int32_t *__restrict__ seam_in,
                                                            int64_t id_offset, int32_t *__restrict__ faces)
{
    const uint32_t a = blockIdx.x * 128 + threadIdx.x;
    if (a >= nact) return;
    const CellRec r = cells[a];
    int i, j, k;
    cell_coords(p, (int64_t)r.lin, i, j, k);
    Cell c;
    classify(p, i, j, k, c);
    if (c.nt) emit_cell_faces(p, i, j, k, c, (int64_t)r.vbase, (int64_t)r.fbase, vid, seam_in, id_offset, faces);
}

struct TableBlob {
    uint8_t amb_mask[256];
    uint16_t var_base[256];
    uint8_t ntri[MC_NUM_ENTRIES];
    uint16_t tri_off[MC_NUM_ENTRIES];
    uint8_t tri_edges[MC_NUM_TRI_IDX];
    uint8_t edge_rank[MC_NUM_ENTRIES * 13];
};

McParams make_params(surs_ctx *ctx)
{
    McParams p;
    p.vol = ctx->mc_vol;
    p.R0 = ctx->mc_res[0]; p.R1 = ctx->mc_res[1]; p.R2 = ctx->mc_res[2];
    p.level = (double)ctx->mc_level;
    p.lower_foreign = (ctx->mc_flags & SURS_MC_LOWER_FOREIGN) ? 1 : 0;
    const TableBlob *b = (const TableBlob *)ctx->mc_tables;
    p.tb.amb_mask = b->amb_mask; p.tb.var_base = b->var_base; p.tb.ntri = b->ntri;
    p.tb.tri_off = b->tri_off; p.tb.tri_edges = b->tri_edges; p.tb.edge_rank = b->edge_rank;
    return p;
}

}  // namespace

int surs_mc_init_tables(surs_ctx *ctx)
{
    TableBlob *h = new TableBlob();
    memcpy(h->amb_mask, mc_amb_mask, sizeof(h->amb_mask));
    memcpy(h->var_base, mc_var_base, sizeof(h->var_base));
    memcpy(h->ntri, mc_ntri, sizeof(h->ntri));
    memcpy(h->tri_off, mc_tri_off, sizeof(h->tri_off));
    memcpy(h->tri_edges, mc_tri_edges, sizeof(h->tri_edges));
    memcpy(h->edge_rank, mc_edge_rank, sizeof(h->edge_rank));
    cudaError_t e = cudaMalloc(&ctx->mc_tables, sizeof(TableBlob));
    if (e == cudaSuccess) e = cudaMemcpy(ctx->mc_tables, h, sizeof(TableBlob), cudaMemcpyHostToDevice);
    delete h;
    if (e == cudaSuccess) e = cudaMemcpyToSymbol(c_corner_off, mc_corner_off, 24);
    if (e == cudaSuccess) e = cudaMemcpyToSymbol(c_edge_corner, mc_edge_corner, 24);
    if (e == cudaSuccess) e = cudaMemcpyToSymbol(c_edge_axis, mc_edge_axis, 12);
    if (e == cudaSuccess) e = cudaMemcpyToSymbol(c_edge_base, mc_edge_base, 36);
    if (e == cudaSuccess) e = cudaMemcpyToSymbol(c_face_corner, mc_face_corner, 24);
    if (e != cudaSuccess) SURS_FAIL(ctx, "marching-cubes table upload failed: %s", cudaGetErrorString(e));
    return 0;
}

extern "C" int surs_mc_count(surs_ctx *ctx, const float *vol, const int res[3], float level, int flags,
                             int64_t *n_verts, int64_t *n_faces, int64_t *n_ambiguous, void *stream)
{
    if (!ctx) return 1;
    cudaStream_t st = (cudaStream_t)stream;
    SURS_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!vol || res[0] < 2 || res[1] < 2 || res[2] < 2) SURS_FAIL(ctx, "surs_mc_count: volume needs at least 2 nodes per axis");
    const int64_t nnode = (int64_t)res[0] * res[1] * res[2];
    if (3 * nnode >= ((int64_t)1 << 31)) SURS_FAIL(ctx, "surs_mc_count: volume too large for 32-bit edge ids (3*res^3 < 2^31)");
    ctx->mc_vol = vol;
    memcpy(ctx->mc_res, res, sizeof(int) * 3);
    ctx->mc_level = level;
    ctx->mc_flags = flags;
    SURS_CUDA(ctx, cudaMemsetAsync(ctx->counter, 0, 64, st));
    McParams p = make_params(ctx);
    static const bool no_fast = getenv("SURS_MC_SLOW") != nullptr;
    ctx->mc_fast = (res[2] % 4 == 0) && ((uintptr_t)vol % 16 == 0) && !no_fast;
    if (ctx->mc_fast) {
        const uint32_t nquad = (uint32_t)(nnode / 4);
        const int64_t nb = ((int64_t)nquad + MC_THREADS - 1) / MC_THREADS;
        if (surs_ensure(ctx, (void **)&ctx->mc_block_tot, &ctx->mc_block_cap, sizeof(uint4) * (size_t)nb)) return 1;
        uint4 *bt = reinterpret_cast<uint4 *>(ctx->mc_block_tot);
        mc_quad_kernel<false><<<(unsigned)nb, MC_THREADS, 0, st>>>(p, level, nquad, bt, ctx->counter + 1, nullptr);
        SURS_LAUNCH_CHECK(ctx, "mc_quad_kernel<count>");
        mc_scan_blocks3_kernel<<<1, 1024, 0, st>>>(bt, nb, ctx->counter + 2);
        SURS_LAUNCH_CHECK(ctx, "mc_scan_blocks3_kernel");
        unsigned long long host[5];
        SURS_CUDA(ctx, cudaMemcpyAsync(host, ctx->counter, sizeof(host), cudaMemcpyDeviceToHost, st));
        SURS_CUDA(ctx, cudaStreamSynchronize(st));
        ctx->mc_nv = (int64_t)host[2];
        ctx->mc_nf = (int64_t)host[3];
        ctx->mc_nact = (int64_t)host[4];
        if (ctx->mc_nv >= ((int64_t)1 << 31) || ctx->mc_nf >= ((int64_t)1 << 31)) SURS_FAIL(ctx, "surs_mc_count: mesh too large for int32 indices");
        if (ctx->mc_nact > 0) {
            if (surs_ensure(ctx, (void **)&ctx->mc_cells, &ctx->mc_cells_cap, sizeof(CellRec) * (size_t)ctx->mc_nact)) return 1;
            mc_quad_kernel<true><<<(unsigned)nb, MC_THREADS, 0, st>>>(p, level, nquad, bt, nullptr, reinterpret_cast<CellRec *>(ctx->mc_cells));
            SURS_LAUNCH_CHECK(ctx, "mc_quad_kernel<compact>");
        }
        if (n_verts) *n_verts = ctx->mc_nv;
        if (n_faces) *n_faces = ctx->mc_nf;
        if (n_ambiguous) *n_ambiguous = (int64_t)host[1];
        return 0;
    }
    const int64_t nblocks = (nnode + MC_THREADS - 1) / MC_THREADS;
    if (surs_ensure(ctx, (void **)&ctx->mc_block_tot, &ctx->mc_block_cap, sizeof(uint2) * (size_t)nblocks)) return 1;
    mc_count_kernel<<<(unsigned)nblocks, MC_THREADS, 0, st>>>(p, nnode, reinterpret_cast<uint2 *>(ctx->mc_block_tot), ctx->counter + 1);
    SURS_LAUNCH_CHECK(ctx, "mc_count_kernel");
    mc_scan_blocks_kernel<<<1, 1024, 0, st>>>(reinterpret_cast<uint2 *>(ctx->mc_block_tot), nblocks, ctx->counter + 2);
    SURS_LAUNCH_CHECK(ctx, "mc_scan_blocks_kernel");
    unsigned long long host[4];
    SURS_CUDA(ctx, cudaMemcpyAsync(host, ctx->counter, sizeof(host), cudaMemcpyDeviceToHost, st));
    SURS_CUDA(ctx, cudaStreamSynchronize(st));
    ctx->mc_nv = (int64_t)host[2];
    ctx->mc_nf = (int64_t)host[3];
    if (ctx->mc_nv >= ((int64_t)1 << 31) || ctx->mc_nf >= ((int64_t)1 << 31)) SURS_FAIL(ctx, "surs_mc_count: mesh too large for int32 indices");
    if (n_verts) *n_verts = ctx->mc_nv;
    if (n_faces) *n_faces = ctx->mc_nf;
    if (n_ambiguous) *n_ambiguous = (int64_t)host[1];
    return 0;
}

extern "C" int surs_mc_emit_verts(surs_ctx *ctx, const double *mat, float *verts, double *verts_world,
                                  float *normals, float *values, int64_t vert_id_offset, int plane_offset,
                                  int32_t *seam_out, void *stream)
{
    if (!ctx) return 1;
    cudaStream_t st = (cudaStream_t)stream;
    SURS_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!ctx->mc_vol) SURS_FAIL(ctx, "surs_mc_emit_verts: call surs_mc_count first");
    if (ctx->mc_nv > 0 && !verts) SURS_FAIL(ctx, "surs_mc_emit_verts: null output");
    const int64_t nnode = (int64_t)ctx->mc_res[0] * ctx->mc_res[1] * ctx->mc_res[2];
    const int64_t nblocks = (nnode + MC_THREADS - 1) / MC_THREADS;
    if (surs_ensure(ctx, (void **)&ctx->mc_vid, &ctx->mc_vid_cap, sizeof(int32_t) * 3 * (size_t)nnode)) return 1;
    McParams p = make_params(ctx);
    McOut o;
    memset(&o, 0, sizeof(o));
    o.verts = verts; o.verts_world = verts_world; o.normals = normals; o.values = values;
    o.has_mat = mat != nullptr;
    if (mat) memcpy(o.mat, mat, sizeof(double) * 12);
    o.vid = ctx->mc_vid;
    o.id_offset = vert_id_offset;
    o.plane_offset = plane_offset;
    ctx->mc_id_offset = vert_id_offset;
    if (ctx->mc_nv > 0 && ctx->mc_fast) {
        mc_list_verts_kernel<<<(unsigned)((ctx->mc_nact + 127) / 128), 128, 0, st>>>(p, reinterpret_cast<const CellRec *>(ctx->mc_cells),
                                                                                     (uint32_t)ctx->mc_nact, o);
        SURS_LAUNCH_CHECK(ctx, "mc_list_verts_kernel");
    } else if (ctx->mc_nv > 0) {
        mc_emit_verts_kernel<<<(unsigned)nblocks, MC_THREADS, 0, st>>>(p, nnode, reinterpret_cast<const uint2 *>(ctx->mc_block_tot), o);
        SURS_LAUNCH_CHECK(ctx, "mc_emit_verts_kernel");
    }
    if (seam_out) {
        dim3 grid((p.R2 + 127) / 128, p.R1);
        mc_seam_export_kernel<<<grid, 128, 0, st>>>(p, ctx->mc_vid, seam_out);
        SURS_LAUNCH_CHECK(ctx, "mc_seam_export_kernel");
    }
    return 0;
}

extern "C" int surs_mc_emit_faces(surs_ctx *ctx, int32_t *faces, const int32_t *seam_in, void *stream)
{
    if (!ctx) return 1;
    cudaStream_t st = (cudaStream_t)stream;
    SURS_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!ctx->mc_vol || !ctx->mc_vid) SURS_FAIL(ctx, "surs_mc_emit_faces: call surs_mc_count and surs_mc_emit_verts first");
    if ((ctx->mc_flags & SURS_MC_LOWER_FOREIGN) && !seam_in) SURS_FAIL(ctx, "surs_mc_emit_faces: seam_in required with SURS_MC_LOWER_FOREIGN");
    if (ctx->mc_nf > 0 && !faces) SURS_FAIL(ctx, "surs_mc_emit_faces: null output");
    const int64_t nnode = (int64_t)ctx->mc_res[0] * ctx->mc_res[1] * ctx->mc_res[2];
    const int64_t nblocks = (nnode + MC_THREADS - 1) / MC_THREADS;
    McParams p = make_params(ctx);
    if (ctx->mc_nf > 0 && ctx->mc_fast) {
        mc_list_faces_kernel<<<(unsigned)((ctx->mc_nact + 127) / 128), 128, 0, st>>>(p, reinterpret_cast<const CellRec *>(ctx->mc_cells),
                                                                                     (uint32_t)ctx->mc_nact, ctx->mc_vid, seam_in,
                                                                                     ctx->mc_id_offset, faces);
        SURS_LAUNCH_CHECK(ctx, "mc_list_faces_kernel");
    } else if (ctx->mc_nf > 0) {
        mc_emit_faces_kernel<<<(unsigned)nblocks, MC_THREADS, 0, st>>>(p, nnode, reinterpret_cast<const uint2 *>(ctx->mc_block_tot), ctx->mc_vid,
                                                                      seam_in, ctx->mc_id_offset, faces);
        SURS_LAUNCH_CHECK(ctx, "mc_emit_faces_kernel");
    }
    return 0;
}

extern "C" int surs_mc_emit(surs_ctx *ctx, const double *mat, float *verts, double *verts_world,
                            int32_t *faces, float *normals, float *values, void *stream)
{
    if (surs_mc_emit_verts(ctx, mat, verts, verts_world, normals, values, 0, 0, nullptr, stream)) return 1;
    return surs_mc_emit_faces(ctx, faces, nullptr, stream);
}
