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
// -> emit faces.  Case tables: csrc/mc_tables.h (derivation in csrc/gen_mc_tables.py).
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
    const uint8_t *ntest;        // [E] interior (tunnel) tests of the entry
    const uint16_t *test_off;    // [E]
    const uint8_t *test_corners; // [T][8]
    const uint8_t *test_sign;    // [T]
    const uint16_t *test_target; // [T]
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
    int interior;       // 0: no interior ambiguity, 1: interior test(s) evaluated, all failed, 2: tunnel variant taken
    double d[8];
};

// Interior (tunnel) test, csrc/gen_mc_tables.py: the regions on the lines A and C (cube edges parallel to one axis,
// cyclic order A B C D) are joined through the cell iff the section at the stationary point of q(t) = At Ct - Bt Dt
// has At, Ct of sign s and q > 0.  Same operation order as oracle/mc_oracle.c and the generator (no contraction).
__device__ __forceinline__ bool interior_test(const double (&d)[8], const uint8_t *cn, int s)
{
    const double a0 = d[cn[0]], a1 = d[cn[1]], b0 = d[cn[2]], b1 = d[cn[3]], c0 = d[cn[4]], c1 = d[cn[5]], d0 = d[cn[6]], d1 = d[cn[7]];
    const double dA = __dsub_rn(a1, a0), dB = __dsub_rn(b1, b0), dC = __dsub_rn(c1, c0), dD = __dsub_rn(d1, d0);
    const double qa = __dsub_rn(__dmul_rn(dA, dC), __dmul_rn(dB, dD));
    const double qb = __dsub_rn(__dadd_rn(__dmul_rn(a0, dC), __dmul_rn(c0, dA)), __dadd_rn(__dmul_rn(b0, dD), __dmul_rn(d0, dB)));
    if (!(qa < 0.0)) return false;
    const double t = __ddiv_rn(-qb, __dmul_rn(2.0, qa));
    if (!(t > 0.0 && t < 1.0)) return false;
    const double At = __dadd_rn(a0, __dmul_rn(dA, t)), Bt = __dadd_rn(b0, __dmul_rn(dB, t));
    const double Ct = __dadd_rn(c0, __dmul_rn(dC, t)), Dt = __dadd_rn(d0, __dmul_rn(dD, t));
    if (s) {
        if (!(At > 0.0 && Ct > 0.0)) return false;
    } else if (At > 0.0 || Ct > 0.0) {
        return false;
    }
    return __dsub_rn(__dmul_rn(At, Ct), __dmul_rn(Bt, Dt)) > 0.0;
}

// 32-bit node index: surs_mc_count refuses volumes with 3 * nodes >= 2^31 (the edge -> vertex-id map is indexed by 3 * node + axis)
__device__ __forceinline__ int32_t node_lin(const McParams &p, int i, int j, int k) { return (i * p.R1 + j) * p.R2 + k; }

__device__ __forceinline__ void classify_vals(const McParams &p, int i, int j, int k, const float (&val)[8], Cell &c);

__device__ __forceinline__ void classify(const McParams &p, int i, int j, int k, Cell &c)
{
    c.entry = -1; c.nv = 0; c.nt = 0; c.owned = 0; c.ambiguous = 0; c.interior = 0;
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
    c.entry = -1; c.nv = 0; c.nt = 0; c.owned = 0; c.ambiguous = 0; c.interior = 0;
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
    const int ntest = p.tb.ntest[c.entry];
    if (ntest) {                                   // rare: two same-sign regions that a tunnel through the cell may join
        c.interior = 1;
        const int t0 = p.tb.test_off[c.entry];
        for (int q = 0; q < ntest; ++q)
            if (interior_test(c.d, p.tb.test_corners + 8 * (t0 + q), p.tb.test_sign[t0 + q])) {
                c.entry = p.tb.test_target[t0 + q];
                c.interior = 2;
                break;
            }
    }
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
    c.nv = c.nt = 0; c.ambiguous = 0; c.interior = 0;
    if (lin < nnode) {
        int i, j, k;
        cell_coords(p, lin, i, j, k);
        classify(p, i, j, k, c);
    }
    unsigned ea, eb, ta, tb;
    block_scan2((unsigned)c.nv, (unsigned)c.nt, ea, eb, ta, tb);
    if (c.interior) atomicAdd(n_amb + 8, 1ull);                 // counter[9]: cells with an interior ambiguity
    if (c.interior == 2) atomicAdd(n_amb + 9, 1ull);            // counter[10]: ... that took the tunnel variant
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

// volume gradient at a node (central differences, one-sided at the border).  Normals are an API-shape obligation only
// (the reference's caller discards them, lib/train_util.py:72) and are compared with a 1e-4 tolerance: fp32 arithmetic
__device__ __forceinline__ float node_grad(const McParams &p, int i, int j, int k, int axis)
{
    const int n = axis == 0 ? p.R0 : (axis == 1 ? p.R1 : p.R2);
    const int q = axis == 0 ? i : (axis == 1 ? j : k);
    const int lo = q > 0 ? q - 1 : q, hi = q < n - 1 ? q + 1 : q;
    const int stride = axis == 0 ? p.R1 * p.R2 : (axis == 1 ? p.R2 : 1);
    const int32_t c = node_lin(p, i, j, k);
    const float dv = __ldg(p.vol + c + (hi - q) * stride) - __ldg(p.vol + c - (q - lo) * stride);
    return (hi - lo) == 2 ? dv * 0.5f : dv;
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

__device__ __forceinline__ void write_vertex(const McOut &o, int64_t v, const double pos[3], const float g[3], double value)
{
    const float fx = (float)pos[0], fy = (float)pos[1], fz = (float)pos[2];
    if (o.verts) { o.verts[3 * v] = fx; o.verts[3 * v + 1] = fy; o.verts[3 * v + 2] = fz; }
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
        const float n2 = g[0] * g[0] + g[1] * g[1] + g[2] * g[2];
        const float inv = n2 > 0.0f ? rsqrtf(n2) : 0.0f;
#pragma unroll
        for (int a = 0; a < 3; ++a) o.normals[3 * v + a] = g[a] * inv;
    }
    if (o.values) o.values[v] = (float)value;
}

// rank of vertex slot `slot` (0..11 cube edges, 12 centre) among the slots this cell creates, in first-use order
__device__ __forceinline__ int owned_rank(const uint8_t *rank, unsigned owned, int slot)
{
    int r = 0;
    for (int e2 = 0; e2 < 13; ++e2)
        if (((owned >> e2) & 1u) && rank[e2] < rank[slot]) ++r;
    return r;
}

// the vertex on cube edge e of cell (i,j,k): da / db = value - level at the edge's lower / upper end
__device__ __forceinline__ void emit_edge_vertex(const McParams &p, const McOut &o, int i, int j, int k, int e, double dlo, double dhi, int64_t v)
{
    const int axis = c_edge_axis[e];
    const int bi = i + c_edge_base[3 * e], bj = j + c_edge_base[3 * e + 1], bk = k + c_edge_base[3 * e + 2];
    const int ei = bi + (axis == 0), ej = bj + (axis == 1), ek = bk + (axis == 2);
    double pos[3] = {(double)(bi + o.plane_offset), (double)bj, (double)bk};
    const double base = pos[axis];
    const double x = edge_point(base, dlo, dhi);
    pos[axis] = x;
    float g[3] = {0.0f, 0.0f, 0.0f};
    if (o.normals) {
        const float tt = (float)(x - base);
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const float ga = node_grad(p, bi, bj, bk, a), gb = node_grad(p, ei, ej, ek, a);
            g[a] = ga + tt * (gb - ga);
        }
    }
    const double va = (double)__ldg(p.vol + node_lin(p, bi, bj, bk)), vb = (double)__ldg(p.vol + node_lin(p, ei, ej, ek));
    write_vertex(o, v, pos, g, va > vb ? va : vb);
    o.vid[3 * node_lin(p, bi, bj, bk) + axis] = (int32_t)(v + o.id_offset);
}

// Lewiner's centre vertex: mean of the cell's edge vertices; d[q] = value - level at corner q
__device__ __forceinline__ void emit_centre_vertex(const McParams &p, const McOut &o, int i, int j, int k, const double (&d)[8], int64_t v)
{
    double s[3] = {0.0, 0.0, 0.0};
    int n = 0;
    for (int e = 0; e < 12; ++e) {
        const int ca = c_edge_corner[2 * e], cb = c_edge_corner[2 * e + 1];
        if ((d[ca] > 0.0) == (d[cb] > 0.0)) continue;
        const int axis = c_edge_axis[e];
        const int lo = c_corner_off[3 * ca + axis] == 0 ? ca : cb, hi = lo == ca ? cb : ca;
        double q[3] = {(double)(i + o.plane_offset + c_edge_base[3 * e]), (double)(j + c_edge_base[3 * e + 1]), (double)(k + c_edge_base[3 * e + 2])};
        q[axis] = edge_point(q[axis], d[lo], d[hi]);
        s[0] = __dadd_rn(s[0], q[0]); s[1] = __dadd_rn(s[1], q[1]); s[2] = __dadd_rn(s[2], q[2]);
        ++n;
    }
    double pos[3] = {__ddiv_rn(s[0], (double)n), __ddiv_rn(s[1], (double)n), __ddiv_rn(s[2], (double)n)};
    float g[3] = {0.0f, 0.0f, 0.0f};
    double vmax = -INFINITY;
    for (int q = 0; q < 8; ++q) {
        const int ci = i + c_corner_off[3 * q], cj = j + c_corner_off[3 * q + 1], ck = k + c_corner_off[3 * q + 2];
        if (o.normals)
            for (int a = 0; a < 3; ++a) g[a] += node_grad(p, ci, cj, ck, a);
        vmax = fmax(vmax, (double)__ldg(p.vol + node_lin(p, ci, cj, ck)));
    }
    write_vertex(o, v, pos, g, vmax);
}

// all vertices created by cell (i,j,k); vbase = index of its first vertex in this volume's list
__device__ __forceinline__ void emit_cell_verts(const McParams &p, const McOut &o, int i, int j, int k, const Cell &c, int64_t vbase)
{
    const uint8_t *rank = p.tb.edge_rank + 13 * c.entry;
#pragma unroll 1
    for (int e = 0; e < 12; ++e) {
        if (!((c.owned >> e) & 1u)) continue;
        const int axis = c_edge_axis[e];
        const int ca = c_edge_corner[2 * e], cb = c_edge_corner[2 * e + 1];
        const int lo = c_corner_off[3 * ca + axis] == 0 ? ca : cb, hi = lo == ca ? cb : ca;
        emit_edge_vertex(p, o, i, j, k, e, c.d[lo], c.d[hi], vbase + owned_rank(rank, c.owned, e));
    }
    if ((c.owned >> 12) & 1u) emit_centre_vertex(p, o, i, j, k, c.d, vbase + owned_rank(rank, c.owned, 12));
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
                                                int64_t id_offset, int32_t *__restrict__ faces, unsigned long long *seam_bad)
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
            if (p.lower_foreign && bi == 0 && axis != 0) {
                // the lower slab owns this vertex; -1 means it saw no sign change on the edge, i.e. the two slabs disagree
                // about an inside / outside bit of the shared plane: counted, reported by surs_mc_seam_violations
                tri[s] = seam_in[((int64_t)(axis - 1) * p.R1 + bj) * p.R2 + bk];
                if (tri[s] < 0) atomicAdd(seam_bad, 1ull);
            } else
                tri[s] = vid[3 * node_lin(p, bi, bj, bk) + axis];
        }
        faces[3 * (fbase + t)] = tri[0];
        faces[3 * (fbase + t) + 1] = tri[1];
        faces[3 * (fbase + t) + 2] = tri[2];
    }
}

__global__ void __launch_bounds__(MC_THREADS) mc_emit_faces_kernel(McParams p, int64_t nnode, const uint2 *block_prefix,
                                                                   const int32_t *__restrict__ vid, const int32_t *__restrict__ seam_in,
                                                                   int64_t id_offset, int32_t *__restrict__ faces, unsigned long long *seam_bad)
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
    emit_cell_faces(p, i, j, k, c, (int64_t)pre.x + ea, (int64_t)pre.y + eb, vid, seam_in, id_offset, faces, seam_bad);
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
// Fast path (R2 % 4 == 0).  The volume is streamed from HBM exactly once:
//   mc_sign_kernel      one warp per node row, 128-byte coalesced loads, __ballot -> 1 bit per node
//                       (17 MB for 512^3: L2 resident for everything that follows)
//   mc_bits_kernel      one thread per 32 cells of a k-row: the cell is active iff its 8 corner bits
//                       differ -- six word loads and a dozen logic ops per 32 cells; pass 1 counts,
//                       pass 2 writes the scan-ordered list of active cells
//   mc_cell_kernel      one thread per ACTIVE cell (~2 % of the cells): classification with the
//                       ambiguity deciders -> vertex / triangle counts, block-relative offsets
//   mc_list_verts / mc_list_faces   one thread per active cell
// ------------------------------------------------------------------------------------------
// vbase / fbase relative to the cell's mc_cell_kernel block; cls = table entry | owned vertex slots << 16: the list
// kernels do not repeat the classification (deciders, interior test), and the face kernel does not touch the volume
struct CellRec { uint32_t lin, vbase, fbase, cls; };

constexpr int MC_SIGN_WARPS = 8;

// float <-> unsigned key with the same ordering (for atomicMin / atomicMax on the value range)
__device__ __forceinline__ unsigned f2key(float f)
{
    const unsigned b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__host__ __device__ __forceinline__ float key2f(unsigned k)
{
    const unsigned b = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    float f;
    memcpy(&f, &b, sizeof(f));
    return f;
}

// One pass over the volume does everything that needs every value: the inside / outside bit per node, the value range
// (what skimage checks the level against: replaces a separate aminmax pass) and, for a float64 source (octree volumes,
// lib/sdf.py keeps them in float64), the float32 copy that skimage makes of its input (replaces a separate cast pass).
// range[0] / range[1]: ordered keys of the minimum / maximum, updated with one guarded atomic per warp.
// sixteen streaming loads, 32 elements apart, issued back to back (one asm block per eight: the compiler otherwise
// interleaves the range arithmetic with the loads and halves the memory-level parallelism)
__device__ __forceinline__ void load16_cs(const float *p, float (&r)[16])
{
#pragma unroll
    for (int h = 0; h < 2; ++h)
        asm volatile("ld.global.cs.f32 %0, [%8];\n\tld.global.cs.f32 %1, [%8+128];\n\tld.global.cs.f32 %2, [%8+256];\n\t"
                     "ld.global.cs.f32 %3, [%8+384];\n\tld.global.cs.f32 %4, [%8+512];\n\tld.global.cs.f32 %5, [%8+640];\n\t"
                     "ld.global.cs.f32 %6, [%8+768];\n\tld.global.cs.f32 %7, [%8+896];"
                     : "=f"(r[8 * h]), "=f"(r[8 * h + 1]), "=f"(r[8 * h + 2]), "=f"(r[8 * h + 3]), "=f"(r[8 * h + 4]), "=f"(r[8 * h + 5]),
                       "=f"(r[8 * h + 6]), "=f"(r[8 * h + 7])
                     : "l"(p + 256 * h));
}
__device__ __forceinline__ void load16_cs(const double *p, double (&r)[16])
{
#pragma unroll
    for (int h = 0; h < 2; ++h)
        asm volatile("ld.global.cs.f64 %0, [%8];\n\tld.global.cs.f64 %1, [%8+256];\n\tld.global.cs.f64 %2, [%8+512];\n\t"
                     "ld.global.cs.f64 %3, [%8+768];\n\tld.global.cs.f64 %4, [%8+1024];\n\tld.global.cs.f64 %5, [%8+1280];\n\t"
                     "ld.global.cs.f64 %6, [%8+1536];\n\tld.global.cs.f64 %7, [%8+1792];"
                     : "=d"(r[8 * h]), "=d"(r[8 * h + 1]), "=d"(r[8 * h + 2]), "=d"(r[8 * h + 3]), "=d"(r[8 * h + 4]), "=d"(r[8 * h + 5]),
                       "=d"(r[8 * h + 6]), "=d"(r[8 * h + 7])
                     : "l"(p + 256 * h));
}

template <typename T, bool COPY32>
__global__ void __launch_bounds__(MC_SIGN_WARPS * 32) mc_sign_kernel(const T *__restrict__ vol, float *__restrict__ vol32, float level, int64_t nrows,
                                                                      int R2, int W, uint32_t *__restrict__ bits, unsigned *__restrict__ range)
{
    // persistent warps: a warp walks rows with the grid's warp count as stride, so the value range costs one
    // reduction per warp and one guarded atomic per block for the whole volume, not per row
    __shared__ unsigned s_lo[MC_SIGN_WARPS], s_hi[MC_SIGN_WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float lo = INFINITY, hi = -INFINITY;
    for (int64_t row = (int64_t)blockIdx.x * MC_SIGN_WARPS + warp; row < nrows; row += (int64_t)gridDim.x * MC_SIGN_WARPS) {
        const T *src = vol + row * R2;
        uint32_t *dst = bits + row * W;
        for (int w0 = 0; w0 < W; w0 += 16) {
            T raw[16];
            if ((w0 + 16) * 32 <= R2) {
                load16_cs(src + w0 * 32 + lane, raw);                    // streamed: nothing re-reads the source from L1
            } else {
#pragma unroll
                for (int u = 0; u < 16; ++u) {
                    const int k = (w0 + u) * 32 + lane;
                    raw[u] = (w0 + u < W && k < R2) ? __ldcs(src + k) : (T)level;
                }
            }
            float v[16];
#pragma unroll
            for (int u = 0; u < 16; ++u) {
                const int k = (w0 + u) * 32 + lane;
                const bool ok = w0 + u < W && k < R2;
                v[u] = (float)raw[u];
                if (COPY32 && ok) vol32[row * R2 + k] = v[u];
                lo = fminf(lo, ok ? v[u] : lo);                          // (selects, no branches: padding lanes do not count)
                hi = fmaxf(hi, ok ? v[u] : hi);
            }
            uint32_t mine = 0;
#pragma unroll
            for (int u = 0; u < 16; ++u) {
                const uint32_t word = __ballot_sync(0xffffffffu, v[u] > level);
                if (lane == u) mine = word;
            }
            if (lane < 16 && w0 + lane < W) dst[w0 + lane] = mine;
        }
    }
    const unsigned klo = __reduce_min_sync(0xffffffffu, f2key(lo)), khi = __reduce_max_sync(0xffffffffu, f2key(hi));
    if (lane == 0) { s_lo[warp] = klo; s_hi[warp] = khi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned a = s_lo[0], b = s_hi[0];
#pragma unroll
        for (int w = 1; w < MC_SIGN_WARPS; ++w) { a = min(a, s_lo[w]); b = max(b, s_hi[w]); }
        if (a < __ldcg(range)) atomicMin(range, a);
        if (b > __ldcg(range + 1)) atomicMax(range + 1, b);
    }
}

// general path (any shape): value range (+ float32 copy) as a grid-stride pass
template <typename T, bool COPY32>
__global__ void mc_range_kernel(const T *__restrict__ vol, float *__restrict__ vol32, int64_t n, unsigned *__restrict__ range)
{
    float lo = INFINITY, hi = -INFINITY;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float v = (float)vol[i];
        lo = fminf(lo, v);
        hi = fmaxf(hi, v);
        if (COPY32) vol32[i] = v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if ((threadIdx.x & 31) == 0) {
        const unsigned klo = f2key(lo), khi = f2key(hi);
        if (klo < __ldcg(range)) atomicMin(range, klo);
        if (khi > __ldcg(range + 1)) atomicMax(range + 1, khi);
    }
}

__device__ __forceinline__ void block_scan1(unsigned a, unsigned &excl, unsigned &tot)
{
    __shared__ unsigned wsum[MC_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned inc = a;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned x = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += x;
    }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    unsigned off = 0, sum = 0;
#pragma unroll
    for (int w = 0; w < MC_THREADS / 32; ++w) {
        if (w == warp) off = sum;
        sum += wsum[w];
    }
    excl = off + inc - a;
    tot = sum;
    __syncthreads();
}

// WRITE pass: `totals` (device) holds the number of active cells; nothing is written when it exceeds `cap` (the host
// then repeats the pass with a larger list -- the first launch runs on an estimate so that no synchronisation is needed
// between counting and compaction)
template <bool WRITE>
__global__ void __launch_bounds__(MC_THREADS) mc_bits_kernel(McParams p, const uint32_t *__restrict__ bits, int W, uint32_t nthreads,
                                                             uint4 *block_tot, CellRec *cells, const unsigned long long *totals, uint32_t cap)
{
    if (WRITE && *totals > (unsigned long long)cap) return;
    const uint32_t t = blockIdx.x * MC_THREADS + threadIdx.x;
    uint32_t act = 0, row = 0;
    int w = 0;
    if (t < nthreads) {
        w = (int)(t % (uint32_t)W);
        row = t / (uint32_t)W;
        const int j = (int)(row % (uint32_t)p.R1), i = (int)(row / (uint32_t)p.R1);
        const int nvalid = p.R2 - 1 - 32 * w;                              // cells k = 32 w + b with k + 1 < R2
        if (i < p.R0 - 1 && j < p.R1 - 1 && nvalid > 0) {
            const uint32_t *r00 = bits + (size_t)row * W + w, *r01 = r00 + W, *r10 = r00 + (size_t)p.R1 * W, *r11 = r10 + W;
            const bool more = w + 1 < W;
            const uint32_t a = __ldg(r00), b = __ldg(r01), c = __ldg(r10), d = __ldg(r11);
            const uint32_t an = more ? __ldg(r00 + 1) : 0u, bn = more ? __ldg(r01 + 1) : 0u, cn = more ? __ldg(r10 + 1) : 0u,
                           dn = more ? __ldg(r11 + 1) : 0u;
            const uint32_t a1 = (a >> 1) | (an << 31), b1 = (b >> 1) | (bn << 31), c1 = (c >> 1) | (cn << 31), d1 = (d >> 1) | (dn << 31);
            const uint32_t any = a | b | c | d | a1 | b1 | c1 | d1, all = a & b & c & d & a1 & b1 & c1 & d1;
            act = any & ~all;
            if (nvalid < 32) act &= (1u << nvalid) - 1u;
        }
    }
    unsigned excl, tot;
    block_scan1((unsigned)__popc(act), excl, tot);
    if (!WRITE) {
        if (threadIdx.x == 0) block_tot[blockIdx.x] = make_uint4(tot, 0, 0, 0);
    } else {
        uint32_t dst = block_tot[blockIdx.x].x + excl;
        const uint32_t lin0 = row * (uint32_t)p.R2 + 32u * (uint32_t)w;
        while (act) {
            const int bpos = __ffs(act) - 1;
            act &= act - 1;
            cells[dst++].lin = lin0 + (uint32_t)bpos;
        }
    }
}

__global__ void __launch_bounds__(MC_THREADS) mc_cell_kernel(McParams p, CellRec *cells, const unsigned long long *nact_dev, uint32_t cap, uint4 *block_tot,
                                                             unsigned long long *n_amb)
{
    const unsigned long long nact_ll = *nact_dev;
    if (nact_ll > (unsigned long long)cap) return;
    const uint32_t nact = (uint32_t)nact_ll;
    if (blockIdx.x * MC_THREADS >= nact) {                      // launched for the capacity: blocks beyond the list only zero their total
        if (threadIdx.x == 0) block_tot[blockIdx.x] = make_uint4(0, 0, 0, 0);
        return;
    }
    const uint32_t a = blockIdx.x * MC_THREADS + threadIdx.x;
    Cell c;
    c.nv = 0; c.nt = 0; c.ambiguous = 0; c.interior = 0;
    if (a < nact) {
        int i, j, k;
        cell_coords(p, (int64_t)cells[a].lin, i, j, k);
        classify(p, i, j, k, c);
    }
    unsigned ev, ef, tv, tf;
    block_scan2((unsigned)c.nv, (unsigned)c.nt, ev, ef, tv, tf);
    if (a < nact) { cells[a].vbase = ev; cells[a].fbase = ef; cells[a].cls = c.entry < 0 ? 0xffffu : ((uint32_t)c.entry | (c.owned << 16)); }
    if (threadIdx.x == 0) block_tot[blockIdx.x] = make_uint4(0, tv, tf, 0);
    if (c.ambiguous) atomicAdd(n_amb, 1ull);
    if (c.interior) atomicAdd(n_amb + 8, 1ull);
    if (c.interior == 2) atomicAdd(n_amb + 9, 1ull);
}

// single-CTA exclusive scan of (x, y, z) over the per-block totals: every thread owns SCAN_ITEMS consecutive entries, so
// 32 K entries take four rounds of (warp scan, one barrier pair) instead of thirty-two
constexpr int SCAN_ITEMS = 8;
__global__ void __launch_bounds__(1024) mc_scan_blocks3_kernel(uint4 *block_tot, int64_t nblocks, unsigned long long *totals,
                                                               const unsigned long long *nitems, int per_block)
{
    if (nitems) {                                    // the list was sized by an estimate: only ceil(n / per_block) blocks hold cells
        const int64_t used = (int64_t)((*nitems + per_block - 1) / per_block);
        nblocks = used < nblocks ? used : nblocks;
    }
    __shared__ unsigned long long wsum[3][32];
    __shared__ unsigned long long carry[3];
    if (threadIdx.x < 3) carry[threadIdx.x] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t base = 0; base < nblocks; base += 1024 * SCAN_ITEMS) {
        const int64_t first = base + (int64_t)threadIdx.x * SCAN_ITEMS;
        uint4 v[SCAN_ITEMS];
        unsigned long long a[3] = {0, 0, 0};
#pragma unroll
        for (int q = 0; q < SCAN_ITEMS; ++q) {
            v[q] = first + q < nblocks ? block_tot[first + q] : make_uint4(0, 0, 0, 0);
            a[0] += v[q].x; a[1] += v[q].y; a[2] += v[q].z;
        }
        const unsigned long long mine[3] = {a[0], a[1], a[2]};
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
        unsigned long long run[3] = {off[0] + a[0] - mine[0], off[1] + a[1] - mine[1], off[2] + a[2] - mine[2]};
#pragma unroll
        for (int q = 0; q < SCAN_ITEMS; ++q) {
            if (first + q < nblocks) block_tot[first + q] = make_uint4((unsigned)run[0], (unsigned)run[1], (unsigned)run[2], 0);
            run[0] += v[q].x; run[1] += v[q].y; run[2] += v[q].z;
        }
        __syncthreads();
        if (threadIdx.x == 1023)
            for (int c = 0; c < 3; ++c) carry[c] = off[c] + a[c];
        __syncthreads();
    }
    if (threadIdx.x == 0) { totals[0] = carry[1]; totals[1] = carry[2]; totals[2] = carry[0]; }
}

// one thread per active cell (four threads per cell, one per owned edge, measured SLOWER: 465 vs 290 us at 512^3 --
// the cell record, the rank table and the gradient taps are then fetched four times)
__global__ void __launch_bounds__(128) mc_list_verts_kernel(McParams p, const CellRec *__restrict__ cells, const uint4 *__restrict__ boff,
                                                            uint32_t nact, McOut o)
{
    const uint32_t a = blockIdx.x * 128 + threadIdx.x;
    if (a >= nact) return;
    CellRec r = cells[a];
    if ((r.cls >> 16) == 0 || (r.cls & 0xffffu) == 0xffffu) return;      // no vertex of its own
    r.vbase += boff[a / MC_THREADS].y;
    int i, j, k;
    cell_coords(p, (int64_t)r.lin, i, j, k);
    Cell c;
    c.entry = (int)(r.cls & 0xffffu);
    c.owned = r.cls >> 16;
#pragma unroll
    for (int q = 0; q < 8; ++q)
        c.d[q] = __dsub_rn((double)__ldg(p.vol + node_lin(p, i + c_corner_off[3 * q], j + c_corner_off[3 * q + 1], k + c_corner_off[3 * q + 2])), p.level);
    emit_cell_verts(p, o, i, j, k, c, (int64_t)r.vbase);
}

__global__ void __launch_bounds__(128) mc_list_faces_kernel(McParams p, const CellRec *__restrict__ cells, const uint4 *__restrict__ boff,
                                                            uint32_t nact, const int32_t *__restrict__ vid, const int32_t *__restrict__ seam_in,
                                                            int64_t id_offset, int32_t *__restrict__ faces, unsigned long long *seam_bad)
{
    const uint32_t a = blockIdx.x * 128 + threadIdx.x;
    if (a >= nact) return;
    CellRec r = cells[a];
    if ((r.cls & 0xffffu) == 0xffffu) return;
    const uint4 bo = boff[a / MC_THREADS];
    r.vbase += bo.y; r.fbase += bo.z;
    int i, j, k;
    cell_coords(p, (int64_t)r.lin, i, j, k);
    Cell c;                                                    // the volume is not read: entry and ownership come from the count pass
    c.entry = (int)(r.cls & 0xffffu);
    c.owned = r.cls >> 16;
    c.nt = p.tb.ntri[c.entry];
    if (c.nt) emit_cell_faces(p, i, j, k, c, (int64_t)r.vbase, (int64_t)r.fbase, vid, seam_in, id_offset, faces, seam_bad);
}

struct TableBlob {
    uint8_t amb_mask[256];
    uint16_t var_base[256];
    uint8_t ntri[MC_NUM_ENTRIES];
    uint16_t tri_off[MC_NUM_ENTRIES];
    uint8_t tri_edges[MC_NUM_TRI_IDX];
    uint8_t edge_rank[MC_NUM_ENTRIES * 13];
    uint8_t ntest[MC_NUM_ENTRIES];
    uint16_t test_off[MC_NUM_ENTRIES];
    uint8_t test_corners[MC_NUM_TESTS * 8];
    uint8_t test_sign[MC_NUM_TESTS];
    uint16_t test_target[MC_NUM_TESTS];
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
    p.tb.ntest = b->ntest; p.tb.test_off = b->test_off; p.tb.test_corners = b->test_corners;
    p.tb.test_sign = b->test_sign; p.tb.test_target = b->test_target;
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
    memcpy(h->ntest, mc_ntest, sizeof(h->ntest));
    memcpy(h->test_off, mc_test_off, sizeof(h->test_off));
    memcpy(h->test_corners, mc_test_corners, sizeof(h->test_corners));
    memcpy(h->test_sign, mc_test_sign, sizeof(h->test_sign));
    memcpy(h->test_target, mc_test_target, sizeof(h->test_target));
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

static int mc_count_impl(surs_ctx *ctx, const float *vol, const double *vol64, const int res[3], float level, int flags,
                         int64_t *n_verts, int64_t *n_faces, int64_t *n_ambiguous, void *stream);

extern "C" int surs_mc_count(surs_ctx *ctx, const float *vol, const int res[3], float level, int flags,
                             int64_t *n_verts, int64_t *n_faces, int64_t *n_ambiguous, void *stream)
{
    return mc_count_impl(ctx, vol, nullptr, res, level, flags, n_verts, n_faces, n_ambiguous, stream);
}

extern "C" int surs_mc_count_f64(surs_ctx *ctx, const double *vol64, float *vol32, const int res[3], float level, int flags,
                                 int64_t *n_verts, int64_t *n_faces, int64_t *n_ambiguous, void *stream)
{
    if (ctx && !vol64) SURS_FAIL(ctx, "surs_mc_count_f64: null source volume");
    return mc_count_impl(ctx, vol32, vol64, res, level, flags, n_verts, n_faces, n_ambiguous, stream);
}

extern "C" int surs_mc_value_range(surs_ctx *ctx, float *vmin, float *vmax)
{
    if (!ctx) return 1;
    if (!ctx->mc_vol) SURS_FAIL(ctx, "surs_mc_value_range: call surs_mc_count first");
    if (vmin) *vmin = ctx->mc_vmin;
    if (vmax) *vmax = ctx->mc_vmax;
    return 0;
}

// vol64 != NULL: the source is float64 and `vol` receives its float32 copy (made by the same pass that takes the bits)
static int mc_count_impl(surs_ctx *ctx, const float *vol, const double *vol64, const int res[3], float level, int flags,
                         int64_t *n_verts, int64_t *n_faces, int64_t *n_ambiguous, void *stream)
{
    if (!ctx) return 1;
    SURS_NVTX("surs_mc_count");
    cudaStream_t st = (cudaStream_t)stream;
    SURS_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!vol || res[0] < 2 || res[1] < 2 || res[2] < 2) SURS_FAIL(ctx, "surs_mc_count: volume needs at least 2 nodes per axis");
    float *vol_w = const_cast<float *>(vol);
    unsigned *range = reinterpret_cast<unsigned *>(ctx->counter + 12);
    {
        const unsigned init[2] = {0xffffffffu, 0u};
        SURS_CUDA(ctx, cudaMemcpyAsync(range, init, sizeof(init), cudaMemcpyHostToDevice, st));
    }
    const int64_t nnode = (int64_t)res[0] * res[1] * res[2];
    if (3 * nnode >= ((int64_t)1 << 31)) SURS_FAIL(ctx, "surs_mc_count: volume too large for 32-bit edge ids (3*res^3 < 2^31)");
    ctx->mc_vol = vol;
    memcpy(ctx->mc_res, res, sizeof(int) * 3);
    ctx->mc_level = level;
    ctx->mc_flags = flags;
    SURS_CUDA(ctx, cudaMemsetAsync(ctx->counter, 0, 64, st));
    SURS_CUDA(ctx, cudaMemsetAsync(ctx->counter + 9, 0, 16, st));
    ctx->mc_count_stream = st;
    McParams p = make_params(ctx);
    static const bool no_fast = getenv("SURS_MC_SLOW") != nullptr;
    ctx->mc_fast = (res[2] % 4 == 0) && ((uintptr_t)vol % 16 == 0) && !no_fast;
    if (ctx->mc_fast) {
        const int W = (res[2] + 31) / 32;
        const int64_t nrows = (int64_t)res[0] * res[1];
        const int64_t nthr = nrows * W;
        if (nthr >= ((int64_t)1 << 32) || nnode >= ((int64_t)1 << 32)) SURS_FAIL(ctx, "surs_mc_count: volume too large for 32-bit cell ids");
        const int64_t nbA = (nthr + MC_THREADS - 1) / MC_THREADS;
        if (surs_ensure(ctx, (void **)&ctx->mc_bits, &ctx->mc_bits_cap, sizeof(uint32_t) * (size_t)nthr)) return 1;
        if (surs_ensure(ctx, (void **)&ctx->mc_block_tot, &ctx->mc_block_cap, sizeof(uint4) * (size_t)nbA)) return 1;
        uint32_t *bits = reinterpret_cast<uint32_t *>(ctx->mc_bits);
        int64_t sb = (nrows + MC_SIGN_WARPS - 1) / MC_SIGN_WARPS;
        if (sb > (int64_t)ctx->sm_count * 8 * 8) sb = (int64_t)ctx->sm_count * 8 * 8;   // 8 resident CTAs per SM, 8 rows in flight per warp slot
        const unsigned sblocks = (unsigned)sb;
        if (vol64) mc_sign_kernel<double, true><<<sblocks, MC_SIGN_WARPS * 32, 0, st>>>(vol64, vol_w, level, nrows, res[2], W, bits, range);
        else mc_sign_kernel<float, false><<<sblocks, MC_SIGN_WARPS * 32, 0, st>>>(vol, nullptr, level, nrows, res[2], W, bits, range);
        SURS_LAUNCH_CHECK(ctx, "mc_sign_kernel");
        uint4 *btA = reinterpret_cast<uint4 *>(ctx->mc_block_tot);
        mc_bits_kernel<false><<<(unsigned)nbA, MC_THREADS, 0, st>>>(p, bits, W, (uint32_t)nthr, btA, nullptr, nullptr, 0);
        SURS_LAUNCH_CHECK(ctx, "mc_bits_kernel<count>");
        mc_scan_blocks3_kernel<<<1, 1024, 0, st>>>(btA, nbA, ctx->counter + 5, nullptr, 0);       // counter[7] = number of active cells
        SURS_LAUNCH_CHECK(ctx, "mc_scan_blocks3_kernel");
        // compaction, per-cell classification and the second scan run on an ESTIMATE of the list length (the last
        // volume's, at least 1/16 of the cells), so the whole count phase needs ONE host synchronisation; if the
        // estimate was too small the kernels wrote nothing and the phase is repeated with the exact length
        unsigned long long host[8];
        int64_t cap = ctx->mc_cap_hint > 0 ? ctx->mc_cap_hint + ctx->mc_cap_hint / 4 : 0;
        if (cap < nnode / 16) cap = nnode / 16;
        if (cap < 4096) cap = 4096;
        if (cap > nnode) cap = nnode;
        for (int attempt = 0; attempt < 2; ++attempt) {
            const int64_t nbB = (cap + MC_THREADS - 1) / MC_THREADS;
            if (surs_ensure(ctx, (void **)&ctx->mc_cells, &ctx->mc_cells_cap, sizeof(CellRec) * (size_t)cap)) return 1;
            if (surs_ensure(ctx, (void **)&ctx->mc_cell_tot, &ctx->mc_cell_tot_cap, sizeof(uint4) * (size_t)nbB)) return 1;
            CellRec *cells = reinterpret_cast<CellRec *>(ctx->mc_cells);
            uint4 *btB = reinterpret_cast<uint4 *>(ctx->mc_cell_tot);
            mc_bits_kernel<true><<<(unsigned)nbA, MC_THREADS, 0, st>>>(p, bits, W, (uint32_t)nthr, btA, cells, ctx->counter + 7, (uint32_t)cap);
            SURS_LAUNCH_CHECK(ctx, "mc_bits_kernel<compact>");
            mc_cell_kernel<<<(unsigned)nbB, MC_THREADS, 0, st>>>(p, cells, ctx->counter + 7, (uint32_t)cap, btB, ctx->counter + 1);
            SURS_LAUNCH_CHECK(ctx, "mc_cell_kernel");
            mc_scan_blocks3_kernel<<<1, 1024, 0, st>>>(btB, nbB, ctx->counter + 2, ctx->counter + 7, MC_THREADS);
            SURS_LAUNCH_CHECK(ctx, "mc_scan_blocks3_kernel");
            SURS_CUDA(ctx, cudaMemcpyAsync(host, ctx->counter, sizeof(host), cudaMemcpyDeviceToHost, st));
            SURS_CUDA(ctx, cudaMemcpyAsync(ctx->mc_range_host, range, 2 * sizeof(unsigned), cudaMemcpyDeviceToHost, st));
            SURS_CUDA(ctx, cudaStreamSynchronize(st));
            ctx->mc_vmin = key2f(ctx->mc_range_host[0]);
            ctx->mc_vmax = key2f(ctx->mc_range_host[1]);
            ctx->mc_nact = (int64_t)host[7];
            if (ctx->mc_nact <= cap) break;
            cap = ctx->mc_nact;                                                       // estimate too small: exact length now
            host[7] = 0;
        }
        ctx->mc_cap_hint = ctx->mc_nact;
        ctx->mc_nv = ctx->mc_nact > 0 ? (int64_t)host[2] : 0;
        ctx->mc_nf = ctx->mc_nact > 0 ? (int64_t)host[3] : 0;
        if (ctx->mc_nv >= ((int64_t)1 << 31) || ctx->mc_nf >= ((int64_t)1 << 31)) SURS_FAIL(ctx, "surs_mc_count: mesh too large for int32 indices");
        if (n_verts) *n_verts = ctx->mc_nv;
        if (n_faces) *n_faces = ctx->mc_nf;
        if (n_ambiguous) *n_ambiguous = (int64_t)host[1];
        return 0;
    }
    const int64_t nblocks = (nnode + MC_THREADS - 1) / MC_THREADS;
    if (surs_ensure(ctx, (void **)&ctx->mc_block_tot, &ctx->mc_block_cap, sizeof(uint2) * (size_t)nblocks)) return 1;
    {
        const unsigned rb = (unsigned)(nblocks < ctx->sm_count * 16 ? nblocks : ctx->sm_count * 16);
        if (vol64) mc_range_kernel<double, true><<<rb, MC_THREADS, 0, st>>>(vol64, vol_w, nnode, range);
        else mc_range_kernel<float, false><<<rb, MC_THREADS, 0, st>>>(vol, nullptr, nnode, range);
        SURS_LAUNCH_CHECK(ctx, "mc_range_kernel");
    }
    mc_count_kernel<<<(unsigned)nblocks, MC_THREADS, 0, st>>>(p, nnode, reinterpret_cast<uint2 *>(ctx->mc_block_tot), ctx->counter + 1);
    SURS_LAUNCH_CHECK(ctx, "mc_count_kernel");
    mc_scan_blocks_kernel<<<1, 1024, 0, st>>>(reinterpret_cast<uint2 *>(ctx->mc_block_tot), nblocks, ctx->counter + 2);
    SURS_LAUNCH_CHECK(ctx, "mc_scan_blocks_kernel");
    unsigned long long host[4];
    SURS_CUDA(ctx, cudaMemcpyAsync(host, ctx->counter, sizeof(host), cudaMemcpyDeviceToHost, st));
    SURS_CUDA(ctx, cudaMemcpyAsync(ctx->mc_range_host, range, 2 * sizeof(unsigned), cudaMemcpyDeviceToHost, st));
    SURS_CUDA(ctx, cudaStreamSynchronize(st));
    ctx->mc_vmin = key2f(ctx->mc_range_host[0]);
    ctx->mc_vmax = key2f(ctx->mc_range_host[1]);
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
    SURS_NVTX("surs_mc_emit_verts");
    if (!ctx) return 1;
    cudaStream_t st = (cudaStream_t)stream;
    SURS_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!ctx->mc_vol) SURS_FAIL(ctx, "surs_mc_emit_verts: call surs_mc_count first");
    if (ctx->mc_nv > 0 && !verts && !(verts_world && mat)) SURS_FAIL(ctx, "surs_mc_emit_verts: null output (verts, or verts_world with mat)");
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
                                                                                     reinterpret_cast<const uint4 *>(ctx->mc_cell_tot),
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
    SURS_NVTX("surs_mc_emit_faces");
    if (!ctx) return 1;
    cudaStream_t st = (cudaStream_t)stream;
    SURS_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!ctx->mc_vol || !ctx->mc_vid) SURS_FAIL(ctx, "surs_mc_emit_faces: call surs_mc_count and surs_mc_emit_verts first");
    if ((ctx->mc_flags & SURS_MC_LOWER_FOREIGN) && !seam_in) SURS_FAIL(ctx, "surs_mc_emit_faces: seam_in required with SURS_MC_LOWER_FOREIGN");
    if (ctx->mc_nf > 0 && !faces) SURS_FAIL(ctx, "surs_mc_emit_faces: null output");
    const int64_t nnode = (int64_t)ctx->mc_res[0] * ctx->mc_res[1] * ctx->mc_res[2];
    const int64_t nblocks = (nnode + MC_THREADS - 1) / MC_THREADS;
    McParams p = make_params(ctx);
    SURS_CUDA(ctx, cudaMemsetAsync(ctx->counter + 8, 0, sizeof(unsigned long long), st));
    ctx->mc_faces_stream = st;
    if (ctx->mc_nf > 0 && ctx->mc_fast) {
        mc_list_faces_kernel<<<(unsigned)((ctx->mc_nact + 127) / 128), 128, 0, st>>>(p, reinterpret_cast<const CellRec *>(ctx->mc_cells),
                                                                                     reinterpret_cast<const uint4 *>(ctx->mc_cell_tot),
                                                                                     (uint32_t)ctx->mc_nact, ctx->mc_vid, seam_in,
                                                                                     ctx->mc_id_offset, faces, ctx->counter + 8);
        SURS_LAUNCH_CHECK(ctx, "mc_list_faces_kernel");
    } else if (ctx->mc_nf > 0) {
        mc_emit_faces_kernel<<<(unsigned)nblocks, MC_THREADS, 0, st>>>(p, nnode, reinterpret_cast<const uint2 *>(ctx->mc_block_tot), ctx->mc_vid,
                                                                      seam_in, ctx->mc_id_offset, faces, ctx->counter + 8);
        SURS_LAUNCH_CHECK(ctx, "mc_emit_faces_kernel");
    }
    return 0;
}

extern "C" int surs_mc_interior_stats(surs_ctx *ctx, int64_t *n_interior_ambiguous, int64_t *n_tunnels)
{
    if (!ctx) return 1;
    unsigned long long n[2] = {0, 0};
    SURS_CUDA(ctx, cudaSetDevice(ctx->device));
    SURS_CUDA(ctx, cudaMemcpyAsync(n, ctx->counter + 9, sizeof(n), cudaMemcpyDeviceToHost, (cudaStream_t)ctx->mc_count_stream));
    SURS_CUDA(ctx, cudaStreamSynchronize((cudaStream_t)ctx->mc_count_stream));
    if (n_interior_ambiguous) *n_interior_ambiguous = (int64_t)n[0];
    if (n_tunnels) *n_tunnels = (int64_t)n[1];
    return 0;
}

extern "C" int64_t surs_mc_seam_violations(surs_ctx *ctx)
{
    if (!ctx) return -1;
    unsigned long long n = 0;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return -1;
    if (cudaMemcpyAsync(&n, ctx->counter + 8, sizeof(n), cudaMemcpyDeviceToHost, (cudaStream_t)ctx->mc_faces_stream) != cudaSuccess) return -1;
    if (cudaStreamSynchronize((cudaStream_t)ctx->mc_faces_stream) != cudaSuccess) return -1;
    return (int64_t)n;
}

extern "C" int surs_mc_emit(surs_ctx *ctx, const double *mat, float *verts, double *verts_world,
                            int32_t *faces, float *normals, float *values, void *stream)
{
    if (surs_mc_emit_verts(ctx, mat, verts, verts_world, normals, values, 0, 0, nullptr, stream)) return 1;
    return surs_mc_emit_faces(ctx, faces, nullptr, stream);
}
