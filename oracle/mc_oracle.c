/* CPU marching cubes -- TEST INFRASTRUCTURE ONLY (the checker for the CUDA kernels in
 * csrc/mc.cu; also the "cpu_baseline" marching-cubes leg of bench.py).  Never linked
 * into, or called from, the product path.
 *
 * PARITY UNPINNED.  The reference's marching cubes is scikit-image 0.17.2
 * `skimage.measure.marching_cubes_lewiner(volume, 0.5)` (reference call sites
 * lib/mesh_util.py:40,45; pin environment.yml:333).  That package is a third-party
 * Cython extension which is neither under /root/reference nor installed in this image,
 * and the reference holds no tests / golden meshes for it.  This file restates the
 * published algorithm (Lewiner et al. 2003, as driven by skimage's defaults
 * spacing=(1,1,1), gradient_direction='descent', step_size=1, allow_degenerate=True):
 *
 *   - volume used as float32; cells scanned with array axis 0 outermost, axis 2 innermost;
 *   - corner / edge numbering of Lewiner (== Bourke) with x = axis 2, y = axis 1, z = axis 0;
 *   - case bit i set iff value_i - level > 0;
 *   - ambiguous faces resolved by the asymptotic decider (Lewiner's face test);
 *   - vertices shared between cells, numbered in order of FIRST USE while scanning cells
 *     and, inside a cell, its triangle list; faces appended in scan order;
 *   - vertex on an edge = mean of the two end points weighted by 1/(FLT_EPSILON + |v-level|),
 *     evaluated in double, stored float32, reported in (axis0, axis1, axis2) order;
 *   - degenerate triangles kept.
 *
 * Deviations (documented in csrc/gen_mc_tables.py): derived case tables (own loop / band
 * triangulation; the centre vertex is used where a loop cannot be triangulated without a
 * diagonal lying in a cube face; the interior (tunnel) test is derived from the trilinear
 * sections rather than Lewiner's per-case reference edges).  The tables are shared with the
 * kernels (mc_tables.h): what checks THEM is the table-free oracle/mc_check.py.  `normals` / `values` are
 * API-shape obligations only (the reference's caller discards them, lib/train_util.py:72):
 * normals = normalised volume gradient (central differences, linear along the edge),
 * values = max of the edge's two end values.
 *
 * This is deliberately the *literal* sequential formulation (an edge -> vertex-id
 * cache filled on first use); the CUDA kernels derive the same numbering from an
 * edge-ownership rule + prefix sums, so agreement is a real cross-check.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -shared -fPIC).
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "mc_tables.h"

#define IDX(i, j, k) (((int64_t)(i) * R1 + (j)) * R2 + (k))

/* Interior (tunnel) test: csrc/gen_mc_tables.py; same operation order as csrc/mc.cu interior_test. */
static int interior_test(const double d[8], const uint8_t *cn, int s)
{
    double a0 = d[cn[0]], a1 = d[cn[1]], b0 = d[cn[2]], b1 = d[cn[3]], c0 = d[cn[4]], c1 = d[cn[5]], d0 = d[cn[6]], d1 = d[cn[7]];
    double dA = a1 - a0, dB = b1 - b0, dC = c1 - c0, dD = d1 - d0;
    double qa = dA * dC - dB * dD;
    double qb = (a0 * dC + c0 * dA) - (b0 * dD + d0 * dB);
    if (!(qa < 0.0)) return 0;
    double t = -qb / (2.0 * qa);
    if (!(t > 0.0 && t < 1.0)) return 0;
    double At = a0 + dA * t, Bt = b0 + dB * t, Ct = c0 + dC * t, Dt = d0 + dD * t;
    if (s) {
        if (!(At > 0.0 && Ct > 0.0)) return 0;
    } else if (At > 0.0 || Ct > 0.0) {
        return 0;
    }
    return At * Ct - Bt * Dt > 0.0;
}

static int64_t g_interior = 0, g_tunnels = 0;      /* statistics of the last run (single-threaded test code) */
void mc_oracle_interior_stats(int64_t *n_interior, int64_t *n_tunnels) { *n_interior = g_interior; *n_tunnels = g_tunnels; }

static int cell_entry(const float *vol, int R1, int R2, int i, int j, int k, double level, double d[8])
{
    int c, f, cas = 0;
    for (c = 0; c < 8; ++c) {
        d[c] = (double)vol[IDX(i + mc_corner_off[3 * c], j + mc_corner_off[3 * c + 1], k + mc_corner_off[3 * c + 2])] - level;
        if (d[c] > 0.0) cas |= 1 << c;
    }
    int mask = mc_amb_mask[cas], var = 0, nb = 0;
    for (f = 0; f < 6; ++f) {
        if (!((mask >> f) & 1)) continue;
        const uint8_t *fc = &mc_face_corner[4 * f];
        double p02 = d[fc[0]] * d[fc[2]];
        double p13 = d[fc[1]] * d[fc[3]];
        int connect = (d[fc[0]] > 0.0) ? (p02 > p13) : (p13 > p02);
        var |= connect << nb;
        ++nb;
    }
    int ent = mc_var_base[cas] + var;
    if (mc_ntest[ent]) {
        ++g_interior;
        for (int q = 0; q < mc_ntest[ent]; ++q) {
            int tq = mc_test_off[ent] + q;
            if (interior_test(d, &mc_test_corners[8 * tq], mc_test_sign[tq])) {
                ++g_tunnels;
                return mc_test_target[tq];
            }
        }
    }
    return ent;
}

static double node_grad(const float *vol, int R0, int R1, int R2, int i, int j, int k, int axis)
{
    int n = axis == 0 ? R0 : (axis == 1 ? R1 : R2);
    int p = axis == 0 ? i : (axis == 1 ? j : k);
    int lo = p > 0 ? p - 1 : p, hi = p < n - 1 ? p + 1 : p;
    int il = i, jl = j, kl = k, ih = i, jh = j, kh = k;
    if (axis == 0) { il = lo; ih = hi; } else if (axis == 1) { jl = lo; jh = hi; } else { kl = lo; kh = hi; }
    double dv = (double)vol[IDX(ih, jh, kh)] - (double)vol[IDX(il, jl, kl)];
    return (hi - lo) == 2 ? dv * 0.5 : dv;
}

/* position along an edge whose lower end sits at coordinate `base`: the mean of the two end
 * points weighted by 1/(FLT_EPSILON + |v - level|) */
static double edge_point(double base, double da, double db)
{
    double wa = 1.0 / ((double)FLT_EPSILON + fabs(da));
    double wb = 1.0 / ((double)FLT_EPSILON + fabs(db));
    return (base * wa + (base + 1.0) * wb) / (wa + wb);
}

/* Lewiner's "13th vertex": mean of the cell's edge vertices (edges in id order). */
static void emit_centre(const float *vol, int R0, int R1, int R2, int i, int j, int k, double level,
                        const double d[8], float *vert, float *normal, float *value)
{
    double s[3] = {0.0, 0.0, 0.0};
    int n = 0;
    for (int e = 0; e < 12; ++e) {
        int a = mc_edge_corner[2 * e], b = mc_edge_corner[2 * e + 1];
        if ((d[a] > 0.0) == (d[b] > 0.0)) continue;
        int axis = mc_edge_axis[e];
        int lo = (mc_corner_off[3 * a + axis] == 0) ? a : b, hi = (lo == a) ? b : a;
        double p[3] = {(double)(i + mc_edge_base[3 * e]), (double)(j + mc_edge_base[3 * e + 1]), (double)(k + mc_edge_base[3 * e + 2])};
        p[axis] = edge_point(p[axis], d[lo], d[hi]);
        s[0] += p[0]; s[1] += p[1]; s[2] += p[2];
        ++n;
    }
    double g[3] = {0.0, 0.0, 0.0}, vmax = -INFINITY, nn = 0.0;
    for (int c = 0; c < 8; ++c) {
        int ci = i + mc_corner_off[3 * c], cj = j + mc_corner_off[3 * c + 1], ck = k + mc_corner_off[3 * c + 2];
        for (int a = 0; a < 3; ++a) g[a] += node_grad(vol, R0, R1, R2, ci, cj, ck, a);
        double v = (double)vol[IDX(ci, cj, ck)];
        if (v > vmax) vmax = v;
    }
    for (int a = 0; a < 3; ++a) { vert[a] = (float)(s[a] / (double)n); nn += g[a] * g[a]; }
    nn = sqrt(nn);
    for (int a = 0; a < 3; ++a) normal[a] = nn > 0.0 ? (float)(g[a] / nn) : 0.0f;
    *value = (float)vmax;
}

/* Runs the algorithm.  If verts == NULL only counts.  Returns 0, or -1 on allocation failure. */
int mc_oracle_run(const float *vol, int R0, int R1, int R2, float level_f,
                  float *verts, int32_t *faces, float *normals, float *values,
                  int64_t *n_verts, int64_t *n_faces, int64_t *n_ambiguous_cells)
{
    const double level = (double)level_f;
    int64_t nnode = (int64_t)R0 * R1 * R2, nv = 0, nf = 0, namb = 0;
    g_interior = g_tunnels = 0;
    int32_t *vid = (int32_t *)malloc(sizeof(int32_t) * 3 * (size_t)nnode);
    if (!vid) return -1;
    memset(vid, 0xff, sizeof(int32_t) * 3 * (size_t)nnode);
    for (int i = 0; i + 1 < R0; ++i)
        for (int j = 0; j + 1 < R1; ++j)
            for (int k = 0; k + 1 < R2; ++k) {
                double d[8];
                int cas_probe = 0;
                /* cheap reject first */
                {
                    int c, pos = 0;
                    for (c = 0; c < 8; ++c)
                        pos += ((double)vol[IDX(i + mc_corner_off[3 * c], j + mc_corner_off[3 * c + 1], k + mc_corner_off[3 * c + 2])] - level) > 0.0;
                    cas_probe = pos;
                }
                if (cas_probe == 0 || cas_probe == 8) continue;
                int ent = cell_entry(vol, R1, R2, i, j, k, level, d);
                int nt = mc_ntri[ent];
                {
                    int cas = 0, c;
                    for (c = 0; c < 8; ++c) if (d[c] > 0.0) cas |= 1 << c;
                    if (mc_amb_mask[cas]) ++namb;
                }
                const uint8_t *te = &mc_tri_edges[3 * mc_tri_off[ent]];
                int32_t centre_id = -1;
                for (int t = 0; t < nt; ++t) {
                    int32_t tri[3];
                    for (int s = 0; s < 3; ++s) {
                        int e = te[3 * t + s];
                        if (e == MC_CENTRE) {
                            if (centre_id < 0) {
                                centre_id = (int32_t)nv;
                                if (verts) emit_centre(vol, R0, R1, R2, i, j, k, level, d, &verts[3 * nv], &normals[3 * nv], &values[nv]);
                                ++nv;
                            }
                            tri[s] = centre_id;
                            continue;
                        }
                        int axis = mc_edge_axis[e];
                        int bi = i + mc_edge_base[3 * e], bj = j + mc_edge_base[3 * e + 1], bk = k + mc_edge_base[3 * e + 2];
                        int64_t slot = 3 * IDX(bi, bj, bk) + axis;
                        if (vid[slot] < 0) {
                            vid[slot] = (int32_t)nv;
                            if (verts) {
                                int ei = bi + (axis == 0), ej = bj + (axis == 1), ek = bk + (axis == 2);
                                double va = (double)vol[IDX(bi, bj, bk)], vb = (double)vol[IDX(ei, ej, ek)];
                                double base = (double)(axis == 0 ? bi : (axis == 1 ? bj : bk));
                                double x = edge_point(base, va - level, vb - level);
                                double tt = x - base;
                                float p[3] = {(float)bi, (float)bj, (float)bk};
                                p[axis] = (float)x;
                                memcpy(&verts[3 * nv], p, sizeof(p));
                                double g[3], nn = 0.0;
                                for (int a = 0; a < 3; ++a) {
                                    double ga = node_grad(vol, R0, R1, R2, bi, bj, bk, a);
                                    double gb = node_grad(vol, R0, R1, R2, ei, ej, ek, a);
                                    g[a] = ga + tt * (gb - ga);
                                    nn += g[a] * g[a];
                                }
                                nn = sqrt(nn);
                                for (int a = 0; a < 3; ++a) normals[3 * nv + a] = nn > 0.0 ? (float)(g[a] / nn) : 0.0f;
                                values[nv] = (float)(va > vb ? va : vb);
                            }
                            ++nv;
                        }
                        tri[s] = vid[slot];
                    }
                    if (faces) memcpy(&faces[3 * nf], tri, sizeof(tri));
                    ++nf;
                }
            }
    free(vid);
    *n_verts = nv;
    *n_faces = nf;
    if (n_ambiguous_cells) *n_ambiguous_cells = namb;
    return 0;
}
