#!/usr/bin/env python
"""Generates the marching-cubes case tables (csrc/mc_tables.h) from first principles.

Why generated: the reference's marching cubes is scikit-image 0.17.2's
``marching_cubes_lewiner`` (lib/mesh_util.py:40,45), a third-party Cython module
that is neither vendored in /root/reference nor installed here, and Lewiner's
hand-made 33-case tables are not available offline.  Instead of recalling ~2k
lines of tables, the triangulation is *derived*:

  * corner / edge numbering as in Lewiner (== Bourke): corner c at offsets
    (dx,dy,dz) with x = array axis 2, y = axis 1, z = axis 0 (skimage scans
    ``im[z][y][x]``);
  * a corner is "positive" iff value - level > 0 (strict);
  * on every cube face the iso-line is traced between the sign-changing edges;
    a face with 4 sign changes is ambiguous and is resolved at run time by the
    asymptotic decider (Lewiner's face test): the positive corners are joined
    across the face iff  a*c > b*d  (a,c the positive, b,d the negative corner
    values minus level).  The tables therefore hold one triangulation per
    (case, decider bits of its ambiguous faces);
  * the face segments are linked into closed loops (ordered by their smallest
    edge id).  A loop is triangulated without any diagonal that lies in a cube
    face (two vertices whose edges share a face): such a diagonal would coincide
    with geometry of the neighbouring cube and make the mesh non-manifold.  The
    first such triangulation in a fixed search order is used; where none exists
    (116 of the 1026 loops: the long 8/9/12-gons of Lewiner's cases 7,10,12,13)
    the loop is fanned around a centre vertex, id 12 -- Lewiner's "13th vertex",
    placed at the mean of the cell's edge vertices.

Because both cubes sharing a face evaluate the same decider on the same four
values, the result is watertight.  Known deviations from Lewiner's MC33 (the
part of the parity that stays UNPINNED): no interior (tunnel) test, and the
triangulation of a loop (which diagonals, when the centre vertex is used) is ours.

Orientation: in output coordinates (axis0, axis1, axis2) taken as a right-handed
frame, (v1-v0)x(v2-v0) points towards increasing volume values (what skimage's
gradient_direction='descent' yields for an occupancy field; the reference then
swaps the winding when it writes the OBJ, lib/mesh_util.py:60).

Run:  python oracle/gen_mc_tables.py   (rewrites csrc/mc_tables.h; output is committed)
"""
import itertools
import os

import numpy as np

# corner -> (d_axis0, d_axis1, d_axis2); Bourke (dx,dy,dz) with x=axis2, y=axis1, z=axis0
BOURKE = [(0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0), (0, 0, 1), (1, 0, 1), (1, 1, 1), (0, 1, 1)]
CORNER = [(dz, dy, dx) for (dx, dy, dz) in BOURKE]
EDGE = [(0, 1), (1, 2), (2, 3), (3, 0), (4, 5), (5, 6), (6, 7), (7, 4), (0, 4), (1, 5), (2, 6), (3, 7)]


def edge_id(a, b):
    for e, (p, q) in enumerate(EDGE):
        if (p, q) == (a, b) or (p, q) == (b, a):
            return e
    raise KeyError((a, b))


def build_faces():
    """6 faces (axis0-, axis0+, axis1-, axis1+, axis2-, axis2+), corners CCW seen from outside."""
    faces = []
    P = np.array(CORNER, dtype=float)
    for axis in range(3):
        for side in (0, 1):
            cs = [c for c in range(8) if CORNER[c][axis] == side]
            n = np.zeros(3)
            n[axis] = 1.0 if side else -1.0
            ctr = P[cs].mean(axis=0)
            # orthonormal in-plane frame (e1, e2) with e1 x e2 = n
            e1 = np.zeros(3)
            e1[(axis + 1) % 3] = 1.0
            e2 = np.cross(n, e1)
            ang = [np.arctan2(np.dot(P[c] - ctr, e2), np.dot(P[c] - ctr, e1)) for c in cs]
            order = [c for _, c in sorted(zip(ang, cs))]
            # rotate so the smallest corner id leads (canonical)
            k = order.index(min(order))
            order = order[k:] + order[:k]
            a, b, c = P[order[0]], P[order[1]], P[order[2]]
            assert np.dot(np.cross(b - a, c - b), n) > 0
            faces.append(order)
    return faces


FACES = build_faces()


def face_is_ambiguous(case, face):
    s = [(case >> c) & 1 for c in face]
    return s[0] == s[2] and s[1] == s[3] and s[0] != s[1]


def triangulate(case, connect_bits):
    """connect_bits[f] = 1 -> positive corners joined across ambiguous face f."""
    nxt = {}
    for f, face in enumerate(FACES):
        s = [(case >> c) & 1 for c in face]
        for i in range(4):
            if s[i] == 1 and s[(i + 1) % 4] == 0:      # + -> - crossing: a segment starts here
                start = edge_id(face[i], face[(i + 1) % 4])
                if face_is_ambiguous(case, face) and connect_bits[f]:
                    j = (i + 1) % 4                       # around the negative corner
                    end = edge_id(face[j], face[(j + 1) % 4])
                else:
                    if face_is_ambiguous(case, face):
                        j = (i - 1) % 4                   # around this positive corner
                    else:
                        j = next(k for k in range(4) if s[k] == 0 and s[(k + 1) % 4] == 1)
                    end = edge_id(face[j], face[(j + 1) % 4])
                assert start not in nxt
                nxt[start] = end
    crossing = sorted(e for e, (a, b) in enumerate(EDGE) if ((case >> a) & 1) != ((case >> b) & 1))
    assert sorted(nxt.keys()) == crossing and sorted(nxt.values()) == crossing
    loops, seen = [], set()
    for e in crossing:
        if e in seen:
            continue
        loop = [e]
        seen.add(e)
        while nxt[loop[-1]] != e:
            loop.append(nxt[loop[-1]])
            seen.add(loop[-1])
        loops.append(loop)
    tris = []
    for loop in loops:
        assert len(loop) >= 3
        t = triangulate_loop(loop)
        if t is None:                       # needs the centre vertex (id 12)
            n = len(loop)
            t = [(CENTRE, loop[i], loop[(i + 1) % n]) for i in range(n)]
        tris.extend(t)
    return tris


CENTRE = 12
EDGE_FACES = None


def share_face(a, b):
    global EDGE_FACES
    if EDGE_FACES is None:
        EDGE_FACES = [{f for f, face in enumerate(FACES) if EDGE[e][0] in face and EDGE[e][1] in face}
                      for e in range(12)]
    return bool(EDGE_FACES[a] & EDGE_FACES[b])


def triangulate_loop(loop):
    """First triangulation (fixed DFS order) whose diagonals never lie in a cube face, or None."""
    n = len(loop)
    memo = {}

    def T(i, j):
        if j == i + 1:
            return []
        if (i, j) in memo:
            return memo[(i, j)]
        res = None
        if (i == 0 and j == n - 1) or not share_face(loop[i], loop[j]):
            for k in range(i + 1, j):
                a = T(i, k)
                if a is None:
                    continue
                b = T(k, j)
                if b is None:
                    continue
                res = [(loop[i], loop[k], loop[j])] + a + b
                break
        memo[(i, j)] = res
        return res

    return T(0, n - 1)


def build():
    amb_mask = []
    var_base = []
    entries = []       # list of triangle lists
    for case in range(256):
        mask = 0
        for f, face in enumerate(FACES):
            if face_is_ambiguous(case, face):
                mask |= 1 << f
        amb_mask.append(mask)
        var_base.append(len(entries))
        amb_faces = [f for f in range(6) if (mask >> f) & 1]
        for v in range(1 << len(amb_faces)):
            bits = [0] * 6
            for i, f in enumerate(amb_faces):
                bits[f] = (v >> i) & 1
            entries.append(triangulate(case, bits))
    return amb_mask, var_base, entries


def self_check(entries, amb_mask, var_base):
    P = np.array(CORNER, dtype=float)
    # single positive corner 0: one triangle whose normal points at the corner (towards increasing values)
    t = entries[var_base[1]]
    assert len(t) == 1
    mid = [0.5 * (P[EDGE[e][0]] + P[EDGE[e][1]]) for e in t[0]]
    nrm = np.cross(mid[1] - mid[0], mid[2] - mid[0])
    assert np.dot(nrm, P[0] - np.mean(mid, axis=0)) > 0
    # every directed boundary edge of a cell's patch set is used once; closedness is tested on volumes (tests/)
    for case in range(256):
        n_amb = bin(amb_mask[case]).count("1")
        for v in range(1 << n_amb):
            tris = entries[var_base[case] + v]
            used = sorted({e for tri in tris for e in tri if e != CENTRE})
            crossing = sorted(e for e, (a, b) in enumerate(EDGE) if ((case >> a) & 1) != ((case >> b) & 1))
            assert used == crossing, (case, v)
            assert len(tris) <= 12
    assert entries[var_base[0]] == [] and entries[var_base[255]] == []


def emit(path):
    amb_mask, var_base, entries = build()
    self_check(entries, amb_mask, var_base)
    tri_off, flat, ntri, rank = [], [], [], []
    for tris in entries:
        tri_off.append(len(flat) // 3)
        ntri.append(len(tris))
        order = []
        for tri in tris:
            for e in tri:
                flat.append(e)
                if e not in order:
                    order.append(e)
        rank.append([order.index(e) if e in order else 255 for e in range(13)])
    nent = len(entries)

    def arr(name, ctype, vals, per=16):
        lines = ["static const %s %s[%d] = {" % (ctype, name, len(vals))]
        for i in range(0, len(vals), per):
            lines.append("  " + ", ".join(str(v) for v in vals[i:i + per]) + ",")
        lines.append("};")
        return "\n".join(lines)

    edge_axis, edge_base = [], []
    for (a, b) in EDGE:
        d = [CORNER[b][i] - CORNER[a][i] for i in range(3)]
        axis = [i for i in range(3) if d[i] != 0][0]
        lo = a if d[axis] > 0 else b
        edge_axis.append(axis)
        edge_base.extend(CORNER[lo])
    out = []
    out.append("/* GENERATED by oracle/gen_mc_tables.py -- do not edit.  See that file for the derivation. */")
    out.append("#ifndef SURS_MC_TABLES_H\n#define SURS_MC_TABLES_H\n#include <stdint.h>")
    out.append("#define MC_NUM_ENTRIES %d" % nent)
    out.append("#define MC_NUM_TRI_IDX %d" % len(flat))
    out.append("/* corner c -> (d_axis0, d_axis1, d_axis2) */")
    out.append(arr("mc_corner_off", "uint8_t", [v for c in CORNER for v in c], 3))
    out.append("/* edge e -> its two corners (Lewiner/Bourke numbering) */")
    out.append(arr("mc_edge_corner", "uint8_t", [v for e in EDGE for v in e], 2))
    out.append("/* edge e -> array axis it runs along, and the offsets of its lower end */")
    out.append(arr("mc_edge_axis", "uint8_t", edge_axis, 12))
    out.append(arr("mc_edge_base", "uint8_t", edge_base, 3))
    out.append("/* face f -> 4 corners, counter-clockwise seen from outside */")
    out.append(arr("mc_face_corner", "uint8_t", [c for f in FACES for c in f], 4))
    out.append("/* case -> 6-bit mask of ambiguous faces; case -> first entry (entry = base + decider bits, compressed) */")
    out.append(arr("mc_amb_mask", "uint8_t", amb_mask))
    out.append(arr("mc_var_base", "uint16_t", var_base))
    out.append("/* entry -> #triangles, first triangle; triangles as edge-id triples */")
    out.append(arr("mc_ntri", "uint8_t", ntri))
    out.append(arr("mc_tri_off", "uint16_t", tri_off))
    out.append(arr("mc_tri_edges", "uint8_t", flat, 24))
    out.append("/* entry -> rank of vertex slot e (0..11 = cube edges, 12 = centre vertex) in first-use order of\n"
               "   the entry's triangle list (255 = unused) */")
    out.append("#define MC_CENTRE 12")
    out.append(arr("mc_edge_rank", "uint8_t", [v for r in rank for v in r], 13))
    out.append("#endif")
    with open(path, "w") as f:
        f.write("\n".join(out) + "\n")
    return nent, len(flat)


if __name__ == "__main__":
    here = os.path.dirname(os.path.abspath(__file__))
    dst = os.path.join(here, "..", "super-resolution-3d-human-shape-from-a-single-low-resolution-image_b200",
                       "csrc", "mc_tables.h")
    print(emit(os.path.normpath(dst)))
