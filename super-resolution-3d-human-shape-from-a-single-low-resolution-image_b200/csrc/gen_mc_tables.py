#!/usr/bin/env python
"""Generates the marching-cubes case tables (csrc/mc_tables.h) from first principles.

Why generated: the reference's marching cubes is scikit-image 0.17.2's
``marching_cubes_lewiner`` (lib/mesh_util.py:40,45), a third-party Cython module
that is neither vendored in the reference tree nor installed here, and Lewiner's
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
values, the result is watertight.

Interior (tunnel) test -- Lewiner's / Chernyaev's cases 4, 6, 7, 10, 12, 13.  The face
decisions fix the loops on the cube surface, and with them the same-sign REGIONS of the
surface (corners joined along same-sign edges and across ambiguous faces as decided).
Two regions Pa, Pc of sign s that border the same opposite-sign region Q are, for the
trilinear interpolant, either separated inside the cube (one disc per loop) or joined by
a tunnel through Q's chamber.  They are joined iff some axis-aligned section at height t
in (0,1) shows them connected: with At..Dt the values on the four cube edges parallel to
the axis (cyclic order; A on Pa's side, C on Pc's), At and Ct of sign s and
At*Ct - Bt*Dt > 0 (the asymptotic decider of the section).  q(t) = At*Ct - Bt*Dt is a
quadratic that is <= 0 where the section leaves the admissible interval, so only its
stationary point t* = -b / (2a) (a < 0) has to be examined -- the structure of Lewiner's
test_interior, but derived, and applied along all three axes.  The tables hold, per entry,
the list of such tests (the eight corners of the lines A..D, the sign s) and the entry to
switch to when one succeeds: the same cell with the two loops joined by a band of
triangles (zipper between the loops, no band edge lying in a cube face).

Known deviations from Lewiner's MC33 (the part of the parity that stays UNPINNED): the
triangulation of a loop / band (which diagonals, when the centre vertex is used) is ours,
and the interior test is the derivation above rather than his per-case reference edges.

Orientation: in output coordinates (axis0, axis1, axis2) taken as a right-handed
frame, (v1-v0)x(v2-v0) points towards increasing volume values (what skimage's
gradient_direction='descent' yields for an occupancy field; the reference then
swaps the winding when it writes the OBJ, lib/mesh_util.py:60).

Run:  python csrc/gen_mc_tables.py   (rewrites csrc/mc_tables.h; output is committed)
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


def trace_loops(case, connect_bits):
    """Closed loops of the iso-line on the cube surface: lists of edge ids, each segment u -> nxt[u] directed so that
    the positive side lies on its left seen from outside.  connect_bits[f] = 1 -> positive corners joined across
    ambiguous face f."""
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
    return loops


def surface_regions(case, connect_bits):
    """corner -> region id: same-sign corners joined along cube edges and across ambiguous faces as decided."""
    parent = list(range(8))

    def find(x):
        while parent[x] != x:
            parent[x] = parent[parent[x]]
            x = parent[x]
        return x

    for a, b in EDGE:
        if ((case >> a) & 1) == ((case >> b) & 1):
            parent[find(a)] = find(b)
    for f, face in enumerate(FACES):
        if face_is_ambiguous(case, face):
            pos = [c for c in face if (case >> c) & 1]
            neg = [c for c in face if not (case >> c) & 1]
            a, b = pos if connect_bits[f] else neg
            parent[find(a)] = find(b)
    return [find(c) for c in range(8)]


def loop_sides(case, loop, region):
    """(positive region, negative region) separated by a loop."""
    pos, neg = set(), set()
    for e in loop:
        a, b = EDGE[e]
        p, n = (a, b) if (case >> a) & 1 else (b, a)
        pos.add(region[p])
        neg.add(region[n])
    assert len(pos) == 1 and len(neg) == 1
    return pos.pop(), neg.pop()


def discs(loops):
    tris = []
    for loop in loops:
        assert len(loop) >= 3
        t = triangulate_loop(loop)
        if t is None:                       # needs the centre vertex (id 12)
            n = len(loop)
            t = [(CENTRE, loop[i], loop[(i + 1) % n]) for i in range(n)]
        tris.extend(t)
    return tris


def triangulate(case, connect_bits, tunnel=None):
    """Triangles (edge-id triples) of one cell.  tunnel = (i, j): loops i and j are joined by a band instead of being
    closed by a disc each."""
    loops = trace_loops(case, connect_bits)
    if tunnel is None:
        return discs(loops)
    i, j = tunnel
    band = triangulate_band(loops[i], loops[j])
    if band is None:
        return None
    return band + discs([l for k, l in enumerate(loops) if k not in (i, j)])


_MID = None


def _mid(e):
    global _MID
    if _MID is None:
        P = np.array(CORNER, dtype=float)
        _MID = [0.5 * (P[a] + P[b]) for a, b in EDGE]
    return _MID[e]


def tri_area(t):
    """Area of a triangle given by three vertex slots, every vertex at the middle of its cube edge."""
    p, q, r = (_mid(e) for e in t)
    return 0.5 * float(np.linalg.norm(np.cross(q - p, r - p)))


def in_one_face(t):
    """All three vertices on one cube face: the triangle would lie in that face."""
    global EDGE_FACES
    share_face(0, 1)
    return bool(EDGE_FACES[t[0]] & EDGE_FACES[t[1]] & EDGE_FACES[t[2]])


def _seg_hits_tri(p0, p1, tri):
    e1, e2 = tri[1] - tri[0], tri[2] - tri[0]
    dirv = p1 - p0
    h = np.cross(dirv, e2)
    a = float(np.dot(e1, h))
    if abs(a) < 1e-12:
        return False
    f = 1.0 / a
    sv = p0 - tri[0]
    u = f * float(np.dot(sv, h))
    q = np.cross(sv, e1)
    v = f * float(np.dot(dirv, q))
    t = f * float(np.dot(e2, q))
    m = 1e-9
    return u > m and v > m and u + v < 1 - m and m < t < 1 - m


def self_intersections(tris, pos=None):
    """Pairs of triangles (not sharing the tested edge's end points) that cut through each other."""
    pos = pos or {e: _mid(e) for t in tris for e in t}
    n = 0
    for a in range(len(tris)):
        for b in range(len(tris)):
            if a == b:
                continue
            ta, tb = tris[a], tris[b]
            T = np.array([pos[e] for e in tb])
            for q in range(3):
                e0, e1 = ta[q], ta[(q + 1) % 3]
                if e0 in tb or e1 in tb:
                    continue
                if _seg_hits_tri(pos[e0], pos[e1], T):
                    n += 1
    return n


def _zipper(la, lb):
    """Smallest-area zipper between two loops: every triangle is one loop segment (in loop direction) plus an apex on
    the other loop; walking forwards along la the band walks backwards along lb.  Area is measured with every vertex at
    the middle of its edge: it picks the untwisted, shortest band -- a twisted one is the same annulus topologically
    but its flat triangles cut through each other.  No triangle may lie in a cube face (all three vertices on it);
    single EDGES lying in a face are counted, not forbidden (see triangulate_band).
    Returns ((area, in-face cross edges), triangles)."""
    n1, n2 = len(la), len(lb)
    best = None
    for j0 in range(n2):
        memo = {}

        def go(i, j, ib, sb, ja, sa):
            # i segments of la and j segments of lb consumed; current cross edge (la[i % n1], lb[(j0 - j) % n2]).
            # ib = i at which lb's first segment was consumed (-1: none yet), sb = lb's segments were consumed at two
            # different i at least; ja / sa likewise.  A loop consumed around ONE apex would fold the band onto itself.
            if i == n1 and j == n2:
                return ((0.0, 0), []) if (sb and sa) else None
            key = (i, j, ib, sb, ja, sa)
            if key in memo:
                return memo[key]
            a, b = la[i % n1], lb[(j0 - j) % n2]
            res = None
            if i < n1:
                a2 = la[(i + 1) % n1]
                closing = (i + 1 == n1 and j == n2)
                rest = None if in_one_face((a, a2, b)) else go(i + 1, j, ib, sb, j if ja < 0 else ja, sa or (ja >= 0 and ja != j))
                if rest is not None:
                    cost = (rest[0][0] + tri_area((a, a2, b)), rest[0][1] + (0 if closing else int(share_face(a2, b))))
                    res = (cost, [(a, a2, b)] + rest[1])
            if j < n2:
                b2 = lb[(j0 - j - 1) % n2]
                closing = (i == n1 and j + 1 == n2)
                rest = None if in_one_face((b2, b, a)) else go(i, j + 1, i if ib < 0 else ib, sb or (ib >= 0 and ib != i), ja, sa)
                if rest is not None:
                    cost = (rest[0][0] + tri_area((b2, b, a)), rest[0][1] + (0 if closing else int(share_face(a, b2))))
                    if res is None or cost < res[0]:
                        res = (cost, [(b2, b, a)] + rest[1])
            memo[key] = res
            return res

        r = go(0, 0, -1, False, -1, False)
        if r is not None:
            cost = (r[0][0], r[0][1] + int(share_face(la[0], lb[j0])))
            if best is None or cost < best[0]:
                best = (cost, r[1])
    return best


def _reductions(loop):
    """Ways of cutting ears off a loop before the band is attached: (kept vertices in loop order, ear triangles).
    The chain between two kept vertices is closed by their chord (never in a cube face) and triangulated like a loop.
    Ordered by the number of kept vertices, largest first (the unreduced loop first)."""
    n = len(loop)
    out = []
    for mask in range((1 << n) - 1, 0, -1):
        keep = [i for i in range(n) if (mask >> i) & 1]
        if len(keep) < 3:
            continue
        tris, ok = [], True
        for q in range(len(keep)):
            i, j = keep[q], keep[(q + 1) % len(keep)]
            chain = [loop[(i + k) % n] for k in range(((j - i) % n) + 1)]
            if len(chain) == 2:
                continue
            if share_face(chain[0], chain[-1]):
                ok = False
                break
            t = triangulate_loop(chain)
            if t is None:
                ok = False
                break
            tris += t
        if ok:
            out.append(([loop[i] for i in keep], tris))
    out.sort(key=lambda r: -len(r[0]))
    return out


def triangulate_band(la, lb):
    """Triangles of the annulus between two loops (both directed as boundary of the surface, so every triangle keeps
    the loops' direction): optionally ears cut off the loops (chords never in a cube face; case 7.4.2's hexagon becomes a
    triangle, as in Lewiner's 9-triangle tiling), then the zipper between the (reduced) loops.  Chosen: no flat
    triangles cutting through each other, then the smallest area.  Unlike a loop's disc, a band cannot always avoid
    single edges that lie in a cube face -- the natural, untwisted band between a corner's triangle and the loop around
    the neighbouring corners runs along the faces they share; forcing it off the faces twists it through itself.  Such
    an edge is interior to THIS cell's patch (two triangles of this cell share it), so the surface stays a closed
    2-manifold; it touches the face along that edge.  Only two face-adjacent tunnel cells choosing the very same
    in-face edge would pinch, which needs two coincident rare events (`in_face_edges` counts them per table)."""
    best = None
    for ka, ea in _reductions(la):
        for kb, eb in _reductions(lb):
            z = _zipper(ka, kb)
            if z is None:
                continue
            tris = z[1] + ea + eb
            cost = (self_intersections(tris), z[0][0] + sum(tri_area(t) for t in ea + eb), z[0][1])
            if best is None or cost < best[0]:
                best = (cost, tris)
    triangulate_band.in_face_edges = getattr(triangulate_band, "in_face_edges", 0) + best[0][2]
    triangulate_band.self_intersecting = getattr(triangulate_band, "self_intersecting", 0) + int(best[0][0] > 0)
    return best[1]


def interior_test_passes(d, corners, s):
    """The run-time test (csrc/mc.cu interior_test, oracle/mc_oracle.c): d[..., c] = value - level at corner c
    (one cube or an array of cubes)."""
    d = np.asarray(d, dtype=np.float64)
    a0, a1, b0, b1, c0, c1, d0, d1 = (d[..., c] for c in corners)
    dA, dB, dC, dD = a1 - a0, b1 - b0, c1 - c0, d1 - d0
    qa = dA * dC - dB * dD
    qb = (a0 * dC + c0 * dA) - (b0 * dD + d0 * dB)
    with np.errstate(divide="ignore", invalid="ignore"):
        t = -qb / (2.0 * qa)
    At, Bt, Ct, Dt = a0 + dA * t, b0 + dB * t, c0 + dC * t, d0 + dD * t
    side = ((At > 0.0) & (Ct > 0.0)) if s else ~((At > 0.0) | (Ct > 0.0))
    return (qa < 0.0) & (t > 0.0) & (t < 1.0) & side & (At * Ct - Bt * Dt > 0.0)


def random_cubes(case, connect_bits, n, rng, batches=12):
    """Up to n random corner-value sets with the signs of `case` and the face decisions of `connect_bits` (rejection
    sampling, bounded: combinations of decisions that hardly ever occur return few or no cubes)."""
    sign = np.array([1.0 if (case >> c) & 1 else -1.0 for c in range(8)])
    out, have = [], 0
    for _ in range(batches):
        d = sign * rng.random((n, 8)) ** 2              # magnitudes skewed towards 0: thin features are the hard cases
        ok = np.ones(len(d), bool)
        for f, face in enumerate(FACES):
            if not face_is_ambiguous(case, face):
                continue
            p02, p13 = d[:, face[0]] * d[:, face[2]], d[:, face[1]] * d[:, face[3]]
            connect = np.where(d[:, face[0]] > 0, p02 > p13, p13 > p02)
            ok &= connect == bool(connect_bits[f])
        out.append(d[ok])
        have += int(ok.sum())
        if have >= n:
            break
    return np.concatenate(out)[:n]


def axis_lines(axis):
    """The four cube edges parallel to `axis` as corner pairs (lower, upper), in cyclic order around the axis."""
    cyc = FACES[2 * axis]
    up = lambda c: CORNER.index(tuple(v + (1 if a == axis else 0) for a, v in enumerate(CORNER[c])))
    assert all(CORNER[c][axis] == 0 for c in cyc)
    return [(c, up(c)) for c in cyc]


def interior_tests(case, connect_bits):
    """[(corners A0 A1 B0 B1 C0 C1 D0 D1, s, (loop i, loop j))]: see the module docstring."""
    loops = trace_loops(case, connect_bits)
    if len(loops) < 2:
        return []
    region = surface_regions(case, connect_bits)
    sides = [loop_sides(case, l, region) for l in loops]
    sign_of = {region[c]: (case >> c) & 1 for c in range(8)}
    tests = []
    for axis in range(3):
        lines = axis_lines(axis)
        for rot in (0, 1):
            A, B, C, D = lines[rot:] + lines[:rot]
            for s in (1, 0):
                def reg(line):
                    for c in line:
                        if ((case >> c) & 1) == s:
                            return region[c]
                    return None
                ra, rc = reg(A), reg(C)
                if ra is None or rc is None or ra == rc:
                    continue
                # the loops that separate ra / rc from a common opposite-sign region Q
                pair = None
                for i, (p, n) in enumerate(sides):
                    for j, (p2, n2) in enumerate(sides):
                        if i == j:
                            continue
                        mine_i, other_i = (p, n) if s else (n, p)
                        mine_j, other_j = (p2, n2) if s else (n2, p2)
                        if mine_i == ra and mine_j == rc and other_i == other_j:
                            pair = (min(i, j), max(i, j))
                if pair is None:
                    continue
                rec = (tuple(A + B + C + D), s, pair)
                if rec not in tests:
                    tests.append(rec)
    return tests


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
    meta = []          # (case, bits) per base entry
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
            meta.append((case, bits))
    # interior tests: per base entry a list of (8 corners, s, target entry); tunnel variants are appended after the
    # base entries (one per (entry, joined loop pair))
    n_base = len(entries)
    tests = [[] for _ in range(n_base)]
    tunnel_entry = {}
    rng = np.random.default_rng(33)
    dropped = 0
    for ent in range(n_base):
        case, bits = meta[ent]
        cand = interior_tests(case, bits)
        if not cand:
            continue
        # The candidates are a superset: for most (case, face decisions) the trilinear interpolant cannot form the tunnel
        # at all (MC33 needs an interior test for its sub-cases 4, 6.1, 7.4, 10.1, 12.1, 13.5 only).  A candidate that
        # never succeeds on 200 000 random cubes with these signs and face decisions is dropped.
        cubes = random_cubes(case, bits, 100000, rng)
        for corners, s, pair in cand:
            if not interior_test_passes(cubes, corners, s).any():
                dropped += 1
                continue
            key = (ent, pair)
            if key not in tunnel_entry:
                tunnel_entry[key] = len(entries)
                entries.append(triangulate(case, bits, tunnel=pair))
                meta.append((case, bits))
            tests[ent].append((corners, s, tunnel_entry[key]))
    build.dropped = dropped
    tests += [[] for _ in range(len(entries) - n_base)]
    return amb_mask, var_base, entries, tests, meta


MC_NUM_BASE = [0]


def self_check(entries, amb_mask, var_base, tests, meta):
    MC_NUM_BASE[0] = sum(1 << bin(m).count("1") for m in amb_mask)
    P = np.array(CORNER, dtype=float)
    # single positive corner 0: one triangle whose normal points at the corner (towards increasing values)
    t = entries[var_base[1]]
    assert len(t) == 1
    mid = [0.5 * (P[EDGE[e][0]] + P[EDGE[e][1]]) for e in t[0]]
    nrm = np.cross(mid[1] - mid[0], mid[2] - mid[0])
    assert np.dot(nrm, P[0] - np.mean(mid, axis=0)) > 0
    for ent, tris in enumerate(entries):
        case, bits = meta[ent]
        used = sorted({e for tri in tris for e in tri if e != CENTRE})
        crossing = sorted(e for e, (a, b) in enumerate(EDGE) if ((case >> a) & 1) != ((case >> b) & 1))
        assert used == crossing, (case, bits)
        assert len(tris) <= 16
        # every directed mesh edge once; the boundary of the patch is exactly the directed loop segments (so that the
        # neighbouring cell, which sees the same segments reversed, closes the surface); interior edges are matched
        loops = trace_loops(case, bits)
        seg = {(l[i], l[(i + 1) % len(l)]) for l in loops for i in range(len(l))}
        directed = [(tri[k], tri[(k + 1) % 3]) for tri in tris for k in range(3)]
        assert len(set(directed)) == len(directed), (case, bits)
        dset = set(directed)
        boundary = {d for d in dset if (d[1], d[0]) not in dset}
        assert boundary == seg, (case, bits, ent)
        # no interior mesh edge lies in a cube face
        if ent < MC_NUM_BASE[0]:
            for (u, v) in dset - seg:
                if CENTRE in (u, v):
                    continue
                assert not share_face(u, v) or (v, u) in seg or (u, v) in seg, (case, bits, ent, u, v)
    assert entries[var_base[0]] == [] and entries[var_base[255]] == []
    # Euler characteristic: discs only -> #loops; one tunnel -> #loops - 2
    n_base = sum(1 for t in tests if True)
    for ent, tris in enumerate(entries):
        case, bits = meta[ent]
        if not tris:
            continue
        verts = {e for tri in tris for e in tri}
        edges = {tuple(sorted((tri[k], tri[(k + 1) % 3]))) for tri in tris for k in range(3)}
        chi = len(verts) - len(edges) + len(tris)
        nloops = len(trace_loops(case, bits))
        is_tunnel = any(ent == tgt for tl in tests for (_, _, tgt) in tl)
        assert chi == (nloops - 2 if is_tunnel else nloops), (case, bits, ent, chi, nloops)


def emit(path):
    amb_mask, var_base, entries, tests, meta = build()
    self_check(entries, amb_mask, var_base, tests, meta)
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
    out.append("/* GENERATED by csrc/gen_mc_tables.py -- do not edit.  See that file for the derivation. */")
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
    ntest, test_off, test_corners, test_sign, test_target = [], [], [], [], []
    for tl in tests:
        ntest.append(len(tl))
        test_off.append(len(test_sign))
        for corners, sgn, tgt in tl:
            test_corners.extend(corners)
            test_sign.append(sgn)
            test_target.append(tgt)
    out.append("/* interior (tunnel) tests: entry -> #tests, first test; test -> corners A0 A1 B0 B1 C0 C1 D0 D1 of the four\n"
               "   parallel cube edges, sign s of the regions it can join (1: positive), entry to use when it succeeds */")
    out.append("#define MC_NUM_BASE_ENTRIES %d" % sum(1 << bin(m).count("1") for m in amb_mask))
    out.append("#define MC_NUM_TESTS %d" % len(test_sign))
    out.append(arr("mc_ntest", "uint8_t", ntest))
    out.append(arr("mc_test_off", "uint16_t", test_off))
    out.append(arr("mc_test_corners", "uint8_t", test_corners, 8))
    out.append(arr("mc_test_sign", "uint8_t", test_sign))
    out.append(arr("mc_test_target", "uint16_t", test_target))
    out.append("#endif")
    with open(path, "w") as f:
        f.write("\n".join(out) + "\n")
    return nent, len(flat), len(test_sign)


if __name__ == "__main__":
    here = os.path.dirname(os.path.abspath(__file__))
    print(emit(os.path.join(here, "mc_tables.h")))
