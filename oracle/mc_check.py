"""TABLE-FREE checker of a marching-cubes mesh.  TEST INFRASTRUCTURE ONLY.

The CUDA kernels (csrc/mc.cu) and their CPU twin (oracle/mc_oracle.c) share the generated case tables
(csrc/mc_tables.h), so "GPU == twin" cannot see a wrong table entry.  This module checks a mesh against the VOLUME
alone, from first principles, without reading any table:

  check_mesh(vol, level, verts, faces)
    1. every vertex lies on a grid edge whose end points straddle the level, at the position of the interpolation
       formula (weights 1 / (FLT_EPSILON + |v - level|), the one skimage's Lewiner code uses), or strictly inside one
       cell (Lewiner's centre vertex); sign-changing grid edges <-> edge vertices is a bijection;
    2. every triangle lives in exactly ONE cell (its three vertices sit on edges of / inside the same cube) and does
       not lie in a cell face;
    3. every directed mesh edge is used once; its reverse is used once too unless the edge lies in the volume's
       border (closed, consistently oriented 2-manifold up to the border);
    4. per-face segment agreement: an edge lying in a cell face is shared by two triangles of the two cells on either
       side of the face (or, counted separately, by two triangles of one tunnel cell);
    5. orientation: (v1 - v0) x (v2 - v0) points towards increasing values of the trilinear interpolant.

  check_cell_topology(vol, level, verts, faces, cells=...)
    6. for the cells with any ambiguity, the TOPOLOGY chosen by the mesh (how many separate positive / negative
       chambers the cell's triangles cut the cube into -- the outcome of the face deciders AND of the interior /
       tunnel test) equals that of the trilinear interpolant, counted by brute force on a lattice inside the cell.
"""
import numpy as np

FLT_EPSILON = float(np.finfo(np.float32).eps)


def _edge_point(base, da, db):
    wa = 1.0 / (FLT_EPSILON + np.abs(da))
    wb = 1.0 / (FLT_EPSILON + np.abs(db))
    return (base * wa + (base + 1.0) * wb) / (wa + wb)


def _classify_vertices(d, verts):
    """-> (is_edge [V], base [V,3] int, axis [V] int) ; centre vertices have axis = -1 and base = their cell."""
    v = np.asarray(verts, np.float64)
    fl = np.floor(v)
    frac = v != fl
    nfrac = frac.sum(1)
    is_edge = nfrac <= 1
    axis = np.where(nfrac == 1, frac.argmax(1), -1)
    base = fl.astype(np.int64)
    # a vertex exactly on a node (interpolation weight saturated): attach it to the sign-changing edge it closes
    exact = np.nonzero(nfrac == 0)[0]
    for q in exact:
        b = base[q]
        found = False
        for a in range(3):
            for s in (0, -1):
                lo = b.copy()
                lo[a] += s
                hi = lo.copy()
                hi[a] += 1
                if lo[a] < 0 or hi[a] >= d.shape[a]:
                    continue
                if (d[tuple(lo)] > 0) != (d[tuple(hi)] > 0):
                    base[q], axis[q], found = lo, a, True
                    break
            if found:
                break
        assert found, "vertex %d sits on a node without a sign-changing edge" % q
    return is_edge, base, axis


def check_mesh(vol, level, verts, faces, pos_tol=1e-5, check_orientation=True):
    vol = np.asarray(vol, np.float32)
    d = vol.astype(np.float64) - float(level)
    inside = d > 0
    verts = np.asarray(verts)
    faces = np.asarray(faces, np.int64)
    R = np.array(vol.shape)
    V = len(verts)
    rep = {"verts": V, "faces": len(faces)}
    assert faces.min() >= 0 and faces.max() < V
    is_edge, base, axis = _classify_vertices(d, verts)
    ev = np.nonzero(is_edge)[0]
    # ---- 1. edge vertices <-> sign-changing grid edges, positions
    crossing = set()
    for a in range(3):
        n = R[a] - 1
        lo = np.take(inside, np.arange(n), axis=a)
        hi = np.take(inside, np.arange(1, n + 1), axis=a)
        idx = np.argwhere(lo != hi)
        crossing.update((int(i), int(j), int(k), a) for i, j, k in idx)
    got = [(int(base[q, 0]), int(base[q, 1]), int(base[q, 2]), int(axis[q])) for q in ev]
    assert len(set(got)) == len(got), "two vertices on one grid edge"
    assert set(got) == crossing, "edge vertices and sign-changing grid edges differ (%d vs %d)" % (len(got), len(crossing))
    b = base[ev]
    a = axis[ev]
    hi = b.copy()
    hi[np.arange(len(ev)), a] += 1
    da, db = d[b[:, 0], b[:, 1], b[:, 2]], d[hi[:, 0], hi[:, 1], hi[:, 2]]
    want = b.astype(np.float64)
    want[np.arange(len(ev)), a] = _edge_point(b[np.arange(len(ev)), a].astype(np.float64), da, db)
    err = np.abs(want - verts[ev].astype(np.float64)).max() if len(ev) else 0.0
    rep["max_position_error"] = float(err)
    assert err <= pos_tol, "vertex position off by %g" % err
    cv = np.nonzero(~is_edge)[0]
    rep["centre_vertices"] = len(cv)
    if len(cv):
        c = verts[cv].astype(np.float64)
        assert ((c > np.floor(c).min(1, keepdims=True) - 1) & (c >= 0) & (c <= R - 1)).all()
    # ---- 2. one cell per triangle
    flo, fhi = _face_cells(is_edge, base, axis, faces)
    assert (flo <= fhi).all(), "a triangle's vertices do not belong to one cell"
    in_face = (flo < fhi).any(1)                    # all three vertices in one grid plane
    # (a zero-area triangle may satisfy this legitimately; real in-face triangles are errors)
    v0, v1, v2 = (verts[faces[:, s]].astype(np.float64) for s in range(3))
    nrm = np.cross(v1 - v0, v2 - v0)
    area2 = np.linalg.norm(nrm, axis=1)
    assert not (in_face & (area2 > 1e-9)).any(), "a triangle lies in a cell face"
    cell = np.clip(fhi, 0, R - 2)
    rep["cells_with_triangles"] = len(np.unique(cell, axis=0))
    # ---- 3. directed edges
    e = np.concatenate([faces[:, [0, 1]], faces[:, [1, 2]], faces[:, [2, 0]]])
    ecell = np.concatenate([cell, cell, cell])
    key = e[:, 0] * V + e[:, 1]
    rkey = e[:, 1] * V + e[:, 0]
    order = np.argsort(key)
    ks = key[order]
    assert (ks[1:] != ks[:-1]).all(), "a directed edge is used twice (non-manifold or flipped triangle)"
    pos = np.searchsorted(ks, rkey)
    has_rev = (pos < len(ks)) & (ks[np.minimum(pos, len(ks) - 1)] == rkey)
    # unmatched edges must lie in the border of the volume
    vf = verts.astype(np.float64)
    un = np.nonzero(~has_rev)[0]
    if len(un):
        p, q = vf[e[un, 0]], vf[e[un, 1]]
        on_border = np.zeros(len(un), bool)
        for o in range(3):
            for plane in (0.0, float(R[o] - 1)):
                on_border |= (p[:, o] == plane) & (q[:, o] == plane)
        assert on_border.all(), "%d open edges away from the volume border" % int((~on_border).sum())
    rep["border_edges"] = int(len(un))
    # ---- 4. edges lying in a cell face: shared across the face (loop segments) or inside one tunnel cell
    p, q = vf[e[:, 0]], vf[e[:, 1]]
    planar = np.zeros(len(e), bool)
    for o in range(3):
        planar |= (p[:, o] == q[:, o]) & (p[:, o] == np.floor(p[:, o]))
    m = np.nonzero(planar & has_rev)[0]
    other = order[pos[m]]
    same_cell = (ecell[m] == ecell[other]).all(1)
    rep["in_face_edges_inside_one_cell"] = int(same_cell.sum() // 2)
    # ---- 5. orientation against the trilinear interpolant
    if check_orientation and len(faces):
        good = area2 > 1e-7
        c = (v0 + v1 + v2)[good] / 3.0
        n = nrm[good] / area2[good, None]
        eps = 2e-3
        f = _trilinear(d, np.clip(c + eps * n, 0, R - 1 - 1e-9)) - _trilinear(d, np.clip(c - eps * n, 0, R - 1 - 1e-9))
        wrong = int((f <= 0).sum())
        rep["orientation_checked"] = int(good.sum())
        rep["orientation_wrong"] = wrong
        rep["orientation_wrong_fraction"] = wrong / max(1, int(good.sum()))
    return rep


def _face_cells(is_edge, base, axis, faces):
    """Per axis the admissible cell indices of a vertex form an interval: [base - 1, base] across an edge that runs
    along another axis, [base, base] along the edge's own axis and for a centre vertex.  A triangle's cell is the
    intersection over its three vertices: (lower bounds, upper bounds), equal for a triangle that is not in a face."""
    lo = base.copy()
    for o in range(3):
        lo[is_edge & (axis != o), o] -= 1
    flo = np.maximum.reduce([lo[faces[:, s]] for s in range(3)])
    fhi = np.minimum.reduce([base[faces[:, s]] for s in range(3)])
    return flo, fhi


def _trilinear(d, p):
    i = np.minimum(np.floor(p).astype(np.int64), np.array(d.shape) - 2)
    t = p - i
    out = 0.0
    for c in range(8):
        o = np.array([(c >> 2) & 1, (c >> 1) & 1, c & 1])
        w = np.prod(np.where(o, t, 1.0 - t), axis=1)
        out = out + w * d[i[:, 0] + o[0], i[:, 1] + o[1], i[:, 2] + o[2]]
    return out


# ---------------------------------------------------------------------------------------------------------------------
def _components(blocked_x, blocked_y, blocked_z, n):
    """Connected components of an n^3 lattice whose axis-neighbour links are removed where blocked_* is True."""
    import scipy.sparse as sp
    from scipy.sparse.csgraph import connected_components
    idx = np.arange(n ** 3).reshape(n, n, n)
    rows, cols = [], []
    for blk, a, b in ((blocked_x, idx[:-1], idx[1:]), (blocked_y, idx[:, :-1], idx[:, 1:]), (blocked_z, idx[:, :, :-1], idx[:, :, 1:])):
        keep = ~blk
        rows.append(a[keep])
        cols.append(b[keep])
    rows, cols = np.concatenate(rows), np.concatenate(cols)
    g = sp.coo_matrix((np.ones(len(rows), np.int8), (rows, cols)), shape=(n ** 3, n ** 3))
    return connected_components(g, directed=False)[1].reshape(n, n, n)


def _segments_hit(p0, p1, tri):   # (kept for ad-hoc debugging of a cell's geometry)
    """Which segments p0->p1 [M,3] intersect triangle tri [3,3] (Moller-Trumbore, closed triangle with a margin)."""
    e1, e2 = tri[1] - tri[0], tri[2] - tri[0]
    dirv = p1 - p0
    h = np.cross(dirv, e2)
    a = h @ e1
    ok = np.abs(a) > 1e-14
    f = np.where(ok, 1.0 / np.where(ok, a, 1.0), 0.0)
    s = p0 - tri[0]
    u = f * np.einsum("ij,ij->i", s, h)
    qv = np.cross(s, e1)
    v = f * np.einsum("ij,ij->i", dirv, qv)
    t = f * (qv @ e2)
    m = 1e-9
    return ok & (u >= -m) & (v >= -m) & (u + v <= 1 + m) & (t >= -m) & (t <= 1 + m)


def trilinear_chambers(dcell, n=24):
    """(#positive, #negative) connected components of the trilinear interpolant of dcell [2,2,2] (value - level at the
    corners) inside the closed unit cube, counted on two lattices that include the cube's faces; None when the
    lattices disagree (a feature thinner than the lattice)."""
    def count(m):
        g = np.linspace(0.0, 1.0, m)
        x, y, z = np.meshgrid(g, g, g, indexing="ij")
        f = np.zeros_like(x)
        for c in range(8):
            o = ((c >> 2) & 1, (c >> 1) & 1, c & 1)
            f += (x if o[0] else 1 - x) * (y if o[1] else 1 - y) * (z if o[2] else 1 - z) * dcell[o]
        pos = f > 0
        lab = _components(pos[:-1] != pos[1:], pos[:, :-1] != pos[:, 1:], pos[:, :, :-1] != pos[:, :, 1:], m)
        return len(np.unique(lab[pos])), len(np.unique(lab[~pos]))
    t1, t2 = count(n), count(2 * n + 1)
    return t1 if t1 == t2 else None


def mesh_chambers(dcell, tri_ids, slot):
    """(#positive, #negative) chambers the cell's triangles cut the cube into, derived combinatorially from the mesh:
    tri_ids [T,3] vertex ids, slot[id] = (corner_a, corner_b) of the cube edge the vertex sits on (corners as (0/1,0/1,0/1)
    tuples) or None for a centre vertex.
      - surface regions: corners joined along same-sign cube edges, and across a face with four sign changes the two
        corners that the patch's boundary segments in that face do NOT cut off;
      - every connected component of the patch separates the cube in two: all regions on the positive side of its
        boundary segments become one chamber, all on the negative side another."""
    corners = [(a, b, c) for a in (0, 1) for b in (0, 1) for c in (0, 1)]
    parent = {c: c for c in corners}

    def find(x):
        while parent[x] != x:
            parent[x] = parent[parent[x]]
            x = parent[x]
        return x

    def union(x, y):
        parent[find(x)] = find(y)

    pos = {c: bool(dcell[c] > 0) for c in corners}
    for a, b in _CUBE_EDGES:
        if pos[a] == pos[b]:
            union(a, b)
    # boundary segments of the patch: edges used by exactly one triangle of this cell
    from collections import Counter, defaultdict
    und = Counter()
    for t in tri_ids:
        for q in range(3):
            und[frozenset((int(t[q]), int(t[(q + 1) % 3])))] += 1
    segs = [tuple(e) for e, c in und.items() if c == 1]
    # faces with four sign changes: the segment between two edge vertices cuts off the corner their cube edges share
    for axis in range(3):
        for side in (0, 1):
            fc = [c for c in corners if c[axis] == side]
            if sum(pos[c] for c in fc) != 2:
                continue
            p = [c for c in fc if pos[c]]
            if sum(abs(u - v) for u, v in zip(p[0], p[1])) != 2:
                continue                                          # adjacent, not diagonal: no ambiguity
            cut = set()
            for u, v in segs:
                su, sv = slot[u], slot[v]
                if su is None or sv is None:
                    continue
                if all(c[axis] == side for c in su + sv):
                    common = set(su) & set(sv)
                    if len(common) == 1:
                        cut.add(common.pop())
            if len(cut) != 2:
                return None                                       # segments in this face are not a pairing: malformed
            joined = [c for c in fc if c not in cut]
            union(joined[0], joined[1])
    # patch components
    tparent = list(range(len(tri_ids)))

    def tfind(x):
        while tparent[x] != x:
            tparent[x] = tparent[tparent[x]]
            x = tparent[x]
        return x

    owner = defaultdict(list)
    for ti, t in enumerate(tri_ids):
        for q in range(3):
            owner[frozenset((int(t[q]), int(t[(q + 1) % 3])))].append(ti)
    for lst in owner.values():
        for o in lst[1:]:
            tparent[tfind(o)] = tfind(lst[0])
    comp_pos, comp_neg = defaultdict(list), defaultdict(list)
    for e, lst in owner.items():
        if len(lst) != 1:
            continue
        comp = tfind(lst[0])
        for vid in e:
            sl = slot[vid]
            if sl is None:
                continue
            a, b = sl
            (comp_pos if pos[a] else comp_neg)[comp].append(a)
            (comp_pos if pos[b] else comp_neg)[comp].append(b)
    for groups in (comp_pos, comp_neg):
        for lst in groups.values():
            for c in lst[1:]:
                union(c, lst[0])
    return (len({find(c) for c in corners if pos[c]}), len({find(c) for c in corners if not pos[c]}))


def check_cell_topology(vol, level, verts, faces, max_cells=400, n=24, seed=0):
    """Check 6 of the module docstring on (a random subset of) the cells that need a decision: the positive or the
    negative corners are not connected along cube edges.  The mesh side is combinatorial (mesh_chambers), the truth is
    brute force on the trilinear interpolant (trilinear_chambers); flat triangles of one cell that cut through each
    other are counted separately (`self_intersecting_cells`)."""
    vol = np.asarray(vol, np.float32)
    d = vol.astype(np.float64) - float(level)
    R = np.array(vol.shape)
    verts = np.asarray(verts, np.float64)
    faces = np.asarray(faces, np.int64)
    is_edge, base, axis = _classify_vertices(d, verts)
    flo, fhi = _face_cells(is_edge, base, axis, faces)
    cell = np.clip(fhi, 0, R - 2)
    key = (cell[:, 0] * R[1] + cell[:, 1]) * R[2] + cell[:, 2]
    order = np.argsort(key, kind="stable")
    ukeys, start = np.unique(key[order], return_index=True)
    ends = dict(zip(ukeys, list(start[1:]) + [len(order)]))
    rng = np.random.default_rng(seed)
    rep = {"cells_needing_a_decision": 0, "cells_checked": 0, "undecidable": 0, "mismatch": 0, "tunnel_cells": 0,
           "self_intersecting_cells": 0, "mismatches": []}
    cand = []
    for u, s0 in zip(ukeys, start):
        i, j, k = int(u // (R[1] * R[2])), int(u // R[2] % R[1]), int(u % R[2])
        p = d[i:i + 2, j:j + 2, k:k + 2] > 0
        if _edge_connected(p) and _edge_connected(~p):
            continue
        cand.append((u, s0, i, j, k))
    rep["cells_needing_a_decision"] = len(cand)
    if len(cand) > max_cells:
        cand = [cand[q] for q in sorted(rng.choice(len(cand), max_cells, replace=False))]
    for u, s0, i, j, k in cand:
        fidx = order[s0:ends[u]]
        tri_ids = faces[fidx]
        origin = np.array([i, j, k])
        slot = {}
        for vid in np.unique(tri_ids):
            if not is_edge[vid]:
                slot[int(vid)] = None
                continue
            o = base[vid] - origin
            hi = o.copy()
            hi[axis[vid]] += 1
            slot[int(vid)] = (tuple(int(x) for x in o), tuple(int(x) for x in hi))
        dc = d[i:i + 2, j:j + 2, k:k + 2]
        truth = None if _near_tie(dc) else trilinear_chambers(dc, n=n)
        if truth is None:
            rep["undecidable"] += 1
            continue
        mesh = mesh_chambers(dc, tri_ids, slot)
        rep["cells_checked"] += 1
        # a patch component with two boundary loops = a tunnel
        if mesh is not None and len(tri_ids) and _has_annulus(tri_ids):
            rep["tunnel_cells"] += 1
        if mesh != truth:
            rep["mismatch"] += 1
            rep["mismatches"].append(((i, j, k), truth, mesh))
        tl = verts[tri_ids] - origin
        if _self_intersects(tl, tri_ids):
            rep["self_intersecting_cells"] += 1
    return rep


def _near_tie(dc, rel=0.02):
    """A face whose asymptotic decider is nearly tied (|ac - bd| small): the connection across it is thinner than any
    lattice, the brute-force count cannot be trusted there."""
    scale = float(np.abs(dc).max()) ** 2
    for axis in range(3):
        for side in (0, 1):
            f = np.take(dc, side, axis=axis)
            a, b, c, e = f[0, 0], f[0, 1], f[1, 1], f[1, 0]
            if (a > 0) == (c > 0) and (b > 0) == (e > 0) and (a > 0) != (b > 0) and abs(a * c - b * e) < rel * scale:
                return True
    return False


def _has_annulus(tri_ids):
    """Some connected component of the triangle set has Euler characteristic 0 (two boundary loops)."""
    from collections import defaultdict
    parent = list(range(len(tri_ids)))

    def find(x):
        while parent[x] != x:
            parent[x] = parent[parent[x]]
            x = parent[x]
        return x

    owner = defaultdict(list)
    for ti, t in enumerate(tri_ids):
        for q in range(3):
            owner[frozenset((int(t[q]), int(t[(q + 1) % 3])))].append(ti)
    for lst in owner.values():
        for o in lst[1:]:
            parent[find(o)] = find(lst[0])
    comps = defaultdict(lambda: [set(), set(), 0])
    for ti, t in enumerate(tri_ids):
        c = comps[find(ti)]
        c[0].update(int(x) for x in t)
        c[1].update(frozenset((int(t[q]), int(t[(q + 1) % 3]))) for q in range(3))
        c[2] += 1
    return any(len(v) - len(e) + f == 0 for v, e, f in comps.values())


def _self_intersects(tl, tri_ids):
    for a in range(len(tl)):
        for b in range(len(tl)):
            if a == b:
                continue
            for q in range(3):
                if tri_ids[a][q] in tri_ids[b] or tri_ids[a][(q + 1) % 3] in tri_ids[b]:
                    continue
                e1, e2 = tl[b][1] - tl[b][0], tl[b][2] - tl[b][0]
                p0, p1 = tl[a][q], tl[a][(q + 1) % 3]
                dirv = p1 - p0
                h = np.cross(dirv, e2)
                det = float(e1 @ h)
                if abs(det) < 1e-12:
                    continue
                f = 1.0 / det
                sv = p0 - tl[b][0]
                u = f * float(sv @ h)
                qv = np.cross(sv, e1)
                v = f * float(dirv @ qv)
                t = f * float(e2 @ qv)
                m = 1e-7
                if u > m and v > m and u + v < 1 - m and m < t < 1 - m:
                    return True
    return False


_CUBE_EDGES = [((0, 0, 0), (0, 0, 1)), ((0, 0, 0), (0, 1, 0)), ((0, 0, 0), (1, 0, 0)), ((0, 0, 1), (0, 1, 1)), ((0, 0, 1), (1, 0, 1)),
               ((0, 1, 0), (0, 1, 1)), ((0, 1, 0), (1, 1, 0)), ((1, 0, 0), (1, 0, 1)), ((1, 0, 0), (1, 1, 0)), ((0, 1, 1), (1, 1, 1)),
               ((1, 0, 1), (1, 1, 1)), ((1, 1, 0), (1, 1, 1))]


def _edge_connected(mask):
    """True when the corners selected by mask [2,2,2] form one set connected along cube edges (or none)."""
    sel = [tuple(int(v) for v in c) for c in np.argwhere(mask)]
    if len(sel) <= 1:
        return True
    seen, todo = {sel[0]}, [sel[0]]
    while todo:
        c = todo.pop()
        for a, b in _CUBE_EDGES:
            for x, y in ((a, b), (b, a)):
                if x == c and mask[y] and y not in seen:
                    seen.add(y)
                    todo.append(y)
    return len(seen) == len(sel)
