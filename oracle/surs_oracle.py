"""CPU restatement of the SuRS reconstruction hot path (numpy, float64).

TEST INFRASTRUCTURE ONLY: imported by ``tests/``, ``__graft_entry__.smoke()`` and
the CPU-baseline legs of ``bench.py``; never by the product path.

Parity status: PINNED.  Every function here is checked against the reference's
own modules (imported unmodified from /root/reference in the build container by
``tests/golden/make_golden.py``); the resulting vectors live in ``tests/golden``
and ``tests/test_oracle_golden.py`` replays them on every run.  The marching
cubes twin lives in ``mc_oracle.c`` and is *unpinned* (see there).

Each function cites the reference file:line it restates (paths relative to the
reference root).
"""
from __future__ import annotations

import numpy as np

LEAKY_SLOPE = 0.01  # torch.nn.functional.leaky_relu default, lib/model/SurfaceClassifier.py:66


# ----------------------------------------------------------------------------
# L0 ops
# ----------------------------------------------------------------------------
def orthogonal(points, calib):
    """lib/geometry.py:15-31 (transforms=None): xyz = R.p + t.

    points [3,N] (cast to float32 first, as lib/mesh_util.py:24 does), calib
    [4,4] or [3,4].  Arithmetic in float64, result rounded to float32 (the
    reference computes a float32 baddbmm; the correctly rounded value differs
    from it by at most 1 ulp for a general calib and is exact for the
    diag(2,-2,2) calib used by gen_mesh, lib/train_util.py:63-66).
    """
    p = np.asarray(points, dtype=np.float32).astype(np.float64)
    c = np.asarray(calib, dtype=np.float32).astype(np.float64)
    xyz = c[:3, :3] @ p + c[:3, 3:4]
    return xyz.astype(np.float32)


def project(points, calib, perspective=False, uv_transform=None):
    """lib/geometry.py:15-31 `orthogonal` / :34-48 `perspective`, with the optional image-space affine `transforms`
    (:27-30,43-46: (u, v) <- scale (u, v) + shift with [scale | shift] = transforms[:2, :3]).  float32 result.
    NB the reference's own `transforms` branch cannot run (it indexes a 2-D matrix and hands it to baddbmm, which wants
    3-D operands); this restates its evident intent, PIFu's 2x3 affine."""
    xyz = orthogonal(points, calib).astype(np.float64)
    if perspective:
        xy = (xyz[:2].astype(np.float32) / xyz[2:3].astype(np.float32)).astype(np.float64)
    else:
        xy = xyz[:2]
    if uv_transform is not None:
        t = np.asarray(uv_transform, dtype=np.float32).astype(np.float64)[:2, :3]
        xy = (t[:, :2] @ xy + t[:, 2:3])
    return np.concatenate([xy, xyz[2:3]], axis=0).astype(np.float32)


def in_image_mask(xyz):
    """lib/model/SuRSNet.py:142 -- inclusive on both ends, float32 compare."""
    u, v = xyz[0], xyz[1]
    return (u >= -1.0) & (u <= 1.0) & (v >= -1.0) & (v <= 1.0)


def depth_feature(z, load_size, z_size):
    """lib/model/DepthNormalizer.py:18: z * (loadSize // 2) / z_size (float32 ops)."""
    z = np.asarray(z, dtype=np.float32)
    return (z * np.float32(load_size // 2)) / np.float32(z_size)


def index(feat, u, v):
    """lib/geometry.py:4-12: grid_sample(feat, uv, align_corners=True), bilinear, zero pad.

    feat [C,H,W]; u indexes W (last axis), v indexes H.  Returns [C,N] float64.
    """
    feat = np.asarray(feat)
    C, H, W = feat.shape
    u = np.asarray(u, dtype=np.float64)
    v = np.asarray(v, dtype=np.float64)
    ix = (u + 1.0) * 0.5 * (W - 1)
    iy = (v + 1.0) * 0.5 * (H - 1)
    x0 = np.floor(ix)
    y0 = np.floor(iy)
    x1 = x0 + 1
    y1 = y0 + 1
    out = np.zeros((C, u.shape[0]), dtype=np.float64)
    for xs, ys, w in (
        (x0, y0, (x1 - ix) * (y1 - iy)),
        (x1, y0, (ix - x0) * (y1 - iy)),
        (x0, y1, (x1 - ix) * (iy - y0)),
        (x1, y1, (ix - x0) * (iy - y0)),
    ):
        ok = (xs >= 0) & (xs < W) & (ys >= 0) & (ys < H)
        xi = np.clip(xs, 0, W - 1).astype(np.int64)
        yi = np.clip(ys, 0, H - 1).astype(np.int64)
        out += feat[:, yi, xi].astype(np.float64) * (w * ok)[None, :]
    return out


def leaky_relu(x):
    return np.where(x >= 0, x, LEAKY_SLOPE * x)


def surface_classifier(weights, biases, feature, res_layers=(2, 3, 4), no_residual=False):
    """lib/model/SurfaceClassifier.py:45-81 with num_views == 1, last_op = Sigmoid.

    weights[i] [Cout,Cin] (the Conv1d weight squeezed), feature [C0,N].
    The skip input is appended AFTER y: cat([y, feature]) (:63-64).
    """
    y = np.asarray(feature, dtype=np.float64)
    f = y
    n = len(weights)
    for i in range(n):
        w = np.asarray(weights[i], dtype=np.float64)
        b = np.asarray(biases[i], dtype=np.float64)
        x = np.concatenate([y, f], axis=0) if (not no_residual and i in res_layers) else y
        y = w @ x + b[:, None]
        if i != n - 1:
            y = leaky_relu(y)
    return 1.0 / (1.0 + np.exp(-y))


# ----------------------------------------------------------------------------
# L1: query_mr + query_sr + get_preds
# ----------------------------------------------------------------------------
def surface_classifier_views(weights, biases, features, res_layers=(2, 3, 4)):
    """lib/model/SurfaceClassifier.py:45-81 with num_views = V > 1 (one subject): features [V,C0,N]; after layer
    len(filters) // 2 = 2 (and its leaky ReLU) y and the skip input are averaged over the views (:70-76)."""
    feats = np.asarray(features, dtype=np.float64)
    V = feats.shape[0]
    ys = [feats[v] for v in range(V)]
    fs = [feats[v] for v in range(V)]
    n = len(weights)
    for i in range(n):
        w = np.asarray(weights[i], dtype=np.float64)
        b = np.asarray(biases[i], dtype=np.float64)
        out = []
        for y, f in zip(ys, fs):
            x = np.concatenate([y, f], axis=0) if i in res_layers else y
            y = w @ x + b[:, None]
            if i != n - 1:
                y = leaky_relu(y)
            out.append(y)
        ys = out
        if i == n // 2 and len(ys) > 1:
            ys = [np.mean(np.stack(ys), axis=0)]
            fs = [np.mean(np.stack(fs), axis=0)]
    return 1.0 / (1.0 + np.exp(-ys[0]))


def query_views(points, calibs, feats_lr, feats_hr, mlp_lr, mlp_hr, load_size=512, z_size=200.0, perspective=False, uv_transform=None):
    """lib/model/SuRSNet.py:131-187 with num_views = V views of one subject: points [V,3,N] (the same points repeated,
    lib/mesh_util.py:22), calibs [V,4,4], feats_lr [V,256,H,W], feats_hr [V,64,H,W] -> (pred_hr, pred_lr) [V,N]."""
    V = len(calibs)
    f321, masks = [], []
    for v in range(V):
        xyz = project(points[v], calibs[v], perspective, uv_transform)
        masks.append(in_image_mask(xyz).astype(np.float64))
        zf = depth_feature(xyz[2], load_size, z_size).astype(np.float64)
        u, w = xyz[0].astype(np.float64), xyz[1].astype(np.float64)
        f321.append(np.concatenate([index(feats_lr[v], u, w), index(feats_hr[v], u, w), zf[None, :]], axis=0))
    masks = np.stack(masks)
    pred_lr = masks * surface_classifier_views(mlp_lr[0], mlp_lr[1], np.stack(f321))[0][None, :]
    f322 = np.stack([np.concatenate([f321[v], pred_lr[v][None, :]], axis=0) for v in range(V)])
    pred_hr = masks * surface_classifier_views(mlp_hr[0], mlp_hr[1], f322)[0][None, :]
    return pred_hr, pred_lr


def query(points, calib, feat_lr, feat_hr, mlp_lr, mlp_hr, load_size=512, z_size=200.0,
          res_layers=(2, 3, 4), return_features=False, perspective=False, uv_transform=None):
    """lib/model/SuRSNet.py:131-187 + lib/model/BaseSuRSNet.py:80-85 (num_views == 1).

    points [3,N]; feat_lr [256,Hl,Wl], feat_hr [64,Hh,Wh]; mlp_* = (weights, biases).
    Returns (pred_hr, pred_lr) float64 [N] -- HR first, as get_preds does.
    """
    xyz = project(points, calib, perspective, uv_transform)
    mask = in_image_mask(xyz).astype(np.float64)
    zf = depth_feature(xyz[2], load_size, z_size).astype(np.float64)
    u = xyz[0].astype(np.float64)
    v = xyz[1].astype(np.float64)
    f321 = np.concatenate([index(feat_lr, u, v), index(feat_hr, u, v), zf[None, :]], axis=0)
    pred_lr = mask * surface_classifier(mlp_lr[0], mlp_lr[1], f321, res_layers)[0]
    f322 = np.concatenate([f321, pred_lr[None, :]], axis=0)
    pred_hr = mask * surface_classifier(mlp_hr[0], mlp_hr[1], f322, res_layers)[0]
    if return_features:
        return pred_hr, pred_lr, f322
    return pred_hr, pred_lr


def query_chunked(points, *args, chunk=65536, **kw):
    n = points.shape[1]
    hr = np.empty(n)
    lr = np.empty(n)
    for s in range(0, n, chunk):
        hr[s:s + chunk], lr[s:s + chunk] = query(points[:, s:s + chunk], *args, **kw)
    return hr, lr


# ----------------------------------------------------------------------------
# L2: lib/sdf.py
# ----------------------------------------------------------------------------
def create_grid(resX, resY, resZ, b_min=np.array([-1, -1, -1]), b_max=np.array([1, 1, 1]), transform=None):
    """lib/sdf.py:4-29.  No half-voxel offset; the last node is b_max - len/res."""
    b_min = np.asarray(b_min, dtype=np.float64)
    b_max = np.asarray(b_max, dtype=np.float64)
    mat = np.eye(4)
    length = b_max - b_min
    mat[0, 0] = length[0] / resX
    mat[1, 1] = length[1] / resY
    mat[2, 2] = length[2] / resZ
    mat[0:3, 3] = b_min
    ii, jj, kk = np.meshgrid(np.arange(resX), np.arange(resY), np.arange(resZ), indexing="ij")
    idx = np.stack([ii, jj, kk]).reshape(3, -1)
    coords = np.matmul(mat[:3, :3], idx) + mat[:3, 3:4]
    if transform is not None:
        coords = np.matmul(transform[:3, :3], coords) + transform[:3, 3:4]
        mat = np.matmul(transform, mat)
    return coords.reshape(3, resX, resY, resZ), mat


def batch_eval(points, eval_func, num_samples=512 * 512 * 512):
    """lib/sdf.py:32-45.  eval_func(points[3,n]) -> (hr, lr), broadcast-assigned into [n] slices."""
    num_pts = points.shape[1]
    sdf_lr = np.zeros(num_pts)
    sdf_hr = np.zeros(num_pts)
    for s in range(0, num_pts, num_samples):
        hr, lr = eval_func(points[:, s:s + num_samples])
        sdf_hr[s:s + num_samples] = np.asarray(hr).reshape(-1)
        sdf_lr[s:s + num_samples] = np.asarray(lr).reshape(-1)
    return sdf_hr, sdf_lr


def eval_grid(coords, eval_func, num_samples=512 * 512 * 512):
    """lib/sdf.py:48-52."""
    resolution = coords.shape[1:4]
    hr, lr = batch_eval(coords.reshape(3, -1), eval_func, num_samples)
    return hr.reshape(resolution), lr.reshape(resolution)


def octree_cell_pass(sdf_hr, sdf_lr, dirty, reso, threshold):
    """One level of the interpolation loop, lib/sdf.py:81-117, vectorised.

    Valid because within one level the sequential loop is order independent: the
    only grid node a block fill overwrites is the cell's own origin, and every
    other cell reading that node has a lexicographically smaller origin
    (SURVEY.md §3.3 property 2).  So: snapshot centre-dirty flags and all corner
    values, then apply the fills.  Cells exist only for origins in
    range(0, R - reso, reso) (the last cell row is never visited, :81-83).
    dirty is shared by HR and LR (:99-101, :115-117).  Arrays are modified in place.
    Returns (#cells filled hr, #cells filled lr).
    """
    R0, R1, R2 = sdf_hr.shape
    xs = np.arange(0, R0 - reso, reso)
    ys = np.arange(0, R1 - reso, reso)
    zs = np.arange(0, R2 - reso, reso)
    if len(xs) == 0 or len(ys) == 0 or len(zs) == 0:
        return 0, 0
    X, Y, Z = np.meshgrid(xs, ys, zs, indexing="ij")
    h = reso // 2
    active = dirty[X + h, Y + h, Z + h].copy()

    def corner_range(vol):
        vmin = np.full(X.shape, np.inf)
        vmax = np.full(X.shape, -np.inf)
        for dx in (0, reso):
            for dy in (0, reso):
                for dz in (0, reso):
                    c = vol[X + dx, Y + dy, Z + dz]
                    vmin = np.minimum(vmin, c)
                    vmax = np.maximum(vmax, c)
        return vmin, vmax

    lo_h, hi_h = corner_range(sdf_hr)
    lo_l, hi_l = corner_range(sdf_lr)
    fill_h = active & ((hi_h - lo_h) < threshold)
    fill_l = active & ((hi_l - lo_l) < threshold)
    for fill, vol, lo, hi in ((fill_h, sdf_hr, lo_h, hi_h), (fill_l, sdf_lr, lo_l, hi_l)):
        cx, cy, cz = X[fill], Y[fill], Z[fill]
        val = (hi[fill] + lo[fill]) / 2
        for dx in range(reso):          # blocks of distinct cells are disjoint
            for dy in range(reso):
                for dz in range(reso):
                    vol[cx + dx, cy + dy, cz + dz] = val
                    dirty[cx + dx, cy + dy, cz + dz] = False
    return int(fill_h.sum()), int(fill_l.sum())


def eval_grid_octree(threshold, coords, eval_func, init_resolution=64, num_samples=512 * 512 * 512,
                     stats=None):
    """lib/sdf.py:55-120 (``opt`` is only read for ``opt.threshold``).

    R < init_resolution gives reso == 0 and returns zeros, as the reference does.
    """
    resolution = coords.shape[1:4]
    sdf_lr = np.zeros(resolution)
    sdf_hr = np.zeros(resolution)
    dirty = np.ones(resolution, dtype=bool)
    grid_mask = np.zeros(resolution, dtype=bool)
    reso = resolution[0] // init_resolution
    while reso > 0:
        grid_mask[0:resolution[0]:reso, 0:resolution[1]:reso, 0:resolution[2]:reso] = True
        test_mask = np.logical_and(grid_mask, dirty)
        points = coords[:, test_mask]
        if stats is not None:
            stats.append((reso, int(test_mask.sum())))
        sdf_hr[test_mask], sdf_lr[test_mask] = batch_eval(points, eval_func, num_samples=num_samples)
        dirty[test_mask] = False
        if reso <= 1:
            break
        octree_cell_pass(sdf_hr, sdf_lr, dirty, reso, threshold)
        reso //= 2
    return sdf_hr, sdf_lr


def eval_grid_octree_sequential(threshold, coords, eval_func, init_resolution=64,
                                num_samples=512 * 512 * 512):
    """Literal loop-for-loop restatement of lib/sdf.py:55-120; pure Python, small cases only.
    Used by the tests to check the vectorised cell pass above without the reference mount."""
    resolution = coords.shape[1:4]
    sdf_lr = np.zeros(resolution)
    sdf_hr = np.zeros(resolution)
    dirty = np.ones(resolution, dtype=bool)
    grid_mask = np.zeros(resolution, dtype=bool)
    reso = resolution[0] // init_resolution
    while reso > 0:
        grid_mask[0:resolution[0]:reso, 0:resolution[1]:reso, 0:resolution[2]:reso] = True
        test_mask = np.logical_and(grid_mask, dirty)
        points = coords[:, test_mask]
        sdf_hr[test_mask], sdf_lr[test_mask] = batch_eval(points, eval_func, num_samples=num_samples)
        dirty[test_mask] = False
        if reso <= 1:
            break
        for x in range(0, resolution[0] - reso, reso):
            for y in range(0, resolution[1] - reso, reso):
                for z in range(0, resolution[2] - reso, reso):
                    if not dirty[x + reso // 2, y + reso // 2, z + reso // 2]:
                        continue
                    for vol in (sdf_hr, sdf_lr):
                        v = vol[x:x + reso + 1:reso, y:y + reso + 1:reso, z:z + reso + 1:reso]
                        v_min, v_max = v.min(), v.max()
                        if (v_max - v_min) < threshold:
                            vol[x:x + reso, y:y + reso, z:z + reso] = (v_max + v_min) / 2
                            dirty[x:x + reso, y:y + reso, z:z + reso] = False
        reso //= 2
    return sdf_hr, sdf_lr


# ----------------------------------------------------------------------------
# L3: lib/mesh_util.py
# ----------------------------------------------------------------------------
def verts_to_world(mat, verts):
    """lib/mesh_util.py:42-43: verts = (mat[:3,:3] @ verts.T + mat[:3,3:4]).T, float64."""
    return (np.matmul(mat[:3, :3], np.asarray(verts).T) + mat[:3, 3:4]).T


def obj_text(verts, faces):
    """lib/mesh_util.py:53-61: 'v %.4f %.4f %.4f' and 1-based 'f a c b' (winding swapped)."""
    out = []
    for v in verts:
        out.append('v %.4f %.4f %.4f\n' % (v[0], v[1], v[2]))
    for f in faces:
        out.append('f %d %d %d\n' % (f[0] + 1, f[2] + 1, f[1] + 1))
    return ''.join(out)
