"""Drop-in for the reference's ``lib/sdf.py`` (same names, arguments and results).

``create_grid`` / ``batch_eval`` / ``eval_grid`` / ``eval_grid_octree`` keep the generic
contract (any Python ``eval_func(points[3,n]) -> (hr, lr)``); their grid bookkeeping is host
numpy exactly like the reference, except that the octree's interpolation loop
(lib/sdf.py:81-117, a pure-Python triple loop in the reference) runs vectorised.  When the
network is the accelerated one, ``lib.mesh_util.reconstruction`` bypasses this module and
keeps the grid, the octree and both volumes on the device (``surs_eval_grid*``).
"""
import numpy as np


def create_grid(resX, resY, resZ, b_min=np.array([-1, -1, -1]), b_max=np.array([1, 1, 1]), transform=None):
    """reference lib/sdf.py:4-29.  Returns (coords float64 [3,resX,resY,resZ], 4x4 matrix)."""
    b_min = np.asarray(b_min, dtype=np.float64)
    b_max = np.asarray(b_max, dtype=np.float64)
    coords_matrix = np.eye(4)
    length = b_max - b_min
    coords_matrix[0, 0] = length[0] / resX
    coords_matrix[1, 1] = length[1] / resY
    coords_matrix[2, 2] = length[2] / resZ
    coords_matrix[0:3, 3] = b_min
    coords = np.mgrid[:resX, :resY, :resZ].reshape(3, -1)
    coords = np.matmul(coords_matrix[:3, :3], coords) + coords_matrix[:3, 3:4]
    if transform is not None:
        coords = np.matmul(transform[:3, :3], coords) + transform[:3, 3:4]
        coords_matrix = np.matmul(transform, coords_matrix)
    return coords.reshape(3, resX, resY, resZ), coords_matrix


def grid_matrix(resolution, b_min, b_max, transform=None):
    """Only the 4x4 index->world matrix of create_grid (no 3 x res^3 float64 coordinate array)."""
    res = (resolution,) * 3 if np.isscalar(resolution) else tuple(resolution)
    b_min = np.asarray(b_min, dtype=np.float64).reshape(3)
    b_max = np.asarray(b_max, dtype=np.float64).reshape(3)
    m = np.eye(4)
    for a in range(3):
        m[a, a] = (b_max[a] - b_min[a]) / res[a]
    m[0:3, 3] = b_min
    return m if transform is None else np.matmul(transform, m)


def batch_eval(points, eval_func, num_samples=512 * 512 * 512):
    """reference lib/sdf.py:32-45."""
    num_pts = points.shape[1]
    sdf_lr = np.zeros(num_pts)
    sdf_hr = np.zeros(num_pts)
    for s in range(0, num_pts, num_samples):
        hr, lr = eval_func(points[:, s:s + num_samples])
        sdf_hr[s:s + num_samples] = np.asarray(hr).reshape(-1)
        sdf_lr[s:s + num_samples] = np.asarray(lr).reshape(-1)
    return sdf_hr, sdf_lr


def eval_grid(coords, eval_func, num_samples=512 * 512 * 512):
    """reference lib/sdf.py:48-52."""
    resolution = coords.shape[1:4]
    sdf_hr, sdf_lr = batch_eval(coords.reshape([3, -1]), eval_func, num_samples=num_samples)
    return sdf_hr.reshape(resolution), sdf_lr.reshape(resolution)


def _cell_pass(sdf_hr, sdf_lr, dirty, reso, threshold):
    """One level of reference lib/sdf.py:81-117.  Decisions are taken from a snapshot of the
    corner values (the sequential loop never reads a value it has already overwritten: the only
    on-grid node a fill touches is the cell's own origin, read only by earlier cells)."""
    R0, R1, R2 = sdf_hr.shape
    xs, ys, zs = (np.arange(0, R - reso, reso) for R in (R0, R1, R2))
    if len(xs) == 0 or len(ys) == 0 or len(zs) == 0:
        return
    X, Y, Z = np.meshgrid(xs, ys, zs, indexing="ij")
    h = reso // 2
    active = dirty[X + h, Y + h, Z + h].copy()
    fills = []
    for vol in (sdf_hr, sdf_lr):
        lo = np.full(X.shape, np.inf)
        hi = np.full(X.shape, -np.inf)
        for dx in (0, reso):
            for dy in (0, reso):
                for dz in (0, reso):
                    c = vol[X + dx, Y + dy, Z + dz]
                    lo = np.minimum(lo, c)
                    hi = np.maximum(hi, c)
        fills.append((active & ((hi - lo) < threshold), (hi + lo) / 2))
    for vol, (fill, mid) in zip((sdf_hr, sdf_lr), fills):
        cx, cy, cz, val = X[fill], Y[fill], Z[fill], mid[fill]
        for dx in range(reso):
            for dy in range(reso):
                for dz in range(reso):
                    vol[cx + dx, cy + dy, cz + dz] = val
                    dirty[cx + dx, cy + dy, cz + dz] = False


def eval_grid_octree(opt, coords, eval_func, init_resolution=64, num_samples=512 * 512 * 512):
    """reference lib/sdf.py:55-120: same evaluated set and same volumes (including the zero
    holes left where the *other* volume was locally uniform)."""
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
        _cell_pass(sdf_hr, sdf_lr, dirty, reso, opt.threshold)
        reso //= 2
    return sdf_hr.reshape(resolution), sdf_lr.reshape(resolution)
