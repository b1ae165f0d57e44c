"""Marching cubes, octree bookkeeping, refinement select and repack on small inputs -- the kernels that synchronise with
__syncthreads / warp primitives only -- for `compute-sanitizer --tool racecheck` (the tcgen05 kernels are excluded with
--kernel-regex: racecheck does not model tcgen05.commit -> mbarrier dependencies)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch

import helpers
from surs_b200 import _capi

ctx = _capi.Context("cuda:0")
dev = ctx.device
rng = np.random.default_rng(0)
for vol in (helpers.sphere_volume(32, 10.0), rng.random((12, 9, 16)).astype(np.float32), rng.random((9, 10, 27)).astype(np.float32)):
    out = ctx.marching_cubes(torch.from_numpy(vol).to(dev), 0.5, np.eye(4)[:3])
    print("mc", vol.shape, out[0].shape[0], out[2].shape[0])
v64 = torch.from_numpy(rng.random((16, 16, 32))).to(dev)
print("mc f64", ctx.mc_count(v64, 0.5), ctx.mc_value_range())
# octree building blocks on an analytic field
R = 32
g = np.stack(np.meshgrid(*[np.linspace(-0.5, 0.5, R, endpoint=False)] * 3, indexing="ij")).reshape(3, -1)
hr, lr = helpers.analytic_eval_func(g)
dirty = torch.ones((R, R, R), dtype=torch.uint8, device=dev)
idx = torch.empty(R ** 3, dtype=torch.int64, device=dev)
sdf_hr = torch.zeros((R, R, R), dtype=torch.float64, device=dev)
sdf_lr = torch.zeros_like(sdf_hr)
full_hr, full_lr = torch.from_numpy(hr.reshape(-1).astype(np.float64)).to(dev), torch.from_numpy(lr.reshape(-1).astype(np.float64)).to(dev)
reso = 4
while reso > 0:
    n = ctx.octree_select((R, R, R), reso, dirty, idx)
    sel = idx[:n]
    sdf_hr.view(-1)[sel] = full_hr[sel]
    sdf_lr.view(-1)[sel] = full_lr[sel]
    if reso <= 1:
        break
    ctx.octree_cells((R, R, R), reso, 0.05, sdf_hr, sdf_lr, dirty)
    reso //= 2
print("octree", int((sdf_hr != 0).sum()))
