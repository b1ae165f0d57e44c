"""Marching-cubes timing on a synthetic 512^3 volume (a noisy blob).  Usage: mc_time.py [res]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from surs_b200 import _capi

res = int(sys.argv[1]) if len(sys.argv) > 1 else 512
dev = torch.device("cuda:0")
ctx = _capi.Context(dev)
ax = torch.linspace(-1, 1, res, device=dev)
x, y, z = torch.meshgrid(ax, ax, ax, indexing="ij")
vol = torch.sigmoid(8 * (0.6 - torch.sqrt(x * x + 1.5 * y * y + z * z) + 0.05 * torch.sin(20 * x) * torch.cos(17 * y + 3 * z))).contiguous()
del x, y, z
for rep in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    nv, nf, amb = ctx.mc_count(vol, 0.5)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    out = ctx.marching_cubes(vol, 0.5)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print("res %d: count %.3f ms, count+emit %.3f ms; %d verts %d faces %d ambiguous" % (res, (t1 - t0) * 1e3, (t2 - t1) * 1e3, nv, nf, amb))
