"""Marching cubes of the two 512^3 volumes of the bench workload, for ncu (kernels mc_*):
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:mc_ \
        --csv --log-file gpurun_out/mc.csv python scripts/mc_profile.py [res] [reps]
Prints the host-side wall time per volume as well (CUDA events, both volumes, count + emit)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from surs_b200 import _capi, synthetic as syn
from surs_b200.lib import sdf as bsdf

res = int(sys.argv[1]) if len(sys.argv) > 1 else 512
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device("cuda:0")
case = syn.SyntheticCase(S=512, seed=0)
ctx = _capi.Context(dev)
t = lambda a: torch.from_numpy(a).to(dev)
ctx.set_weights([t(w) for w in case.mlp_lr[0]], [t(b) for b in case.mlp_lr[1]], [t(w) for w in case.mlp_hr[0]], [t(b) for b in case.mlp_hr[1]],
                syn.MLP_DIM_LR, syn.MLP_DIM_HR, syn.RES_LAYERS)
ctx.set_features(t(case.feat_lr), t(case.feat_hr))
b_min, b_max = np.array([-0.5] * 3), np.array([0.5] * 3)
vols = ctx.eval_grid((res,) * 3, b_min, b_max, case.calib, float(case.load_size // 2), float(case.z_size), precision=_capi.PREC_FP16)
mat = bsdf.grid_matrix(res, b_min, b_max)[:3, :4]
for rep in range(reps):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    tot = 0
    for vol in vols:
        nv, nf, _ = ctx.mc_count(vol, 0.5)
        ctx.mc_emit_verts(nv, mat)
        ctx.mc_emit_faces(nf)
        tot += vol.numel() * 4 + nv * 52 + nf * 12
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    print("rep %d: both volumes %.3f ms, %.1f MB algorithmic -> %.0f GB/s" % (rep, ms, tot / 1e6, tot / ms / 1e6))
