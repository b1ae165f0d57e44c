"""One octree reconstruction (eval_grid_octree + marching cubes of both volumes) at R^3, S = 512 -- for the ncu
launch list of configs C2 / C4.  argv: resolution (default 512), precision (fp16 | fp16x3)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from surs_b200 import _capi, synthetic as syn
from surs_b200.lib import sdf as bsdf
dev = torch.device("cuda:0")
res = int(sys.argv[1]) if len(sys.argv) > 1 else 512
prec = {"fp16": _capi.PREC_FP16, "fp16x3": _capi.PREC_FP16X3}[sys.argv[2] if len(sys.argv) > 2 else "fp16"]
case = syn.SyntheticCase(S=512, seed=0)
ctx = _capi.Context(dev)
t = lambda a: torch.from_numpy(a).to(dev)
ctx.set_weights([t(w) for w in case.mlp_lr[0]], [t(b) for b in case.mlp_lr[1]], [t(w) for w in case.mlp_hr[0]], [t(b) for b in case.mlp_hr[1]],
                syn.MLP_DIM_LR, syn.MLP_DIM_HR, syn.RES_LAYERS)
ctx.set_features(t(case.feat_lr), t(case.feat_hr))
bmin, bmax = np.array([-0.5] * 3), np.array([0.5] * 3)
mat = bsdf.grid_matrix(res, bmin, bmax)[:3, :4]
reps = int(os.environ.get("SURS_REPS", "1"))
for _ in range(reps):
    hr, lr, n_eval = ctx.eval_grid_octree((res,) * 3, bmin, bmax, case.calib, float(case.load_size // 2), float(case.z_size), 0.05, precision=prec)
    for v in (ctx.cast_f64_f32(hr), ctx.cast_f64_f32(lr)):
        nv, nf, na = ctx.mc_count(v, 0.5)
        ctx.mc_emit_verts(nv, mat)
        ctx.mc_emit_faces(nf)
torch.cuda.synchronize()
print("ok", n_eval, nv, nf)
