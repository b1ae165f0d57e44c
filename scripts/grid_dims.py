"""eval_grid of an arbitrary [R0, R1, R2] grid (debug helper): grid_dims.py R0 R1 R2 [precision]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from surs_b200 import _capi, synthetic as syn

res = tuple(int(v) for v in sys.argv[1:4])
prec = {"fp16": _capi.PREC_FP16, "fp16r": _capi.PREC_FP16R, "fp16x3": _capi.PREC_FP16X3, "fp32": _capi.PREC_FP32}[sys.argv[4] if len(sys.argv) > 4 else "fp16"]
dev = torch.device("cuda:0")
case = syn.SyntheticCase(S=32, seed=0)
ctx = _capi.Context(dev)
t = lambda a: torch.from_numpy(a).to(dev)
ctx.set_weights([t(w) for w in case.mlp_lr[0]], [t(b) for b in case.mlp_lr[1]], [t(w) for w in case.mlp_hr[0]], [t(b) for b in case.mlp_hr[1]],
                syn.MLP_DIM_LR, syn.MLP_DIM_HR, syn.RES_LAYERS)
ctx.set_features(t(case.feat_lr), t(case.feat_hr))
vols = ctx.eval_grid(res, np.array([-0.5] * 3), np.array([0.5] * 3), case.calib, float(case.load_size // 2), float(case.z_size), precision=prec)
ref = ctx.eval_grid(res, np.array([-0.5] * 3), np.array([0.5] * 3), case.calib, float(case.load_size // 2), float(case.z_size), precision=_capi.PREC_FP32)
torch.cuda.synchronize()
print("grid", res, "ok; max |d| vs fp32", max((a - b).abs().max().item() for a, b in zip(vols, ref)))
