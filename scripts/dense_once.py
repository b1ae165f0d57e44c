"""One dense grid evaluation at R^3, S = 512 (for ncu captures).  argv: resolution (default 256), precision (fp16 | fp16x3)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from surs_b200 import _capi, synthetic as syn
dev = torch.device("cuda:0")
res = int(sys.argv[1]) if len(sys.argv) > 1 else 256
prec = {"fp16": _capi.PREC_FP16, "fp16x3": _capi.PREC_FP16X3}[sys.argv[2] if len(sys.argv) > 2 else "fp16"]
case = syn.SyntheticCase(S=512, seed=0)
ctx = _capi.Context(dev)
t = lambda a: torch.from_numpy(a).to(dev)
ctx.set_weights([t(w) for w in case.mlp_lr[0]], [t(b) for b in case.mlp_lr[1]], [t(w) for w in case.mlp_hr[0]], [t(b) for b in case.mlp_hr[1]],
                syn.MLP_DIM_LR, syn.MLP_DIM_HR, syn.RES_LAYERS)
ctx.set_features(t(case.feat_lr), t(case.feat_hr))
hr, lr = ctx.eval_grid((res,) * 3, [-0.5] * 3, [0.5] * 3, case.calib, float(case.load_size // 2), float(case.z_size), precision=prec)
torch.cuda.synchronize()
print("ok", float(hr.mean()))
