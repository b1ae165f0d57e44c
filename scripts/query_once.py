"""One raw query of N random points (BASELINE config 5) for ncu.  Usage: query_once.py [log2 N] [fp16|fp16r|fp16x3|fp32]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from surs_b200 import _capi, synthetic as syn

lg = int(sys.argv[1]) if len(sys.argv) > 1 else 22
prec = {"fp16": _capi.PREC_FP16, "fp16r": _capi.PREC_FP16R, "fp16x3": _capi.PREC_FP16X3, "fp32": _capi.PREC_FP32}[sys.argv[2] if len(sys.argv) > 2 else "fp16r"]
dev = torch.device("cuda:0")
case = syn.SyntheticCase(S=512, seed=0)
ctx = _capi.Context(dev)
t = lambda a: torch.from_numpy(a).to(dev)
ctx.set_weights([t(w) for w in case.mlp_lr[0]], [t(b) for b in case.mlp_lr[1]], [t(w) for w in case.mlp_hr[0]], [t(b) for b in case.mlp_hr[1]],
                syn.MLP_DIM_LR, syn.MLP_DIM_HR, syn.RES_LAYERS)
ctx.set_features(t(case.feat_lr), t(case.feat_hr))
pts = torch.rand((3, 1 << lg), device=dev) - 0.5
for rep in range(2):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    ctx.query(pts, case.calib, float(case.load_size // 2), float(case.z_size), precision=prec)
    b.record()
    torch.cuda.synchronize()
    print("2^%d points: %.2f ms = %.1f M queries/s" % (lg, a.elapsed_time(b), (1 << lg) / a.elapsed_time(b) / 1e3))
