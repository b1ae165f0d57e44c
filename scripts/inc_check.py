"""Dense-grid kernels compared on one synthetic case: incremental layer 1 (query_inc.cu, default) vs
the GEMM kernel (SURS_COL_GEMM=1, query_col.cu) vs the fp32 mode.  Usage: inc_check.py [S] [res]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from surs_b200 import _capi, synthetic as syn

S = int(sys.argv[1]) if len(sys.argv) > 1 else 64
res = int(sys.argv[2]) if len(sys.argv) > 2 else 128
dev = torch.device("cuda:0")
case = syn.SyntheticCase(S=S, seed=0)
ctx = _capi.Context(dev)
t = lambda a: torch.from_numpy(a).to(dev)
ctx.set_weights([t(w) for w in case.mlp_lr[0]], [t(b) for b in case.mlp_lr[1]], [t(w) for w in case.mlp_hr[0]], [t(b) for b in case.mlp_hr[1]],
                syn.MLP_DIM_LR, syn.MLP_DIM_HR, syn.RES_LAYERS)
ctx.set_features(t(case.feat_lr), t(case.feat_hr))
zn, zd = float(case.load_size // 2), float(case.z_size)
bmin, bmax = [-0.5] * 3, [0.5] * 3
shape = (res,) * 3 if len(sys.argv) < 4 else (res, res, int(sys.argv[3]))
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    a = ctx.eval_grid(shape, bmin, bmax, case.calib, zn, zd, precision=_capi.PREC_FP16)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    print("fp16 grid %s: %.2f ms" % (shape, (t1 - t0) * 1e3))
if os.environ.get("SURS_SKIP_REF"):
    sys.exit(0)
b = ctx.eval_grid(shape, bmin, bmax, case.calib, zn, zd, precision=_capi.PREC_FP32)
for name, x, y in (("hr", a[0], b[0]), ("lr", a[1], b[1])):
    d = (x - y).abs()
    print(name, "max|d| %.3g mean|d| %.3g  flips %d of %d" % (d.max().item(), d.mean().item(), ((x > 0.5) != (y > 0.5)).sum().item(), d.numel()))
a2 = ctx.eval_grid(shape, bmin, bmax, case.calib, zn, zd, precision=_capi.PREC_FP16)
print("deterministic:", torch.equal(a[0], a2[0]) and torch.equal(a[1], a2[1]))
