"""One dense grid evaluation of the bench workload (S = 512 synthetic input), for ncu:
    ncu --set full --clock-control none --import-source on -k regex:query_col_kernel -c 1 -o gpurun_out/col512 \
        python scripts/grid_once.py 512 fp16
Usage: grid_once.py [res] [fp16|fp16r|fp16x3|fp32] [planes]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from surs_b200 import _capi, synthetic as syn

res = int(sys.argv[1]) if len(sys.argv) > 1 else 512
prec = {"fp16": _capi.PREC_FP16, "fp16r": _capi.PREC_FP16R, "fp16x3": _capi.PREC_FP16X3, "fp32": _capi.PREC_FP32}[sys.argv[2] if len(sys.argv) > 2 else "fp16"]
planes = int(sys.argv[3]) if len(sys.argv) > 3 else res
dev = torch.device("cuda:0")
case = syn.SyntheticCase(S=512, seed=0)
ctx = _capi.Context(dev)
t = lambda a: torch.from_numpy(a).to(dev)
ctx.set_weights([t(w) for w in case.mlp_lr[0]], [t(b) for b in case.mlp_lr[1]], [t(w) for w in case.mlp_hr[0]], [t(b) for b in case.mlp_hr[1]],
                syn.MLP_DIM_LR, syn.MLP_DIM_HR, syn.RES_LAYERS)
ctx.set_features(t(case.feat_lr), t(case.feat_hr))
reps = int(os.environ.get("GRID_REPS", "1"))          # > 1: warm repeats, min / median reported (A/B runs: scripts/ab_grid.sh)
times = []
for _ in range(reps):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    vols = ctx.eval_grid((res,) * 3, np.array([-0.5] * 3), np.array([0.5] * 3), case.calib, float(case.load_size // 2), float(case.z_size), precision=prec,
                         plane_lo=0, plane_hi=planes)
    b.record()
    torch.cuda.synchronize()
    times.append(a.elapsed_time(b))
import hashlib

sha = hashlib.sha256(b"".join(v.cpu().numpy().tobytes() for v in vols)).hexdigest()[:16]
name = sys.argv[2] if len(sys.argv) > 2 else "fp16"
if reps == 1:
    print("eval_grid %d^3 (%d planes) precision %s: %.2f ms  sha %s" % (res, planes, name, times[0], sha))
else:
    print("eval_grid %d^3 (%d planes) precision %s: min %.2f median %.2f ms over %d warm calls (first %.2f)  sha %s" %
          (res, planes, name, min(times[1:]), float(np.median(times[1:])), reps - 1, times[0], sha))
