"""One query of 2^24 random points through surs_query (for ncu captures of query_tc_kernel)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from surs_b200 import _capi, synthetic as syn
dev = torch.device("cuda:0")
case = syn.SyntheticCase(S=512, seed=0)
ctx = _capi.Context(dev)
t = lambda a: torch.from_numpy(a).to(dev)
ctx.set_weights([t(w) for w in case.mlp_lr[0]], [t(b) for b in case.mlp_lr[1]], [t(w) for w in case.mlp_hr[0]], [t(b) for b in case.mlp_hr[1]],
                syn.MLP_DIM_LR, syn.MLP_DIM_HR, syn.RES_LAYERS)
ctx.set_features(t(case.feat_lr), t(case.feat_hr))
n = 1 << int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 24
pts = torch.rand(3, n, device=dev) - 0.5
for _ in range(2):
    hr, lr = ctx.query(pts, case.calib, float(case.load_size // 2), float(case.z_size))
torch.cuda.synchronize()
print("ok", float(hr.mean()))
