import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch, torch.distributed as dist
from surs_b200 import _capi, parallel, synthetic as syn
from surs_b200.lib import sdf as bsdf
rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
case = syn.SyntheticCase(S=512, seed=0)
ctx = _capi.Context(dev)
t = lambda a: torch.from_numpy(a).to(dev)
ctx.set_weights([t(w) for w in case.mlp_lr[0]], [t(b) for b in case.mlp_lr[1]], [t(w) for w in case.mlp_hr[0]], [t(b) for b in case.mlp_hr[1]], syn.MLP_DIM_LR, syn.MLP_DIM_HR, syn.RES_LAYERS)
ctx.set_features(t(case.feat_lr), t(case.feat_hr))
res = 512
b_min, b_max = np.array([-0.5] * 3), np.array([0.5] * 3)
mat = bsdf.grid_matrix(res, b_min, b_max)
for i in range(4):
    parallel.reconstruct_slab(ctx, (res,) * 3, b_min, b_max, case.calib, 512.0, 200.0, mat[:3, :4])
dist.barrier(); torch.cuda.synchronize()
os.environ["SURS_TIMING"] = "1"
for i in range(2):
    t0 = time.perf_counter()
    parallel.reconstruct_slab(ctx, (res,) * 3, b_min, b_max, case.calib, 512.0, 200.0, mat[:3, :4])
    torch.cuda.synchronize()
    if rank == 0: print("total %.3f ms" % ((time.perf_counter() - t0) * 1e3), flush=True)
dist.destroy_process_group()
