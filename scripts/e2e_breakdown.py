"""Times the pieces of one end-to-end 512^3 reconstruction (host buffers in, host arrays out)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from surs_b200 import _capi, synthetic as syn
from surs_b200.lib import sdf as bsdf

res = int(sys.argv[1]) if len(sys.argv) > 1 else 512
S = int(sys.argv[2]) if len(sys.argv) > 2 else 512
dev = torch.device("cuda:0")
case = syn.SyntheticCase(S=S, seed=0)
ctx = _capi.Context(dev)
t = lambda a: torch.from_numpy(a).to(dev)
ctx.set_weights([t(w) for w in case.mlp_lr[0]], [t(b) for b in case.mlp_lr[1]], [t(w) for w in case.mlp_hr[0]], [t(b) for b in case.mlp_hr[1]],
                syn.MLP_DIM_LR, syn.MLP_DIM_HR, syn.RES_LAYERS)
f_lr = torch.from_numpy(case.feat_lr).pin_memory(); f_hr = torch.from_numpy(case.feat_hr).pin_memory()
zn, zd = float(case.load_size // 2), float(case.z_size)
bmin, bmax = np.array([-0.5] * 3), np.array([0.5] * 3)
mat = bsdf.grid_matrix(res, bmin, bmax)

def tick(label, t0):
    torch.cuda.synchronize(); t1 = time.perf_counter(); print("%-28s %8.2f ms" % (label, (t1 - t0) * 1e3)); return t1

for rep in range(2):
    print("--- rep", rep)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    a, b = f_lr.to(dev, non_blocking=True), f_hr.to(dev, non_blocking=True); t0 = tick("H2D features", t0)
    ctx.set_features(a, b); t0 = tick("set_features (repack)", t0)
    hr, lr = ctx.eval_grid((res,) * 3, bmin, bmax, case.calib, zn, zd); t0 = tick("eval_grid", t0)
    for name, vol in (("hr", hr), ("lr", lr)):
        mn, mx = torch.aminmax(vol); float(mn); t0 = tick("aminmax " + name, t0)
        nv, nf, na = ctx.mc_count(vol, 0.5); t0 = tick("mc_count " + name, t0)
        verts, world, normals, values = ctx.mc_emit_verts(nv, mat[:3, :4]); t0 = tick("mc_emit_verts " + name, t0)
        faces = ctx.mc_emit_faces(nf); t0 = tick("mc_emit_faces " + name, t0)
        outs = [x.cpu().numpy() for x in (world, faces, normals, values)]; t0 = tick("D2H pageable %d MB" % (sum(o.nbytes for o in outs) >> 20), t0)
        pin = [torch.empty(x.shape, dtype=x.dtype, pin_memory=True) for x in (world, faces, normals, values)]; t0 = tick("alloc pinned", t0)
        for p, x in zip(pin, (world, faces, normals, values)): p.copy_(x, non_blocking=True)
        t0 = tick("D2H pinned", t0)
