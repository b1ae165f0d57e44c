"""Runs the BASELINE.json configs that fit one GPU and prints one JSON line per config
(documentation numbers for profiles/; bench.py stays the driver-facing entry point).

  C1  dense 128^3 (the reference's CPU-runnable case)       S = 512 input
  C2  octree 256^3                                           S = 512
  C3  dense 512^3                                            S = 512   (= bench.py's workload)
  C4  octree 512^3 (one image per GPU = replicas)            S = 512
  C5  raw query sweep, 2^20 .. 2^26 random points            S = 512
Each: median of 3 timed runs after one warm-up, CUDA events, through the public reconstruction API
(device-side time of the whole call incl. marching cubes; host mesh copies excluded).
"""
import json
import os
import sys
import time
import types

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from surs_b200 import _capi, synthetic as syn
from surs_b200.lib import sdf as bsdf

FLOP = 4564998


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2], out


def main():
    S = int(os.environ.get("SURS_S", "512"))
    dev = torch.device("cuda:0")
    case = syn.SyntheticCase(S=S, seed=0)
    ctx = _capi.Context(dev)
    t = lambda a: torch.from_numpy(a).to(dev)
    ctx.set_weights([t(w) for w in case.mlp_lr[0]], [t(b) for b in case.mlp_lr[1]], [t(w) for w in case.mlp_hr[0]], [t(b) for b in case.mlp_hr[1]],
                    syn.MLP_DIM_LR, syn.MLP_DIM_HR, syn.RES_LAYERS)
    ctx.set_features(t(case.feat_lr), t(case.feat_hr))
    zn, zd = float(case.load_size // 2), float(case.z_size)
    bmin, bmax = np.array([-0.5] * 3), np.array([0.5] * 3)

    prec = {"fp16": _capi.PREC_FP16, "fp16x3": _capi.PREC_FP16X3, "fp16r": _capi.PREC_FP16R, "fp32": _capi.PREC_FP32}[os.environ.get("SURS_PRECISION", "fp16")]

    def recon(res, octree):
        mat = bsdf.grid_matrix(res, bmin, bmax)[:3, :4]
        if octree:
            hr, lr, n_eval = ctx.eval_grid_octree((res,) * 3, bmin, bmax, case.calib, zn, zd, 0.05, precision=prec)
            vols = (ctx.cast_f64_f32(hr), ctx.cast_f64_f32(lr))
        else:
            vols = ctx.eval_grid((res,) * 3, bmin, bmax, case.calib, zn, zd, precision=prec)
            n_eval = res ** 3
        meshes = []
        for v in vols:
            nv, nf, na = ctx.mc_count(v, 0.5)
            ctx.mc_emit_verts(nv, mat)
            ctx.mc_emit_faces(nf)
            meshes.append((nv, nf, na))
        return n_eval, meshes

    for name, res, octree in (("C1 dense 128^3", 128, False), ("C2 octree 256^3", 256, True), ("C3 dense 512^3", 512, False),
                              ("C4 octree 512^3", 512, True)):
        ms, (n_eval, meshes) = timed(lambda: recon(res, octree))
        print(json.dumps({"config": name, "precision": os.environ.get("SURS_PRECISION", "fp16"), "input_side": S, "ms_per_mesh": ms, "grid_nodes": res ** 3, "network_evaluations": n_eval,
                          "evaluated_fraction": n_eval / res ** 3, "grid_nodes_per_s": res ** 3 / ms * 1e3,
                          "evaluations_per_s": n_eval / ms * 1e3, "verts_faces_ambiguous_hr_lr": meshes}))
    if prec != _capi.PREC_FP16:
        return
    for lg in (20, 22, 24, 26):
        n = 1 << lg
        pts = torch.rand(3, n, device=dev, generator=torch.Generator(device=dev).manual_seed(lg)) - 0.5
        ms, _ = timed(lambda: ctx.query(pts, case.calib, zn, zd))
        print(json.dumps({"config": "C5 query sweep", "points": n, "ms": ms, "queries_per_s": n / ms * 1e3,
                          "algorithmic_tflops": n * FLOP / ms * 1e-9}))
        ms32, _ = (timed(lambda: ctx.query(pts[:, :1 << 20], case.calib, zn, zd, precision=_capi.PREC_FP32)) if lg == 20 else (None, None))
        if ms32:
            print(json.dumps({"config": "C5 query sweep (fp32 exact mode)", "points": 1 << 20, "ms": ms32, "queries_per_s": (1 << 20) / ms32 * 1e3}))


if __name__ == "__main__":
    main()
