"""Parity report for the tensor-core (fp16) paths against the fp32 mode and the CPU oracle:
max / mean |d occupancy| pre-threshold, number of 0.5-classification flips, and the number of
near-threshold voxels (north_star: "bit-exact except in a reported count of near-threshold voxels").
Writes one JSON document (profiles/r1_parity_report.json when run with an output path)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from oracle import surs_oracle as O
from surs_b200 import _capi, synthetic as syn


def stats(a, b, band):
    d = (a - b).abs()
    flips = (a > 0.5) != (b > 0.5)
    near = (b - 0.5).abs() < band
    return {"max_abs_diff": float(d.max()), "mean_abs_diff": float(d.mean()), "p999_abs_diff": float(torch.quantile(d.flatten()[:: max(1, d.numel() // 4000000)], 0.999)),
            "n": int(d.numel()), "classification_flips": int(flips.sum()), "flips_outside_band": int((flips & ~near).sum()),
            "near_threshold_voxels(|occ-0.5|<%g)" % band: int(near.sum())}


def main():
    S = int(os.environ.get("SURS_S", "512"))
    res = int(os.environ.get("SURS_RES", "256"))
    dev = torch.device("cuda:0")
    case = syn.SyntheticCase(S=S, seed=0)
    ctx = _capi.Context(dev)
    t = lambda a: torch.from_numpy(a).to(dev)
    ctx.set_weights([t(w) for w in case.mlp_lr[0]], [t(b) for b in case.mlp_lr[1]], [t(w) for w in case.mlp_hr[0]], [t(b) for b in case.mlp_hr[1]],
                    syn.MLP_DIM_LR, syn.MLP_DIM_HR, syn.RES_LAYERS)
    ctx.set_features(t(case.feat_lr), t(case.feat_hr))
    zn, zd = float(case.load_size // 2), float(case.z_size)
    bmin, bmax = [-0.5] * 3, [0.5] * 3
    band = 1e-2
    out = {"input_side": S, "resolution": res, "stated_tolerance_fp16": {"max": 1e-2, "mean": 3e-4}, "band": band}
    col = ctx.eval_grid((res,) * 3, bmin, bmax, case.calib, zn, zd, precision=_capi.PREC_FP16)
    ref = ctx.eval_grid((res,) * 3, bmin, bmax, case.calib, zn, zd, precision=_capi.PREC_FP32)
    out["dense_grid_column_kernel_vs_fp32_mode"] = {"hr": stats(col[0], ref[0], band), "lr": stats(col[1], ref[1], band)}
    x3 = ctx.eval_grid((res,) * 3, bmin, bmax, case.calib, zn, zd, precision=_capi.PREC_FP16X3)
    out["dense_grid_split_operand_x3_vs_fp32_mode(band 1e-4)"] = {"hr": stats(x3[0], ref[0], 1e-4), "lr": stats(x3[1], ref[1], 1e-4)}
    del x3
    out["occupancy_histogram_hr(10 bins)"] = torch.histc(ref[0], bins=10, min=0, max=1).tolist()
    out["occupancy_histogram_lr(10 bins)"] = torch.histc(ref[1], bins=10, min=0, max=1).tolist()
    pts = torch.rand(3, 1 << 22, device=dev, generator=torch.Generator(device=dev).manual_seed(7)) * 1.1 - 0.55
    a = ctx.query(pts, case.calib, zn, zd, precision=_capi.PREC_FP16)
    b = ctx.query(pts, case.calib, zn, zd, precision=_capi.PREC_FP32)
    out["random_points_generic_kernel_vs_fp32_mode"] = {"hr": stats(a[0], b[0], band), "lr": stats(a[1], b[1], band)}
    n_or = int(os.environ.get("SURS_ORACLE_POINTS", "60000"))
    p_small = pts[:, :n_or].cpu().numpy()
    ohr, olr = O.query_chunked(p_small, case.calib, case.feat_lr, case.feat_hr, case.mlp_lr, case.mlp_hr, load_size=case.load_size, chunk=20000)
    for name, (x, y) in (("fp32_mode", b), ("fp16_generic", a)):
        out["%s_vs_cpu_oracle" % name] = {"hr": stats(x[:n_or].cpu().double(), torch.from_numpy(ohr), band),
                                          "lr": stats(y[:n_or].cpu().double(), torch.from_numpy(olr), band)}
    txt = json.dumps(out, indent=1)
    print(txt)
    if len(sys.argv) > 1:
        with open(sys.argv[1], "w") as f:
            f.write(txt + "\n")


if __name__ == "__main__":
    main()
