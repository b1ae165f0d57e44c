#!/usr/bin/env python
"""Image encoder on the B200: eager fp32 (TF32 off), eager with PyTorch's default TF32 convolutions
(how the reference itself runs on a GPU), and the "fast" configuration (channels_last + bf16 autocast + CUDA graph) --
time per image, feature-map error against fp32, and what that error does to the occupancy.
    python scripts/encoder_bench.py [S]     -> one JSON line
"""
import json
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from surs_b200 import _capi, synthetic as syn  # noqa: E402
from surs_b200.lib.model import SuRSNet  # noqa: E402


def main():
    S = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    dev = torch.device("cuda:0")
    opt = types.SimpleNamespace(num_views=1, no_residual=False, mlp_dim_lr=[321, 1024, 512, 256, 128, 1], mlp_dim_hr=[322, 1024, 512, 256, 128, 1],
                                mlp_res_layers_lr=[2, 3, 4], mlp_res_layers_hr=[2, 3, 4], loadSize=2 * S, z_size=200.0, threshold=0.05,
                                num_stack_lr=3, num_stack_hr=1, hg_depth=2, hg_dim=256, norm="group", n_block=[2, 2, 2], rgb_range=255, scale=2, residual=True)
    torch.manual_seed(0)
    net = SuRSNet(opt, encoder_mode="eager").to(dev).eval()
    # give the encoder a realistic dynamic range: the reference init (N(0, 0.02)) makes every feature ~1e-6
    for m in net.modules():
        if isinstance(m, torch.nn.Conv2d) and m.weight.shape[1] > 3:
            torch.nn.init.kaiming_normal_(m.weight, a=0.2)
    case = syn.SyntheticCase(S=S, seed=0)
    for mlp, wb in ((net.mlp_lr, case.mlp_lr), (net.mlp_hr, case.mlp_hr)):
        for conv, w, b in zip(mlp.layers(), wb[0], wb[1]):
            conv.weight.data = torch.from_numpy(w)[:, :, None].to(dev)
            conv.bias.data = torch.from_numpy(b).to(dev)
    g = torch.Generator().manual_seed(1991)
    img = (torch.rand(1, 3, S, S, generator=g) * 2 - 1).to(dev)
    fast = SuRSNet(opt, encoder_mode="fast").to(dev).eval()
    fast.load_state_dict(net.state_dict())
    bf16 = SuRSNet(opt, encoder_mode="bf16").to(dev).eval()
    bf16.load_state_dict(net.state_dict())

    def run(n, reps=5):
        def once():
            _, flr, fhr = n.super_res(img)
            n.filter_hr(fhr)
            n.filter_lr(flr)
        with torch.no_grad():
            for _ in range(3):
                once()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(reps):
                once()
            b.record()
            torch.cuda.synchronize()
        return a.elapsed_time(b) / reps, n.im_feat_list_lr[-1].clone(), n.im_feat_list_hr[0].clone()

    out = {"S": S}
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    ms32, lr32, hr32 = run(net)
    torch.backends.cudnn.allow_tf32 = True
    ms_tf, lr_tf, hr_tf = run(net)
    ms_f, lr_f, hr_f = run(fast)
    ms_b, lr_b, hr_b = run(bf16)
    rel = lambda a, b: float((a - b).norm() / b.norm())
    out["ms"] = {"eager_fp32": ms32, "eager_tf32_default": ms_tf, "fast_tf32_cuda_graph": ms_f, "bf16_channels_last_cuda_graph": ms_b}
    out["feature_rel_l2_vs_fp32"] = {"tf32": [rel(lr_tf, lr32), rel(hr_tf, hr32)], "fast": [rel(lr_f, lr32), rel(hr_f, hr32)], "bf16": [rel(lr_b, lr32), rel(hr_b, hr32)]}
    # effect on the occupancy: the same 2^20 random points through the fp32 kernel with each pair of maps
    pts = torch.from_numpy(syn.random_points(1 << 20, seed=5)).to(dev)
    ctx = net.surs_context()
    ctx.set_weights([c.weight for c in net.mlp_lr.layers()], [c.bias for c in net.mlp_lr.layers()], [c.weight for c in net.mlp_hr.layers()],
                    [c.bias for c in net.mlp_hr.layers()], opt.mlp_dim_lr, opt.mlp_dim_hr, opt.mlp_res_layers_lr)
    occ = {}
    for name, (flr, fhr) in (("fp32", (lr32, hr32)), ("tf32", (lr_tf, hr_tf)), ("fast", (lr_f, hr_f)), ("bf16", (lr_b, hr_b))):
        ctx.set_features(flr, fhr)
        occ[name] = ctx.query(pts, case.calib, float(S), 200.0, precision=_capi.PREC_FP16X3)
    out["occupancy_max_abs_vs_fp32_features"] = {k: [float((occ[k][0] - occ["fp32"][0]).abs().max()), float((occ[k][1] - occ["fp32"][1]).abs().max())]
                                                 for k in ("tf32", "fast", "bf16")}
    out["occupancy_mean_abs_vs_fp32_features"] = {k: [float((occ[k][0] - occ["fp32"][0]).abs().mean()), float((occ[k][1] - occ["fp32"][1]).abs().mean())]
                                                  for k in ("tf32", "fast", "bf16")}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
