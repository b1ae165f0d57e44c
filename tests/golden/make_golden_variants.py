#!/usr/bin/env python
"""Golden vectors for the query variants of SURVEY §8(f)-4, from the UNMODIFIED reference (run in the build container
only; output tests/golden/variants_golden.npz is committed):

  perspective   lib.model.SuRSNet(opt, projection_mode='perspective') -> query_mr + query_sr + get_preds
  multi-view    opt.num_views = 2, two views of one subject: features [2,C,H,W], calibs [2,4,4], points [2,3,N]
                (lib/model/SurfaceClassifier.py:70-76 mean over views after layer 2)

The reference's image-space `transforms` branch cannot run at all (lib/geometry.py:27-30 hands 2-D matrices to
baddbmm), so there is nothing to record for it; the oracle restates its intent and the CUDA kernels are checked
against the oracle.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import ref_import as R  # noqa: E402
from surs_b200 import synthetic as syn  # noqa: E402


def ref_net(opt, mode, cases):
    from lib.model import SuRSNet
    with R.quiet():
        net = SuRSNet(opt, projection_mode=mode).eval()
    case = cases[0]
    for mlp, wb in ((net.mlp_lr, case.mlp_lr), (net.mlp_hr, case.mlp_hr)):
        for i, (w, b) in enumerate(zip(*wb)):
            conv = getattr(mlp, "conv%d" % i)
            conv.weight.data = torch.from_numpy(w)[:, :, None].clone()
            conv.bias.data = torch.from_numpy(b).clone()
    net.im_feat_list_lr = [torch.from_numpy(np.stack([c.feat_lr for c in cases]))]
    net.im_feat_list_hr = [torch.from_numpy(np.stack([c.feat_hr for c in cases]))]
    return net


def main():
    R.import_reference()
    S = 32
    case = syn.SyntheticCase(S=S, seed=0)
    other = syn.SyntheticCase(S=S, seed=7)            # the second view's feature maps
    pts = syn.random_points(1500, seed=11, lo=-0.6, hi=0.6)
    out = {"points": pts}
    # ---- perspective: a pin-hole looking down -z from z = 2.5
    opt = R.make_opt(["--residual", "--loadSize", str(2 * S)])
    calib_p = np.array([[3.6, 0, 0, 0.02], [0, -3.6, 0, -0.01], [0, 0, 1.0, 2.5], [0, 0, 0, 1]], np.float32)
    net = ref_net(opt, "perspective", [case])
    with torch.no_grad(), R.quiet():
        net.query_mr(torch.from_numpy(pts)[None], torch.from_numpy(calib_p)[None])
        net.query_sr(torch.from_numpy(pts)[None], torch.from_numpy(calib_p)[None])
        hr, lr = net.get_preds()
    out.update(persp_calib=calib_p, persp_hr=hr[0, 0].numpy(), persp_lr=lr[0, 0].numpy())
    # ---- two views of one subject: front (the gen_mesh calib) and a camera rotated by 90 degrees about y
    opt2 = R.make_opt(["--residual", "--loadSize", str(2 * S), "--num_views", "2"])
    c0 = case.calib
    rot = np.array([[0, 0, 1, 0], [0, 1, 0, 0], [-1, 0, 0, 0], [0, 0, 0, 1]], np.float32)
    c1 = (c0 @ rot).astype(np.float32)
    c1[:3, 3] = [0.03, 0.0, -0.02]
    calibs = np.stack([c0, c1])
    net = ref_net(opt2, "orthogonal", [case, other])
    p2 = torch.from_numpy(np.stack([pts, pts]))
    with torch.no_grad(), R.quiet():
        net.query_mr(p2, torch.from_numpy(calibs))
        net.query_sr(p2, torch.from_numpy(calibs))
        hr, lr = net.get_preds()
    assert hr.shape == (2, 1, pts.shape[1])
    out.update(mv_calibs=calibs, mv_hr=hr[:, 0].numpy(), mv_lr=lr[:, 0].numpy(), mv_other_seed=7)
    np.savez_compressed(os.path.join(HERE, "variants_golden.npz"), **out)
    print("perspective: out-of-image", int((out["persp_hr"] == 0).sum()), "range", out["persp_hr"].min(), out["persp_hr"].max())
    print("multi-view: views differ where one mask is 0:", int(((out["mv_hr"][0] == 0) != (out["mv_hr"][1] == 0)).sum()))


if __name__ == "__main__":
    main()
