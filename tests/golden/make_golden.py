#!/usr/bin/env python
"""Generates the golden fixtures in tests/golden/ by running the UNMODIFIED reference
(/root/reference, imported in place with scikit-image stubbed -- see oracle/ref_import.py).

Run in the build container only:  python tests/golden/make_golden.py
The outputs (*.npz, *.txt) are committed; the reference mount does not exist on the GPU box.

Fixtures
  query_golden.npz    lib/model/SuRSNet.py query_mr + query_sr + get_preds on seeded points
                      (in- and out-of-image), synthetic weights/features of surs_b200.synthetic.
  grid_golden.npz     lib/sdf.py create_grid for several boxes / transforms.
  octree_golden.npz   lib/sdf.py eval_grid + eval_grid_octree driven by the analytic eval_func of
                      tests/helpers.py (64^3 with init_resolution=16 stored in full; 128^3 by hash).
  recon_golden.npz    lib/mesh_util.py reconstruction() volumes (marching cubes stubbed: skimage is
                      absent) for the real network on CPU, dense and octree, small resolution.
  obj_golden.txt      lib/mesh_util.py save_obj_mesh output for a small mesh.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import ref_import as R  # noqa: E402
import helpers  # noqa: E402
from surs_b200 import synthetic as syn  # noqa: E402


def ref_net(case, opt):
    from lib.model import SuRSNet
    with R.quiet():
        net = SuRSNet(opt).eval()
    for mlp, wb in ((net.mlp_lr, case.mlp_lr), (net.mlp_hr, case.mlp_hr)):
        for i, (w, b) in enumerate(zip(*wb)):
            conv = getattr(mlp, "conv%d" % i)
            conv.weight.data = torch.from_numpy(w)[:, :, None].clone()
            conv.bias.data = torch.from_numpy(b).clone()
    net.im_feat_list_lr = [torch.from_numpy(case.feat_lr)[None]]
    net.im_feat_list_hr = [torch.from_numpy(case.feat_hr)[None]]
    return net


def main():
    lib = R.import_reference()
    torch.manual_seed(0)

    # ---- query ------------------------------------------------------------
    S = 32
    case = syn.SyntheticCase(S=S, seed=0)
    opt = R.make_opt(["--residual", "--loadSize", str(2 * S)])
    net = ref_net(case, opt)
    pts = syn.random_points(2048, seed=3, lo=-0.6, hi=0.6)          # some fall outside the image
    pts[:, :8] = np.array([[-0.5, 0.5, 0.5, -0.5, 0.0, 0.5, -0.5, 0.25],
                           [-0.5, 0.5, -0.5, 0.5, 0.0, 0.0, 0.25, -0.5],
                           [0.0, 0.1, -0.2, 0.3, 0.0, 0.5, -0.5, 0.0]], dtype=np.float32)  # edges: |u|,|v| == 1
    calib = torch.from_numpy(case.calib)[None]
    with torch.no_grad(), R.quiet():
        net.query_mr(torch.from_numpy(pts)[None], calib)
        net.query_sr(torch.from_numpy(pts)[None], calib)
        hr, lr = net.get_preds()
    # a general (non-diagonal) calibration too
    calib2 = case.calib.copy()
    calib2[:3, :3] = calib2[:3, :3] @ np.array([[0.96, 0.0, 0.28], [0.0, 1.0, 0.0], [-0.28, 0.0, 0.96]], np.float32)
    calib2[:3, 3] = [0.05, -0.03, 0.1]
    with torch.no_grad(), R.quiet():
        net.query_mr(torch.from_numpy(pts)[None], torch.from_numpy(calib2)[None])
        net.query_sr(torch.from_numpy(pts)[None], torch.from_numpy(calib2)[None])
        hr2, lr2 = net.get_preds()
    np.savez_compressed(os.path.join(HERE, "query_golden.npz"), S=S, seed=0, points=pts,
                        pred_hr=hr[0, 0].numpy(), pred_lr=lr[0, 0].numpy(),
                        calib2=calib2, pred_hr2=hr2[0, 0].numpy(), pred_lr2=lr2[0, 0].numpy(),
                        feat_lr_sha=helpers.sha(case.feat_lr), feat_hr_sha=helpers.sha(case.feat_hr),
                        w_sha=helpers.sha(np.concatenate([w.ravel() for w in case.mlp_lr[0] + case.mlp_hr[0]])))
    print("query golden: hr range", float(hr.min()), float(hr.max()), "out-of-image", int((hr == 0).sum()))

    # ---- create_grid --------------------------------------------------------
    T = np.array([[0.9, 0.1, 0.0, 0.01], [-0.1, 0.9, 0.05, -0.02], [0.0, -0.05, 1.1, 0.03], [0, 0, 0, 1.0]])
    grids = {}
    for name, (res, bmin, bmax, tr) in {
        "unit16": ((16, 16, 16), [-0.5] * 3, [0.5] * 3, None),
        "aniso": ((8, 12, 20), [-1.0, -0.25, 0.0], [1.0, 1.75, 0.5], None),
        "pifu": ((32, 32, 32), [-128., -28., -128.], [128., 228., 128.], None),
        "xform": ((10, 10, 10), [-0.5] * 3, [0.5] * 3, T),
    }.items():
        c, m = lib.sdf.create_grid(*res, np.array(bmin), np.array(bmax), transform=tr)
        grids[name + "_coords"] = c
        grids[name + "_mat"] = m
        grids[name + "_args"] = np.array(list(res) + bmin + bmax, dtype=np.float64)
    grids["xform_T"] = T
    np.savez_compressed(os.path.join(HERE, "grid_golden.npz"), **grids)

    # ---- eval_grid / eval_grid_octree with the analytic field ----------------
    oct_out = {}
    o = types.SimpleNamespace(threshold=0.05)
    coords, _ = lib.sdf.create_grid(64, 64, 64, np.array([-0.5] * 3), np.array([0.5] * 3))
    hr_d, lr_d = lib.sdf.eval_grid(coords, helpers.analytic_eval_func, num_samples=50000)
    hr_o, lr_o = lib.sdf.eval_grid_octree(o, coords, helpers.analytic_eval_func, init_resolution=16, num_samples=50000)
    oct_out.update(dense64_hr=hr_d.astype(np.float32), dense64_lr=lr_d.astype(np.float32), oct64_hr=hr_o, oct64_lr=lr_o)
    o2 = types.SimpleNamespace(threshold=0.11)
    hr_o2, lr_o2 = lib.sdf.eval_grid_octree(o2, coords, helpers.analytic_eval_func, init_resolution=8, num_samples=7777)
    oct_out.update(oct64b_hr=hr_o2, oct64b_lr=lr_o2)
    coords, _ = lib.sdf.create_grid(128, 128, 128, np.array([-0.5] * 3), np.array([0.5] * 3))
    hr_o, lr_o = lib.sdf.eval_grid_octree(o, coords, helpers.analytic_eval_func, num_samples=50000)
    oct_out.update(oct128_hr_sha=helpers.sha(hr_o), oct128_lr_sha=helpers.sha(lr_o),
                   oct128_hr_zeros=int((hr_o == 0).sum()), oct128_lr_zeros=int((lr_o == 0).sum()),
                   oct128_hr_slice=hr_o[64], oct128_lr_slice=lr_o[:, 64])
    np.savez_compressed(os.path.join(HERE, "octree_golden.npz"), **oct_out)
    print("octree golden: 128^3 zero voxels", oct_out["oct128_hr_zeros"], oct_out["oct128_lr_zeros"])

    # ---- reconstruction() volumes through the real network -------------------
    captured = {}

    def fake_mc(vol, level):
        captured.setdefault("vols", []).append(np.array(vol))
        return (np.zeros((3, 3), np.float32), np.zeros((1, 3), np.int32), np.zeros((3, 3), np.float32),
                np.zeros(3, np.float32))

    lib.mesh_util.measure.marching_cubes_lewiner = fake_mc
    rec = {}
    b_min, b_max = np.array([-0.5] * 3), np.array([0.5] * 3)
    with torch.no_grad(), R.quiet():
        lib.mesh_util.reconstruction(opt, net, torch.device("cpu"), calib, 32, b_min, b_max,
                                     use_octree=False, num_samples=10000)
    rec["dense32_hr"], rec["dense32_lr"] = [v.astype(np.float32) for v in captured.pop("vols")]
    opt64 = R.make_opt(["--residual", "--loadSize", str(2 * S), "--threshold", "0.05"])
    with torch.no_grad(), R.quiet():
        lib.mesh_util.reconstruction(opt64, net, torch.device("cpu"), calib, 64, b_min, b_max,
                                     use_octree=True, num_samples=10000)
    # at 64^3 the reference octree degenerates to reso=1 (dense); keep it as the dense-64 golden
    rec["oct64_hr"], rec["oct64_lr"] = [v.astype(np.float32) for v in captured.pop("vols")]
    np.savez_compressed(os.path.join(HERE, "recon_golden.npz"), S=S, seed=0, **rec)
    print("recon golden: inside fraction hr/lr", float((rec["dense32_hr"] > 0.5).mean()), float((rec["dense32_lr"] > 0.5).mean()))

    # ---- OBJ writer ----------------------------------------------------------
    rng = np.random.default_rng(5)
    verts = rng.standard_normal((7, 3)) * 100.0
    verts[0] = [0.00005, -0.00005, 1.23455]
    faces = np.array([[0, 1, 2], [2, 3, 4], [4, 5, 6], [6, 0, 3]], dtype=np.int32)
    path = os.path.join(HERE, "obj_golden.txt")
    lib.mesh_util.save_obj_mesh(path, verts, faces)
    np.savez(os.path.join(HERE, "obj_golden_input.npz"), verts=verts, faces=faces)
    print("done")


if __name__ == "__main__":
    main()
