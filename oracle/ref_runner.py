"""TEST / BENCH INFRASTRUCTURE ONLY.  Runs the UNMODIFIED reference's own reconstruction path on the host cores:

    lib.mesh_util.reconstruction(opt, net, cpu, calib, resolution, b_min, b_max, use_octree, num_samples)
        (lib/mesh_util.py:8-49: create_grid -> eval_grid / eval_grid_octree -> batch_eval -> eval_func ->
         SuRSNet.query_mr + query_sr + get_preds -> marching cubes x2 -> world transform)

from ``/root/reference`` (build container) or its verbatim copy ``oracle/_ref`` (GPU box; ``oracle/make_ref.py``).
The network is the reference's ``lib.model.SuRSNet`` built from the reference's ``BaseOptions``; the two feature maps
and the MLP parameters of the synthetic case are placed exactly where ``filter_lr`` / ``filter_hr`` / ``load_state_dict``
would put them (the image encoder is not part of the measured path).  scikit-image is not installed: the reference's
``measure.marching_cubes_lewiner`` call is routed to the oracle's C twin (``oracle/mc_oracle.c``, one thread, like
skimage's Cython loop) -- every number produced here says so (``mc = "mc_oracle"``).
"""
import time

import numpy as np
import torch

from . import ref_import as R


def available():
    return R.available()


class ReferenceRun:
    def __init__(self, case, threads=None):
        if threads:
            torch.set_num_threads(int(threads))
        self.lib = R.import_reference()
        if R.have_real_skimage():
            self.mc = "skimage"
        else:
            from . import mc_oracle
            R.use_marching_cubes(mc_oracle.marching_cubes_lewiner)
            self.mc = "mc_oracle"
        self.opt = R.make_opt(["--residual", "--loadSize", str(case.load_size), "--z_size", str(case.z_size),
                               "--num_samples", "50000", "--threshold", "0.05"])
        from lib.model import SuRSNet                      # the reference's class (REF_ROOT is on sys.path)
        with R.quiet():
            net = SuRSNet(self.opt).eval()
        for mlp, wb in ((net.mlp_lr, case.mlp_lr), (net.mlp_hr, case.mlp_hr)):
            for i, (w, b) in enumerate(zip(*wb)):
                conv = getattr(mlp, "conv%d" % i)
                conv.weight.data = torch.from_numpy(w)[:, :, None].clone()
                conv.bias.data = torch.from_numpy(b).clone()
        net.im_feat_list_lr = [torch.from_numpy(case.feat_lr)[None]]
        net.im_feat_list_hr = [torch.from_numpy(case.feat_hr)[None]]
        self.net = net
        self.calib = torch.from_numpy(case.calib)[None]
        self.threads = torch.get_num_threads()

    def reconstruction(self, resolution, use_octree=False, num_samples=50000, b_min=(-0.5,) * 3, b_max=(0.5,) * 3):
        """One call of the reference's reconstruction; returns (8-tuple, seconds)."""
        b_min, b_max = np.asarray(b_min, np.float64), np.asarray(b_max, np.float64)
        t0 = time.perf_counter()
        with torch.no_grad(), R.quiet():
            out = self.lib.mesh_util.reconstruction(self.opt, self.net, torch.device("cpu"), self.calib, resolution, b_min, b_max,
                                                    use_octree=use_octree, num_samples=num_samples)
        return out, time.perf_counter() - t0
