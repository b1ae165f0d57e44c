"""CPU baseline: the reference's query path restated with the SAME torch ops it uses, on the host.

TEST / BENCH INFRASTRUCTURE ONLY (``bench.py``'s ``cpu_baseline`` and ``--impl reference`` legs,
and a cross-check in tests/).  The reference is pure Python and is not present on the GPU box, so
its CPU path is reproduced op for op:
    lib/geometry.py:25      torch.baddbmm            (orthogonal)
    lib/model/SuRSNet.py:142 in-image mask
    lib/model/DepthNormalizer.py:18
    lib/geometry.py:11      F.grid_sample(align_corners=True) x4 (query_mr and query_sr both gather)
    lib/model/SurfaceClassifier.py:53-79  Conv1d(k=1) x5, torch.cat skip, leaky_relu, sigmoid
including the reference's redundancy (query_sr recomputes projection and both gathers), because
that is what its CPU time consists of.  Pinned by tests/test_oracle_golden.py against the goldens.
"""
import numpy as np
import torch
import torch.nn.functional as F


class TorchPort:
    def __init__(self, case, device="cpu"):
        self.dev = torch.device(device)
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(self.dev)
        self.f_lr = t(case.feat_lr)[None]
        self.f_hr = t(case.feat_hr)[None]
        self.calib = t(case.calib)[None]
        self.mlps = []
        for ws, bs in (case.mlp_lr, case.mlp_hr):
            self.mlps.append([(t(w)[:, :, None], t(b)) for w, b in zip(ws, bs)])
        self.z_num = float(case.load_size // 2)
        self.z_den = float(case.z_size)

    @staticmethod
    def _index(feat, uv):
        return F.grid_sample(feat, uv.transpose(1, 2).unsqueeze(2), align_corners=True)[:, :, :, 0]

    @staticmethod
    def _mlp(layers, feature, res_layers=(2, 3, 4)):
        y = feature
        for i, (w, b) in enumerate(layers):
            y = F.conv1d(torch.cat([y, feature], 1) if i in res_layers else y, w, b)
            if i != len(layers) - 1:
                y = F.leaky_relu(y)
        return torch.sigmoid(y)

    def _local(self, points, calib):
        xyz = torch.baddbmm(calib[:, :3, 3:4], calib[:, :3, :3], points)
        xy, z = xyz[:, :2, :], xyz[:, 2:3, :]
        in_img = (xy[:, 0] >= -1.0) & (xy[:, 0] <= 1.0) & (xy[:, 1] >= -1.0) & (xy[:, 1] <= 1.0)
        z_feat = z * self.z_num / self.z_den
        feat = torch.cat([torch.cat([self._index(self.f_lr, xy), self._index(self.f_hr, xy)], 1), z_feat], 1)
        return feat, in_img[:, None].float()

    @torch.no_grad()
    def query(self, points, calib=None):
        """points float32 [3,n] numpy -> (pred_hr, pred_lr) numpy [n]."""
        pts = torch.from_numpy(np.ascontiguousarray(points, dtype=np.float32)).to(self.dev)[None]
        calib = self.calib if calib is None else torch.from_numpy(np.asarray(calib, np.float32)).to(self.dev)[None]
        feat, mask = self._local(pts, calib)                       # query_mr
        pred_lr = mask * self._mlp(self.mlps[0], feat)
        feat2, mask2 = self._local(pts, calib)                     # query_sr recomputes everything
        pred_hr = mask2 * self._mlp(self.mlps[1], torch.cat([feat2, pred_lr], 1))
        return pred_hr[0, 0].cpu().numpy(), pred_lr[0, 0].cpu().numpy()
