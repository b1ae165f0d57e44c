"""Deterministic synthetic inputs for the reconstruction hot path.

There is no network: ``weights/netG_epoch_12`` is absent (reference
``.MISSING_LARGE_BLOBS``) and no dataset can be fetched, so tests and ``bench.py``
use the seeded inputs built here.  numpy only (plus torch for
the bilinear up-sampling of the feature fields); identical on every machine that
runs the same image.

* ``make_mlp``      -- a SurfaceClassifier state (5 conv1x1 layers with skip concat
                       at ``res_layers``), init as lib/net_util.py:99-132 does
                       (normal(0, 0.02), zero bias) but with a gain so that the
                       occupancy field saturates like a trained network's.
* ``make_features`` -- the two encoder outputs consumed by the path:
                       F_lr [256,S/2,S/2] (HGFilter 'low_res', SuRSNet.py:101-110)
                       and F_hr [64,2S,2S] (HGFilter 'high_res', :112-122).
* ``make_calib``    -- diag(2,-2,2,1), lib/train_util.py:63-66.
"""
from __future__ import annotations

import numpy as np

MLP_DIM_LR = (321, 1024, 512, 256, 128, 1)   # lib/options.py:88
MLP_DIM_HR = (322, 1024, 512, 256, 128, 1)   # lib/options.py:90
RES_LAYERS = (2, 3, 4)                       # lib/options.py:94-96


def make_mlp(dims, seed, gain=6.0, bias_std=0.1, res_layers=RES_LAYERS, z_gain=4.0, n_img_feat=320):
    """Returns (weights, biases): weights[i] float32 [Cout, Cin(+C0 on skip layers)]."""
    rng = np.random.default_rng(seed)
    ws, bs = [], []
    for l in range(len(dims) - 1):
        cin = dims[l] + (dims[0] if l in res_layers else 0)
        w = rng.standard_normal((dims[l + 1], cin)).astype(np.float32) * np.float32(0.02 * gain)
        # make depth (and, for the HR MLP, the LR prediction) matter next to 320 image channels
        zcol0 = cin - dims[0] + n_img_feat if l in res_layers else (n_img_feat if l == 0 else None)
        if zcol0 is not None:
            w[:, zcol0:] *= np.float32(z_gain)
        b = (rng.standard_normal(dims[l + 1]) * bias_std).astype(np.float32)
        ws.append(w)
        bs.append(b)
    return ws, bs


def _smooth_field(rng, channels, h, w, coarse):
    import torch
    import torch.nn.functional as F
    lo = torch.from_numpy(rng.standard_normal((1, channels, coarse, coarse)).astype(np.float32))
    return F.interpolate(lo, size=(h, w), mode="bilinear", align_corners=True)[0].numpy()


def silhouette(h, w):
    """A soft 'capsule person' mask in [0,1] on an h x w image (rows = v, cols = u)."""
    v, u = np.meshgrid(np.linspace(-1, 1, h), np.linspace(-1, 1, w), indexing="ij")
    body = ((u / 0.33) ** 2 + ((v - 0.12) / 0.62) ** 2)
    head = ((u / 0.16) ** 2 + ((v + 0.68) / 0.17) ** 2)
    d = np.minimum(body, head)
    return (1.0 / (1.0 + np.exp((d - 1.0) * 6.0))).astype(np.float32)


def make_features(S, seed=1991):
    """(F_lr [256,S/2,S/2], F_hr [64,2S,2S]) float32, C-contiguous NCHW planes."""
    rng = np.random.default_rng(seed)
    hl, hh = S // 2, 2 * S
    f_lr = _smooth_field(rng, 256, hl, hl, 12) * (0.25 + silhouette(hl, hl))[None]
    f_hr = _smooth_field(rng, 64, hh, hh, 24) * (0.25 + silhouette(hh, hh))[None]
    return np.ascontiguousarray(f_lr, dtype=np.float32), np.ascontiguousarray(f_hr, dtype=np.float32)


def make_calib():
    c = np.identity(4, dtype=np.float32) * 2
    c[1, 1] = -2
    c[3, 3] = 1
    return c


def random_points(n, seed=0, lo=-0.5, hi=0.5):
    """[3,n] float32 uniform in [lo,hi)^3 (BASELINE config 5)."""
    rng = np.random.default_rng(seed)
    return (rng.random((3, n), dtype=np.float32) * np.float32(hi - lo) + np.float32(lo)).astype(np.float32)


class SyntheticCase:
    """Everything one query needs, as numpy arrays."""

    def __init__(self, S=64, seed=0, gain=6.0, bias_std=0.1, z_gain=4.0, z_size=200.0):
        self.S = S
        self.load_size = 2 * S            # README.md:38: loadSize = 2 x input side
        self.z_size = float(z_size)
        self.calib = make_calib()
        self.feat_lr, self.feat_hr = make_features(S, seed=1991 + seed)
        self.mlp_lr = make_mlp(MLP_DIM_LR, 100 + seed, gain, bias_std, z_gain=z_gain)
        self.mlp_hr = make_mlp(MLP_DIM_HR, 200 + seed, gain, bias_std, z_gain=z_gain)
