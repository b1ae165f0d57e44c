"""Drop-in for the reference's ``lib/geometry.py`` (``index``, ``orthogonal``, ``perspective``).

These torch versions serve the variants the fused CUDA query does not cover (multi-view,
image-space ``transforms``, perspective projection) and the training-time forward; the
inference hot path runs them inside csrc/query_*.cu instead.
"""
import torch
import torch.nn.functional as F


def index(feat, uv):
    """reference lib/geometry.py:4-12: feat [B,C,H,W], uv [B,2,N] in [-1,1] -> [B,C,N];
    bilinear, zero padding, align_corners=True (u indexes width, v height)."""
    grid = uv.transpose(1, 2).unsqueeze(2)
    return F.grid_sample(feat, grid, mode="bilinear", padding_mode="zeros", align_corners=True)[..., 0]


def _image_space(xy, transforms):
    return torch.baddbmm(transforms[:2, 2:3], transforms[:2, :2], xy)


def orthogonal(points, calibrations, transforms=None):
    """reference lib/geometry.py:15-31: R.p + t, optional 2x3 affine on (u,v)."""
    pts = torch.baddbmm(calibrations[:, :3, 3:4], calibrations[:, :3, :3], points)
    if transforms is not None:
        pts[:, :2, :] = _image_space(pts[:, :2, :], transforms)
    return pts


def perspective(points, calibrations, transforms=None):
    """reference lib/geometry.py:34-48 (never selected by the shipped configs)."""
    homo = torch.baddbmm(calibrations[:, :3, 3:4], calibrations[:, :3, :3], points)
    xy = homo[:, :2, :] / homo[:, 2:3, :]
    if transforms is not None:
        xy = _image_space(xy, transforms)
    return torch.cat([xy, homo[:, 2:3, :]], 1)
