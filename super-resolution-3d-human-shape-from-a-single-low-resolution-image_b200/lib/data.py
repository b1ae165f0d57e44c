"""Input contract of the reference's eval dataset (lib/data/EvalDataset_LR_v2.py:185-254,389-410)
without torchvision / trimesh: ``<dataroot>/image_final/X.{jpg,png}`` + ``<dataroot>/mask_final/X.{png,jpg}``
-> ``{'name', 'img_LR' [1,3,S,S] in [-1,1] multiplied by the mask, 'calib' [1,4,4], 'b_min', 'b_max'}``."""
import os

import numpy as np
import torch
from PIL import Image


def _first_existing(folder, stem, exts):
    for e in exts:
        p = os.path.join(folder, stem + e)
        if os.path.isfile(p):
            return p
    raise FileNotFoundError(os.path.join(folder, stem + exts[0]))


def to_tensor(img):
    """PIL image -> float tensor [C,H,W] in [0,1] (what torchvision's ToTensor does for uint8 images)."""
    a = np.array(img, dtype=np.uint8)
    if a.ndim == 2:
        a = a[:, :, None]
    return torch.from_numpy(np.ascontiguousarray(a.transpose(2, 0, 1))).float().div(255)


class EvalImageFolder:
    def __init__(self, opt):
        self.render = os.path.join(opt.dataroot, "image_final")
        self.mask = os.path.join(opt.dataroot, "mask_final")
        self.b_min = np.array(opt.b_min, dtype=float)
        self.b_max = np.array(opt.b_max, dtype=float)
        self.subjects = sorted(os.listdir(self.render))

    def __len__(self):
        return len(self.subjects)

    def __getitem__(self, index):
        stem = os.path.splitext(self.subjects[index])[0]
        image = Image.open(_first_existing(self.render, stem, (".jpg", ".png"))).convert("RGB")
        mask = Image.open(_first_existing(self.mask, stem, (".png", ".jpg"))).convert("L")
        m = to_tensor(mask)
        rgb = (to_tensor(image) - 0.5) / 0.5                      # Normalize((0.5,)*3, (0.5,)*3)
        calib = torch.diag(torch.tensor([2.0, -2.0, 2.0, 1.0]))
        return {"name": stem, "img_LR": (m.expand_as(rgb) * rgb)[None], "calib": calib[None],
                "b_min": self.b_min, "b_max": self.b_max}
