"""The eval input contract (surs_b200.lib.data) on CPU."""
import os

import numpy as np
import torch
from PIL import Image

from surs_b200.lib.data import EvalImageFolder
from surs_b200.lib.options import BaseOptions


def test_eval_image_folder(tmp_path):
    os.makedirs(tmp_path / "image_final")
    os.makedirs(tmp_path / "mask_final")
    rng = np.random.default_rng(0)
    rgb = rng.integers(0, 256, (16, 16, 3), dtype=np.uint8)
    mask = (rng.random((16, 16)) > 0.5).astype(np.uint8) * 255
    Image.fromarray(rgb).save(tmp_path / "image_final" / "b_subject.png")
    Image.fromarray(mask).save(tmp_path / "mask_final" / "b_subject.png")
    Image.fromarray(rgb[::-1].copy()).save(tmp_path / "image_final" / "a_subject.png")
    Image.fromarray(mask).save(tmp_path / "mask_final" / "a_subject.png")
    opt = BaseOptions().parse(["--dataroot", str(tmp_path), "--b_min", "-0.5", "-0.5", "-0.5", "--b_max", "0.5", "0.5", "0.5"])
    ds = EvalImageFolder(opt)
    assert len(ds) == 2 and ds[0]["name"] == "a_subject"
    item = ds[1]
    want = (torch.from_numpy(rgb.transpose(2, 0, 1)).float() / 255 - 0.5) / 0.5 * (torch.from_numpy(mask).float() / 255)[None]
    assert item["img_LR"].shape == (1, 3, 16, 16) and torch.equal(item["img_LR"][0], want)
    assert torch.equal(item["calib"][0], torch.diag(torch.tensor([2.0, -2.0, 2.0, 1.0])))
    assert item["b_min"].dtype == np.float64 and np.array_equal(item["b_max"], [0.5, 0.5, 0.5])
    # defaults mirror the reference's options
    d = BaseOptions().parse([])
    assert d.resolution == 512 and d.num_samples == 50000 and d.threshold == 0.05 and d.mlp_dim_hr[0] == 322 and d.z_size == 200.0
