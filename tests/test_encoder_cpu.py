"""The PyTorch image encoder (surs_b200.lib.model.encoder) against the reference's modules.
The comparison with the live reference runs in the build container only (the reference mount is
absent on the GPU box); the key / shape inventory is checked everywhere."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import ref_import as R

HERE = os.path.dirname(os.path.abspath(__file__))


def _opt():
    import types
    return types.SimpleNamespace(
        num_views=1, no_residual=False, mlp_dim_lr=[321, 1024, 512, 256, 128, 1], mlp_dim_hr=[322, 1024, 512, 256, 128, 1],
        mlp_res_layers_lr=[2, 3, 4], mlp_res_layers_hr=[2, 3, 4], loadSize=64, z_size=200.0, threshold=0.05,
        num_stack_lr=3, num_stack_hr=1, hg_depth=2, hg_dim=256, norm="group", n_block=[2, 2, 2], rgb_range=255, scale=2, residual=True)


def test_state_dict_inventory_matches_reference(golden_dir):
    from surs_b200.lib.model import SuRSNet
    net = SuRSNet(_opt())
    mine = {k: list(v.shape) for k, v in net.state_dict().items()}
    with open(os.path.join(golden_dir, "state_dict_keys.json")) as f:
        want = json.load(f)
    assert mine == want          # the reference checkpoint format: 553 tensors, same names and shapes


@pytest.mark.skipif(not R.available(), reason="reference mount not present (build container only)")
def test_encoder_matches_reference_forward():
    lib = R.import_reference()
    from lib.model import SuRSNet as RefNet
    from surs_b200.lib.model import SuRSNet
    opt = R.make_opt(["--residual", "--loadSize", "64"])
    torch.manual_seed(0)
    with R.quiet():
        ref = RefNet(opt).eval()
    net = SuRSNet(opt).eval()
    net.load_state_dict(ref.state_dict(), strict=True)
    img = torch.randn(1, 3, 32, 32)
    with torch.no_grad(), R.quiet():
        a = ref.super_res(img)
        ref.filter_hr(a[2]); ref.filter_lr(a[1])
        b = net.super_res(img)
        net.filter_hr(b[2]); net.filter_lr(b[1])
    for x, y in zip(a, b):
        assert torch.allclose(x, y, atol=1e-6, rtol=1e-5)
    assert len(net.im_feat_list_lr) == 1 and len(net.im_feat_list_hr) == 1
    assert torch.allclose(ref.im_feat_list_lr[0], net.im_feat_list_lr[0], atol=1e-5, rtol=1e-4)
    assert torch.allclose(ref.im_feat_list_hr[0], net.im_feat_list_hr[0], atol=1e-6, rtol=1e-5)
    # training mode keeps all three hourglass outputs (SuRSNet.py:109)
    ref.train(); net.train()
    with torch.no_grad(), R.quiet():
        ref.filter_lr(a[1]); net.filter_lr(b[1])
    assert len(net.im_feat_list_lr) == 3
    for x, y in zip(ref.im_feat_list_lr, net.im_feat_list_lr):
        assert torch.allclose(x, y, atol=1e-5, rtol=1e-4)
