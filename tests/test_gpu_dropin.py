"""GPU tests of the drop-in Python layer (surs_b200.lib.*): the reference's call shapes give the
reference's results.  Run on the B200 box: pytest -m gpu."""
import os
import types

import numpy as np
import pytest
import torch

import helpers
from oracle import mc_oracle
from oracle import surs_oracle as O
from surs_b200 import _capi
from surs_b200 import synthetic as syn
from surs_b200.lib import mesh_util, sdf, train_util
from surs_b200.lib.model import SuRSNet

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


class FakeEncoder(torch.nn.Module):
    """Stands in for SuRSSR_v3 + HGFilter: returns the synthetic feature maps."""

    def __init__(self, case):
        super().__init__()
        self.f_lr = torch.from_numpy(case.feat_lr)[None].to(DEV)
        self.f_hr = torch.from_numpy(case.feat_hr)[None].to(DEV)

    def super_res(self, images):
        return images, self.f_lr, self.f_hr

    def filter_lr(self, x):
        return [x * 0.5, x * 0.7, x]          # three hourglass outputs; eval mode keeps the last

    def filter_hr(self, x):
        return [x]


def make_net(case, precision=None, encoder=None):
    opt = helpers.make_opt(loadSize=case.load_size, z_size=case.z_size)
    net = (SuRSNet(opt, encoder=encoder) if precision is None else SuRSNet(opt, precision=precision, encoder=encoder)).to(DEV).eval()
    sd = {}
    for name, (ws, bs) in (("mlp_lr", case.mlp_lr), ("mlp_hr", case.mlp_hr)):
        for i, (w, b) in enumerate(zip(ws, bs)):
            sd["%s.conv%d.weight" % (name, i)] = torch.from_numpy(w)[:, :, None]
            sd["%s.conv%d.bias" % (name, i)] = torch.from_numpy(b)
    missing, unexpected = net.load_state_dict(sd, strict=False)      # reference checkpoint keys for the MLPs
    assert not unexpected and not [k for k in missing if k.startswith("mlp_")]
    net.im_feat_list_lr = [torch.from_numpy(case.feat_lr)[None].to(DEV)]
    net.im_feat_list_hr = [torch.from_numpy(case.feat_hr)[None].to(DEV)]
    return opt, net


@pytest.fixture(scope="module")
def case32():
    return syn.SyntheticCase(S=32, seed=0)


def test_query_api_matches_reference_semantics(case32, golden_dir):
    g = np.load(os.path.join(golden_dir, "query_golden.npz"))
    opt, net = make_net(case32, precision=_capi.PREC_FP32)
    pts = torch.from_numpy(g["points"])[None].to(DEV)
    calib = torch.from_numpy(case32.calib)[None].to(DEV)
    net.query_mr(pts, calib)
    assert net.preds_lr.shape == (1, 1, pts.shape[2])
    net.query_sr(pts, calib)
    hr, lr = net.get_preds()                                       # HR first (BaseSuRSNet.py:85)
    assert np.abs(hr[0, 0].cpu().numpy() - g["pred_hr"]).max() < 2e-5
    assert np.abs(lr[0, 0].cpu().numpy() - g["pred_lr"]).max() < 2e-5
    assert torch.equal(net.query(pts, calib), hr)                  # PIFu alias
    # the torch path for uncovered variants gives the same numbers (and warns)
    net.train()
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False                        # cuDNN's TF32 convs are ~5e-4 off
    try:
        with pytest.warns(UserWarning):
            net.query_mr(pts, calib)
            net.query_sr(pts, calib)
    finally:
        torch.backends.cudnn.allow_tf32 = tf32
    net.eval()
    assert np.abs(net.preds_hr[0, 0].detach().cpu().numpy() - g["pred_hr"]).max() < 1e-4


def test_default_precision_meets_the_tolerance_through_the_public_api(case32, golden_dir):
    """SuRSNet() without a precision argument = SURS_PREC_FP16R; its query_mr / query_sr / get_preds agree with the
    reference's golden predictions within THE tolerance (helpers.TOL = 1e-3), flips reported."""
    g = np.load(os.path.join(golden_dir, "query_golden.npz"))
    opt, net = make_net(case32)
    assert net.precision == _capi.PREC_FP16R == _capi.PREC_DEFAULT
    pts = torch.from_numpy(g["points"])[None].to(DEV)
    calib = torch.from_numpy(case32.calib)[None].to(DEV)
    net.query_mr(pts, calib)
    net.query_sr(pts, calib)
    hr, lr = net.get_preds()
    helpers.parity_report(hr[0, 0].cpu(), g["pred_hr"], label="public API, default precision, HR")
    helpers.parity_report(lr[0, 0].cpu(), g["pred_lr"], label="public API, default precision, LR")


def test_public_api_variants_run_on_the_fused_kernels(case32, golden_dir):
    """SURVEY §8(f)-4 through the reference's own call shapes: SuRSNet(opt, 'perspective'), opt.num_views = 2 with
    [V,...] features / calibs / points, and query_mr / query_sr with an image-space `transforms` -- all without the
    torch fallback (no warning), against the reference goldens / the oracle."""
    import warnings
    g = np.load(os.path.join(golden_dir, "variants_golden.npz"))
    pts = torch.from_numpy(g["points"])[None].to(DEV)
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        # perspective
        opt = helpers.make_opt(loadSize=case32.load_size, z_size=case32.z_size)
        _, base = make_net(case32, precision=_capi.PREC_FP32)
        net = SuRSNet(opt, "perspective", precision=_capi.PREC_FP32, encoder=None).to(DEV).eval()
        net.load_state_dict(base.state_dict())
        net.im_feat_list_lr, net.im_feat_list_hr = base.im_feat_list_lr, base.im_feat_list_hr
        calib = torch.from_numpy(g["persp_calib"])[None].to(DEV)
        net.query_mr(pts, calib)
        net.query_sr(pts, calib)
        hr, lr = net.get_preds()
        assert np.abs(hr[0, 0].cpu().numpy() - g["persp_hr"]).max() < 2e-5 and np.abs(lr[0, 0].cpu().numpy() - g["persp_lr"]).max() < 2e-5
        # two views of one subject
        other = syn.SyntheticCase(S=32, seed=int(g["mv_other_seed"]))
        opt2 = helpers.make_opt(loadSize=case32.load_size, z_size=case32.z_size, num_views=2)
        net2 = SuRSNet(opt2, precision=_capi.PREC_FP32, encoder=None).to(DEV).eval()
        net2.load_state_dict(base.state_dict())
        net2.im_feat_list_lr = [torch.from_numpy(np.stack([case32.feat_lr, other.feat_lr])).to(DEV)]
        net2.im_feat_list_hr = [torch.from_numpy(np.stack([case32.feat_hr, other.feat_hr])).to(DEV)]
        p2 = pts.expand(2, -1, -1).contiguous()
        c2 = torch.from_numpy(g["mv_calibs"]).to(DEV)
        net2.query_mr(p2, c2)
        assert net2.preds_lr.shape == (2, 1, pts.shape[2])
        net2.query_sr(p2, c2)
        hr, lr = net2.get_preds()
        assert np.abs(hr[:, 0].cpu().numpy() - g["mv_hr"]).max() < 2e-5 and np.abs(lr[:, 0].cpu().numpy() - g["mv_lr"]).max() < 2e-5
        # image-space transforms on the single-view net
        T = torch.tensor([[0.9, 0.05, 0.02], [-0.04, 1.1, -0.03]], device=DEV)
        calib = torch.from_numpy(case32.calib)[None].to(DEV)
        base.query_mr(pts, calib, transforms=T)
        base.query_sr(pts, calib, transforms=T)
        hr, lr = base.get_preds()
        ohr, olr = O.query(g["points"], case32.calib, case32.feat_lr, case32.feat_hr, case32.mlp_lr, case32.mlp_hr, load_size=case32.load_size,
                           uv_transform=T.cpu().numpy())
        assert np.abs(hr[0, 0].cpu().numpy() - ohr).max() < 2e-5 and np.abs(lr[0, 0].cpu().numpy() - olr).max() < 2e-5
        # ... and the next plain call is not affected by the transform of the previous one
        plain = base.query(pts, calib)
        ohr, _ = O.query(g["points"], case32.calib, case32.feat_lr, case32.feat_hr, case32.mlp_lr, case32.mlp_hr, load_size=case32.load_size)
        assert np.abs(plain[0, 0].cpu().numpy() - ohr).max() < 2e-5


def test_cached_state_is_invalidated_soundly(case32):
    """The library keeps packed weights, repacked features and query_mr's HR result; none of them may go stale:
    weights edited through .data (no version bump), a new feature tensor that reuses a freed tensor's memory,
    features replaced behind the net's back, query_sr on other points than query_mr saw."""
    opt, net = make_net(case32, precision=_capi.PREC_FP32)
    calib = torch.from_numpy(case32.calib)[None].to(DEV)
    pts = torch.from_numpy(syn.random_points(3000, seed=3))[None].to(DEV)
    base = net.query(pts, calib).clone()
    # 1. in-place write through .data: the version counter does not move, the checksum does
    w = net.mlp_hr.conv4.weight
    v0 = w._version
    w.data.mul_(0.5)
    assert w._version == v0
    changed = net.query(pts, calib).clone()
    assert not torch.equal(changed, base)
    w.data.mul_(2.0)
    assert torch.equal(net.query(pts, calib), base)
    # 2. load_state_dict / .to() mark the weights dirty
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    sd["mlp_lr.conv0.bias"] += 0.25
    net.load_state_dict(sd)
    assert not torch.equal(net.query(pts, calib), base)
    sd["mlp_lr.conv0.bias"] -= 0.25
    net.load_state_dict(sd)
    assert torch.allclose(net.query(pts, calib), base, atol=1e-6)
    base = net.query(pts, calib).clone()
    # 3. a new feature tensor in recycled memory (same data_ptr / version / shape as the freed one)
    other = syn.SyntheticCase(S=32, seed=5)
    net.im_feat_list_lr, net.im_feat_list_hr = [], []
    ptrs = set()
    f_lr = torch.from_numpy(case32.feat_lr)[None].to(DEV); f_hr = torch.from_numpy(case32.feat_hr)[None].to(DEV)
    net.im_feat_list_lr, net.im_feat_list_hr = [f_lr], [f_hr]
    a = net.query(pts, calib).clone()
    ptrs.add(f_lr.data_ptr())
    del f_lr, f_hr
    net.im_feat_list_lr, net.im_feat_list_hr = [], []
    g_lr = torch.from_numpy(other.feat_lr)[None].to(DEV); g_hr = torch.from_numpy(other.feat_hr)[None].to(DEV)
    net.im_feat_list_lr, net.im_feat_list_hr = [g_lr], [g_hr]
    b = net.query(pts, calib).clone()
    print("recycled feature memory:", g_lr.data_ptr() in ptrs)
    assert not torch.equal(a, b)
    # 4. somebody replaces the context's features behind the net's back
    ctx = net.surs_context()
    ctx.set_features(torch.from_numpy(case32.feat_lr).to(DEV), torch.from_numpy(case32.feat_hr).to(DEV))
    assert torch.equal(net.query(pts, calib), b)                       # the net re-uploads ITS features
    # 5. query_sr on other points than query_mr: recomputed, never the cached HR of the other points
    pts2 = torch.from_numpy(syn.random_points(3000, seed=4))[None].to(DEV)
    want2 = net.query(pts2, calib).clone()
    net.query_mr(pts, calib)
    net.query_sr(pts2, calib)
    assert torch.equal(net.get_preds()[0], want2)
    # ... also when the second tensor reuses the first one's memory
    net.query_mr(pts, calib)
    p_tmp = pts.clone()
    net.query_mr(p_tmp, calib)
    del p_tmp
    p_new = pts2.clone()
    net.query_sr(p_new, calib)
    assert torch.equal(net.get_preds()[0], want2)
    # 6. in-place edit of the points between query_mr and query_sr
    p3 = pts.clone()
    net.query_mr(p3, calib)
    p3.copy_(pts2)
    net.query_sr(p3, calib)
    assert torch.equal(net.get_preds()[0], want2)


@pytest.mark.parametrize("use_octree,res,prec", [(False, 64, _capi.PREC_FP16R), (True, 128, _capi.PREC_FP16R), (False, 64, _capi.PREC_FP16),
                                                 (False, 64, _capi.PREC_FP16X3), (True, 128, _capi.PREC_FP16X3)])
def test_reconstruction_fast_path(case32, use_octree, res, prec):
    opt, net = make_net(case32, precision=prec)
    calib = torch.from_numpy(case32.calib)[None].to(DEV)
    b_min, b_max = np.array([-0.5] * 3), np.array([0.5] * 3)
    out, stats = mesh_util.reconstruction(opt, net, DEV, calib, res, b_min, b_max, use_octree=use_octree, return_stats=True)
    assert len(out) == 8
    ctx = net.surs_context()
    zn, zd = net.depth_scale()
    if use_octree:
        a, b, n_eval = ctx.eval_grid_octree((res,) * 3, b_min, b_max, calib, zn, zd, opt.threshold, precision=prec)
        vols = (a.float(), b.float())
        assert stats["n_evaluated"] == n_eval < res ** 3
    else:
        vols = ctx.eval_grid((res,) * 3, b_min, b_max, calib, zn, zd, precision=prec)
    _, mat = O.create_grid(res, res, res, b_min, b_max)
    for k, vol in enumerate(vols):
        v, f, n, val = mc_oracle.marching_cubes_lewiner(vol.cpu().numpy(), 0.5)
        verts, faces, normals, values = out[4 * k:4 * k + 4]
        assert verts.dtype == np.float64 and faces.dtype == np.int32 and normals.dtype == np.float32
        assert np.array_equal(faces, f)
        assert np.allclose(verts, O.verts_to_world(mat, v), rtol=0, atol=1e-12)
        assert np.array_equal(values, val)
    # PIFu call shape (no opt) gives the same mesh
    out2 = mesh_util.reconstruction(net, DEV, calib, res, b_min, b_max, use_octree)
    assert all(np.array_equal(x, y) for x, y in zip(out, out2))


def test_reconstruction_generic_path_with_foreign_net():
    """Any object with query_mr / query_sr / get_preds works (eval_func closure, lib/mesh_util.py:20-28)."""

    class Foreign:
        num_views = 1

        def query_mr(self, samples, calib):
            self.p = samples

        def query_sr(self, samples, calib):
            pass

        def get_preds(self):
            hr, lr = helpers.analytic_eval_func(self.p[0].cpu().numpy().astype(np.float64))
            return torch.from_numpy(hr), torch.from_numpy(lr)

    opt = helpers.make_opt()
    b_min, b_max = np.array([-0.5] * 3), np.array([0.5] * 3)
    for use_octree in (False, True):
        out = mesh_util.reconstruction(opt, Foreign(), DEV, None, 64, b_min, b_max, use_octree=use_octree, num_samples=30000)
        coords, mat = O.create_grid(64, 64, 64, b_min, b_max)
        if use_octree:
            hr, lr = O.eval_grid_octree(opt.threshold, coords, helpers.analytic_eval_func, num_samples=30000)
        else:
            hr, lr = O.eval_grid(coords, helpers.analytic_eval_func, num_samples=30000)
        v, f, _, _ = mc_oracle.marching_cubes_lewiner(hr.astype(np.float32), 0.5)
        assert np.array_equal(out[1], f) and np.allclose(out[0], O.verts_to_world(mat, v), atol=1e-12)
        v, f, _, _ = mc_oracle.marching_cubes_lewiner(lr.astype(np.float32), 0.5)
        assert np.array_equal(out[5], f)


def test_sdf_module_matches_oracle():
    coords, mat = sdf.create_grid(16, 12, 20, np.array([-1.0, 0, 0.5]), np.array([1.0, 2, 1.5]))
    oc, om = O.create_grid(16, 12, 20, np.array([-1.0, 0, 0.5]), np.array([1.0, 2, 1.5]))
    assert np.array_equal(coords, oc) and np.array_equal(mat, om)
    assert np.array_equal(sdf.grid_matrix((16, 12, 20), [-1.0, 0, 0.5], [1.0, 2, 1.5]), om)
    coords, _ = sdf.create_grid(32, 32, 32, np.array([-0.5] * 3), np.array([0.5] * 3))
    opt = helpers.make_opt()
    a = sdf.eval_grid_octree(opt, coords, helpers.analytic_eval_func, init_resolution=8, num_samples=5000)
    b = O.eval_grid_octree_sequential(0.05, coords, helpers.analytic_eval_func, init_resolution=8, num_samples=5000)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_marching_cubes_errors_mirror_skimage():
    with pytest.raises(ValueError):
        mesh_util.marching_cubes_lewiner(np.zeros((8, 8, 8), np.float32), 0.5)
    v, f, n, val = mesh_util.marching_cubes_lewiner(helpers.sphere_volume(24, 7.2), 0.5)
    assert helpers.mesh_euler_closed(v, f) == 2


def test_gen_mesh_writes_reference_obj(case32, tmp_path):
    opt, net = make_net(case32, encoder=FakeEncoder(case32))
    opt.resolution = 64
    data = {"img_LR": torch.zeros(1, 3, 32, 32), "b_min": np.array([-0.5] * 3), "b_max": np.array([0.5] * 3)}
    path = str(tmp_path / "subject.obj")
    train_util.gen_mesh(opt, net, DEV, data, path, use_octree=False)
    calib = train_util.make_calib(DEV)
    out = mesh_util.reconstruction(opt, net, DEV, calib, 64, data["b_min"], data["b_max"], use_octree=False)
    with open(path[:-4] + "_HR.obj") as fh:
        assert fh.read() == O.obj_text(out[0], out[1])
    with open(path[:-4] + "_LR.obj") as fh:
        assert fh.read() == O.obj_text(out[4], out[5])


def test_gen_mesh_with_builtin_encoder(tmp_path):
    """The whole reference flow (lib/train_util.py:53-85) with the PyTorch encoder + CUDA reconstruction."""
    opt = helpers.make_opt(loadSize=128, num_stack_lr=3, num_stack_hr=1, hg_depth=2, hg_dim=256, norm="group",
                           n_block=[2, 2, 2], rgb_range=255, scale=2, residual=True, resolution=64)
    torch.manual_seed(0)
    net = SuRSNet(opt).to(DEV).eval()
    for mlp in (net.mlp_lr, net.mlp_hr):                     # widen the occupancy range of the random-init MLPs
        for conv in mlp.layers():
            conv.weight.data *= 6.0
        mlp.conv4.bias.data += 0.3                           # inside the image: occupancy > 0.5 ...
    g = torch.Generator().manual_seed(1)
    # ... and the box is larger than the image, where the mask forces occupancy 0: a surface always exists
    data = {"img_LR": torch.rand(1, 3, 64, 64, generator=g) * 2 - 1, "b_min": np.array([-0.7] * 3), "b_max": np.array([0.7] * 3)}
    path = str(tmp_path / "person.obj")
    with torch.no_grad():
        train_util.gen_mesh(opt, net, DEV, data, path, use_octree=True)
    assert net.im_feat_list_lr[0].shape == (1, 256, 32, 32) and net.im_feat_list_hr[0].shape == (1, 64, 128, 128)
    for suffix in ("_HR.obj", "_LR.obj"):
        with open(path[:-4] + suffix) as fh:
            lines = fh.read().splitlines()
        assert lines[0].startswith("v ") and lines[-1].startswith("f ") and len(lines) > 100


def test_fast_encoder_mode_matches_eager(tmp_path):
    """encoder_mode='fast' (TF32 convolutions + ONE CUDA graph of super_res + filter_hr + filter_lr) produces the eager
    modules' feature maps up to TF32 rounding (2e-3 relative, the reference's own GPU numerics), replays correctly for
    new images and after a weight update."""
    import types
    opt = types.SimpleNamespace(**vars(helpers.make_opt(loadSize=64, resolution=64)), num_stack_lr=3, num_stack_hr=1, hg_depth=2, hg_dim=256,
                                norm="group", n_block=[2, 2, 2], rgb_range=255, scale=2, residual=True)
    torch.manual_seed(0)
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False                   # the eager side in true fp32
    eager = SuRSNet(opt, encoder_mode="eager").to(DEV).eval()
    for m in eager.modules():
        if isinstance(m, torch.nn.Conv2d) and m.weight.shape[1] > 3:
            torch.nn.init.kaiming_normal_(m.weight, a=0.2)
    fast = SuRSNet(opt, encoder_mode="fast").to(DEV).eval()
    fast.load_state_dict(eager.state_dict())
    rel = lambda a, b: float((a - b).norm() / b.norm())
    with torch.no_grad():
        for seed in (1, 2, 1):
            img = torch.randn(1, 3, 32, 32, generator=torch.Generator().manual_seed(seed)).to(DEV)
            for n in (eager, fast):
                _, flr, fhr = n.super_res(img)
                n.filter_hr(fhr)
                n.filter_lr(flr)
            e_lr, e_hr = rel(fast.im_feat_list_lr[-1], eager.im_feat_list_lr[-1]), rel(fast.im_feat_list_hr[0], eager.im_feat_list_hr[0])
            print("fast encoder vs eager: relative L2 error lr %.3g hr %.3g" % (e_lr, e_hr))
            assert e_lr < 5e-3 and e_hr < 5e-3
            assert fast.im_feat_list_lr[-1].dtype == torch.float32 and fast.im_feat_list_lr[-1].is_contiguous()
        assert len(fast._enc_graphs) == 1                      # one capture, replayed
        # a weight update is seen by the captured graph (in-place load_state_dict)
        sd = {k: v.clone() for k, v in eager.state_dict().items()}
        sd["image_filter_hr.conv5.bias"] += 1.0
        eager.load_state_dict(sd)
        fast.load_state_dict(sd)
        for n in (eager, fast):
            _, flr, fhr = n.super_res(img)
            n.filter_hr(fhr)
        assert rel(fast.im_feat_list_hr[0], eager.im_feat_list_hr[0]) < 5e-3
        # a different feature tensor than the one super_res returned goes through the eager filter
        other = fhr.clone()
        fast.filter_hr(other)
        assert rel(fast.im_feat_list_hr[0], eager.image_filter_hr(other)[-1]) < 5e-3
    torch.backends.cudnn.allow_tf32 = tf32


def test_eval_cli_end_to_end(tmp_path):
    """apps/eval_SuRS.py: image folder + checkpoint file -> OBJ files (reference flags)."""
    import importlib.util
    from PIL import Image
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("eval_surs_app", os.path.join(root, "apps", "eval_SuRS.py"))
    app = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(app)
    from surs_b200.lib.options import BaseOptions
    os.makedirs(tmp_path / "data" / "image_final")
    os.makedirs(tmp_path / "data" / "mask_final")
    rng = np.random.default_rng(0)
    for name in ("p0", "p1"):
        Image.fromarray(rng.integers(0, 256, (64, 64, 3), dtype=np.uint8)).save(tmp_path / "data" / "image_final" / (name + ".png"))
        Image.fromarray(np.full((64, 64), 255, np.uint8)).save(tmp_path / "data" / "mask_final" / (name + ".png"))
    args = ["--dataroot", str(tmp_path / "data"), "--results_path", str(tmp_path / "res"), "--name", "t", "--residual",
            "--resolution", "64", "--loadSize", "128", "--b_min", "-0.7", "-0.7", "-0.7", "--b_max", "0.7", "0.7", "0.7"]
    torch.manual_seed(0)
    net = SuRSNet(BaseOptions().parse(args))
    for mlp in (net.mlp_lr, net.mlp_hr):
        for conv in mlp.layers():
            conv.weight.data *= 6.0
        mlp.conv4.bias.data += 0.3
    ckpt = str(tmp_path / "netG_epoch_12")
    torch.save(net.state_dict(), ckpt)                     # the reference's checkpoint format (flat state_dict)
    done = app.main(args + ["--load_netG_checkpoint_path", ckpt])
    assert len(done) == 2
    for p in done:
        assert os.path.getsize(p[:-4] + "_HR.obj") > 1000 and os.path.getsize(p[:-4] + "_LR.obj") > 1000
