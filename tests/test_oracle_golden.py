"""Pins the CPU oracle (oracle/surs_oracle.py) against vectors produced by the reference's own
modules (tests/golden/make_golden.py).  CPU only."""
import os
import types

import numpy as np
import pytest

import helpers
from oracle import surs_oracle as O
from surs_b200 import synthetic as syn

TOL_QUERY = 2e-5   # oracle is float64, the reference float32: measured 4e-6


@pytest.fixture(scope="module")
def case32():
    return syn.SyntheticCase(S=32, seed=0)


def test_synthetic_inputs_are_the_ones_the_goldens_were_made_with(golden_dir, case32):
    g = np.load(os.path.join(golden_dir, "query_golden.npz"))
    assert helpers.sha(case32.feat_lr) == str(g["feat_lr_sha"])
    assert helpers.sha(case32.feat_hr) == str(g["feat_hr_sha"])
    w = np.concatenate([w.ravel() for w in case32.mlp_lr[0] + case32.mlp_hr[0]])
    assert helpers.sha(w) == str(g["w_sha"])


def test_query_matches_reference(golden_dir, case32):
    g = np.load(os.path.join(golden_dir, "query_golden.npz"))
    c = case32
    hr, lr = O.query(g["points"], c.calib, c.feat_lr, c.feat_hr, c.mlp_lr, c.mlp_hr, load_size=c.load_size)
    assert np.abs(hr - g["pred_hr"]).max() < TOL_QUERY
    assert np.abs(lr - g["pred_lr"]).max() < TOL_QUERY
    # out-of-image points are exactly zero in both (mask multiply, SuRSNet.py:156,183)
    assert np.array_equal(hr == 0, g["pred_hr"] == 0)
    hr2, lr2 = O.query(g["points"], g["calib2"], c.feat_lr, c.feat_hr, c.mlp_lr, c.mlp_hr, load_size=c.load_size)
    assert np.abs(hr2 - g["pred_hr2"]).max() < TOL_QUERY
    assert np.abs(lr2 - g["pred_lr2"]).max() < TOL_QUERY


def test_create_grid_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "grid_golden.npz"))
    for name in ("unit16", "aniso", "pifu", "xform"):
        a = g[name + "_args"]
        res = [int(v) for v in a[:3]]
        tr = g["xform_T"] if name == "xform" else None
        coords, mat = O.create_grid(*res, a[3:6], a[6:9], transform=tr)
        assert np.array_equal(coords, g[name + "_coords"]), name
        assert np.array_equal(mat, g[name + "_mat"]), name


def test_eval_grid_and_octree_match_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "octree_golden.npz"))
    coords, _ = O.create_grid(64, 64, 64, np.array([-0.5] * 3), np.array([0.5] * 3))
    hr, lr = O.eval_grid(coords, helpers.analytic_eval_func, num_samples=50000)
    assert np.array_equal(hr.astype(np.float32), g["dense64_hr"]) and np.array_equal(lr.astype(np.float32), g["dense64_lr"])
    hr, lr = O.eval_grid_octree(0.05, coords, helpers.analytic_eval_func, init_resolution=16, num_samples=50000)
    assert np.array_equal(hr, g["oct64_hr"]) and np.array_equal(lr, g["oct64_lr"])
    hr, lr = O.eval_grid_octree(0.11, coords, helpers.analytic_eval_func, init_resolution=8, num_samples=7777)
    assert np.array_equal(hr, g["oct64b_hr"]) and np.array_equal(lr, g["oct64b_lr"])
    # the octree is NOT an approximation of the dense grid: zero holes
    assert (g["oct64_hr"] == 0).sum() > 0


def test_octree_128_matches_reference_hash(golden_dir):
    g = np.load(os.path.join(golden_dir, "octree_golden.npz"))
    coords, _ = O.create_grid(128, 128, 128, np.array([-0.5] * 3), np.array([0.5] * 3))
    hr, lr = O.eval_grid_octree(0.05, coords, helpers.analytic_eval_func, num_samples=50000)
    assert helpers.sha(hr) == str(g["oct128_hr_sha"])
    assert helpers.sha(lr) == str(g["oct128_lr_sha"])
    assert int((hr == 0).sum()) == int(g["oct128_hr_zeros"])


def test_octree_vectorised_equals_literal_loop():
    coords, _ = O.create_grid(32, 32, 32, np.array([-0.5] * 3), np.array([0.5] * 3))
    a = O.eval_grid_octree(0.05, coords, helpers.analytic_eval_func, init_resolution=8, num_samples=5000)
    b = O.eval_grid_octree_sequential(0.05, coords, helpers.analytic_eval_func, init_resolution=8, num_samples=5000)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    # R < init_resolution -> reso == 0 -> zeros (lib/sdf.py:66-68)
    z = O.eval_grid_octree(0.05, coords, helpers.analytic_eval_func, init_resolution=64)
    assert not z[0].any() and not z[1].any()


def test_reconstruction_volumes_match_reference(golden_dir, case32):
    g = np.load(os.path.join(golden_dir, "recon_golden.npz"))
    c = case32

    def eval_func(points):
        return O.query(points, c.calib, c.feat_lr, c.feat_hr, c.mlp_lr, c.mlp_hr, load_size=c.load_size)

    coords, _ = O.create_grid(32, 32, 32, np.array([-0.5] * 3), np.array([0.5] * 3))
    hr, lr = O.eval_grid(coords, eval_func, num_samples=10000)
    assert np.abs(hr - g["dense32_hr"]).max() < TOL_QUERY
    assert np.abs(lr - g["dense32_lr"]).max() < TOL_QUERY


def test_obj_text_matches_reference(golden_dir):
    inp = np.load(os.path.join(golden_dir, "obj_golden_input.npz"))
    with open(os.path.join(golden_dir, "obj_golden.txt")) as f:
        want = f.read()
    assert O.obj_text(inp["verts"], inp["faces"]) == want


def test_torch_port_matches_reference(golden_dir, case32):
    """The CPU-baseline port (oracle/torch_port.py) reproduces the reference's outputs."""
    from oracle import torch_port
    g = np.load(os.path.join(golden_dir, "query_golden.npz"))
    port = torch_port.TorchPort(case32)
    hr, lr = port.query(g["points"])
    assert np.abs(hr - g["pred_hr"]).max() < 1e-5 and np.abs(lr - g["pred_lr"]).max() < 1e-5
    hr2, lr2 = port.query(g["points"], g["calib2"])
    assert np.abs(hr2 - g["pred_hr2"]).max() < 1e-5 and np.abs(lr2 - g["pred_lr2"]).max() < 1e-5


def test_oracle_variants_match_reference_goldens(golden_dir):
    """Perspective projection and two views of one subject (lib/geometry.py:34-48, lib/model/SurfaceClassifier.py:70-76):
    the oracle's restatement against vectors recorded from the unmodified reference (tests/golden/make_golden_variants.py)."""
    from surs_b200 import synthetic as syn
    g = np.load(os.path.join(golden_dir, "variants_golden.npz"))
    case = syn.SyntheticCase(S=32, seed=0)
    other = syn.SyntheticCase(S=32, seed=int(g["mv_other_seed"]))
    pts = g["points"]
    hr, lr = O.query(pts, g["persp_calib"], case.feat_lr, case.feat_hr, case.mlp_lr, case.mlp_hr, load_size=case.load_size, perspective=True)
    assert np.abs(hr - g["persp_hr"]).max() < 1e-5 and np.abs(lr - g["persp_lr"]).max() < 1e-5
    assert np.array_equal(hr == 0, g["persp_hr"] == 0) and (g["persp_hr"] == 0).sum() > 10
    hr, lr = O.query_views(np.stack([pts, pts]), g["mv_calibs"], [case.feat_lr, other.feat_lr], [case.feat_hr, other.feat_hr],
                           case.mlp_lr, case.mlp_hr, load_size=case.load_size)
    assert hr.shape == g["mv_hr"].shape
    assert np.abs(hr - g["mv_hr"]).max() < 1e-5 and np.abs(lr - g["mv_lr"]).max() < 1e-5
    assert np.array_equal(hr == 0, g["mv_hr"] == 0)
