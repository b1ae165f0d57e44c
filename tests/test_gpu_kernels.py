"""GPU parity tests: every CUDA path is driven through the C-ABI (surs_b200._capi.Context) and
compared with the CPU oracle / golden fixtures.  Run on the B200 box: pytest -m gpu."""
import os

import numpy as np
import pytest
import torch

import helpers
from oracle import mc_oracle
from oracle import surs_oracle as O
from surs_b200 import synthetic as syn

pytestmark = pytest.mark.gpu

# THE tolerance (helpers.TOL = 1e-3, the north_star's): pre-threshold occupancy against the reference's fp32 path, for
# the default precision SURS_PREC_FP16R (dense grids: on the nodes marching cubes reads; point queries / octrees: on
# every point), SURS_PREC_FP16X3 and SURS_PREC_FP32; inside / outside flips only within TOL of 0.5, count reported.
TOL = helpers.TOL
# Regression guards (NOT tolerances: tighter bounds on what the kernels measurably achieve, so a numerical regression
# is caught long before it reaches TOL).  FP32 mode: an fp32 FMA chain like the reference's (measured 1.0e-5).
TOL_FP32 = 2e-5
# FP16X3 (split hi/lo fp16 operands, three tensor-core passes, fp32 accumulate): measured 6.0e-5 max, 1.1e-6 mean.
TOL_X3_MAX = 1e-4
TOL_X3_MEAN = 5e-6
X3_FLIP_BAND = 1e-4
# SURS_PREC_FP16 (ONE pass, fp16 operands) is an explicit opt-in that does NOT meet TOL: measured 1.5e-2 max on the
# synthetic saturating weights (profiles/r1_parity_report.json).  Its bounds below only guard against regressions.
TOL_FP16_MAX = 2e-2
TOL_FP16_MEAN = 5e-4
FLIP_BAND = 1e-2


@pytest.fixture(scope="module")
def ctx():
    from surs_b200 import _capi
    c = _capi.Context("cuda:0")
    yield c
    c.close()


@pytest.fixture(scope="module")
def case32(ctx):
    case = syn.SyntheticCase(S=32, seed=0)
    load_case(ctx, case)
    return case


def load_case(ctx, case):
    dev = ctx.device
    t = lambda a: torch.from_numpy(a).to(dev)
    ctx.set_weights([t(w) for w in case.mlp_lr[0]], [t(b) for b in case.mlp_lr[1]],
                    [t(w) for w in case.mlp_hr[0]], [t(b) for b in case.mlp_hr[1]],
                    syn.MLP_DIM_LR, syn.MLP_DIM_HR, syn.RES_LAYERS)
    ctx.set_features(t(case.feat_lr), t(case.feat_hr))


def znum(case):
    return float(case.load_size // 2), float(case.z_size)


def test_umma_cta_pair_selftest(ctx):
    """tcgen05.mma.cta_group::2 (one cluster of two CTAs, M = 256): building block for the next kernels."""
    from surs_b200 import _capi
    g = torch.Generator().manual_seed(1)
    for (N, K) in ((256, 64), (256, 128), (128, 192), (64, 64)):
        A = torch.randn(256, K, generator=g)
        B = torch.randn(N, K, generator=g)
        D = _capi.selftest_umma2(ctx, A, B).cpu()
        assert (D - A.half().float() @ B.half().float().T).abs().max().item() < 2e-3, (N, K)


def test_umma_selftest(ctx):
    """The tcgen05 plumbing in isolation: descriptors, swizzled K-major layout, TMEM read-back."""
    g = torch.Generator().manual_seed(0)
    for (N, K, tail) in ((256, 64, False), (256, 128, False), (144, 128, False), (256, 80, True), (144, 192, False), (64, 16, True)):
        A = torch.randn(128, K, generator=g)
        B = torch.randn(N, K, generator=g)
        D = ctx.selftest_umma(A, B, tail16=tail).cpu()
        ref = A.half().float() @ B.half().float().T
        err = (D - ref).abs().max().item()
        assert err < 2e-3, (N, K, tail, err)


@pytest.mark.parametrize("prec", ["fp32", "fp16", "default"])
def test_query_matches_golden_and_oracle(ctx, case32, golden_dir, prec):
    from surs_b200 import _capi
    g = np.load(os.path.join(golden_dir, "query_golden.npz"))
    pts = torch.from_numpy(g["points"]).to(ctx.device)
    assert _capi.PREC_DEFAULT == _capi.PREC_FP16R
    p = {"fp32": _capi.PREC_FP32, "fp16": _capi.PREC_FP16, "default": _capi.PREC_DEFAULT}[prec]
    for calib, khr, klr in ((case32.calib, "pred_hr", "pred_lr"), (g["calib2"], "pred_hr2", "pred_lr2")):
        hr, lr = ctx.query(pts, calib, *znum(case32), precision=p)
        hr, lr = hr.cpu().numpy(), lr.cpu().numpy()
        dh, dl = np.abs(hr - g[khr]), np.abs(lr - g[klr])
        print("query %s: max|d| hr %.3g lr %.3g  mean hr %.3g lr %.3g" % (prec, dh.max(), dl.max(), dh.mean(), dl.mean()))
        if prec == "fp32":
            assert dh.max() < TOL_FP32 and dl.max() < TOL_FP32
        elif prec == "default":                       # the reference's golden predictions, THE tolerance
            helpers.parity_report(hr, g[khr], label="default precision, HR")
            helpers.parity_report(lr, g[klr], label="default precision, LR")
        else:
            assert dh.max() < TOL_FP16_MAX and dl.max() < TOL_FP16_MAX
            assert dh.mean() < TOL_FP16_MEAN and dl.mean() < TOL_FP16_MEAN
        # out-of-image points are exactly 0 (mask multiply), bit exact in both modes
        assert np.array_equal(hr == 0, g[khr] == 0) and np.array_equal(lr == 0, g[klr] == 0)


@pytest.mark.parametrize("n", [1, 127, 128, 129, 5000, 40000])
def test_query_ragged_sizes_fp16_vs_fp32(ctx, case32, n):
    from surs_b200 import _capi
    pts = torch.from_numpy(syn.random_points(n, seed=n)).to(ctx.device)
    a = ctx.query(pts, case32.calib, *znum(case32), precision=_capi.PREC_FP32)
    b = ctx.query(pts, case32.calib, *znum(case32), precision=_capi.PREC_FP16)
    for x, y in zip(a, b):
        d = (x - y).abs()
        assert d.max().item() < TOL_FP16_MAX, d.max().item()
    # classification at 0.5 must agree except for near-threshold points (reported)
    flips = ((a[0] > 0.5) != (b[0] > 0.5))
    near = (a[0] - 0.5).abs() < FLIP_BAND
    assert not (flips & ~near).any()


def test_query_empty_and_host_path(ctx, case32):
    from surs_b200 import _capi
    hr, lr = ctx.query(torch.empty(3, 0, device=ctx.device), case32.calib, *znum(case32))
    assert hr.numel() == 0 and lr.numel() == 0
    pts = syn.random_points(3000, seed=9)
    h1, l1 = ctx.query_host(pts, case32.calib, *znum(case32), precision=_capi.PREC_FP32)
    h2, l2 = ctx.query(torch.from_numpy(pts).to(ctx.device), case32.calib, *znum(case32), precision=_capi.PREC_FP32)
    assert np.array_equal(h1, h2.cpu().numpy()) and np.array_equal(l1, l2.cpu().numpy())


def test_dense_grid_matches_reference_volumes(ctx, case32, golden_dir):
    from surs_b200 import _capi
    g = np.load(os.path.join(golden_dir, "recon_golden.npz"))
    for prec, tol in ((_capi.PREC_FP32, TOL_FP32), (_capi.PREC_FP16, TOL_FP16_MAX), (_capi.PREC_DEFAULT, TOL)):
        hr, lr = ctx.eval_grid((32, 32, 32), [-0.5] * 3, [0.5] * 3, case32.calib, *znum(case32), precision=prec)
        assert np.abs(hr.cpu().numpy() - g["dense32_hr"]).max() < tol
        assert np.abs(lr.cpu().numpy() - g["dense32_lr"]).max() < tol
    # slab evaluation = the corresponding planes of the full grid, bit exact
    full = ctx.eval_grid((32, 32, 32), [-0.5] * 3, [0.5] * 3, case32.calib, *znum(case32), precision=_capi.PREC_FP16)
    slab = ctx.eval_grid((32, 32, 32), [-0.5] * 3, [0.5] * 3, case32.calib, *znum(case32), precision=_capi.PREC_FP16,
                         plane_lo=7, plane_hi=20)
    assert torch.equal(full[0][7:20], slab[0]) and torch.equal(full[1][7:20], slab[1])
    # with a transform the grid equals explicit points from the oracle's create_grid
    T = np.array([[0.9, 0.1, 0.0, 0.01], [-0.1, 0.9, 0.05, -0.02], [0.0, -0.05, 1.1, 0.03], [0, 0, 0, 1.0]])
    coords, _ = O.create_grid(16, 12, 20, np.array([-0.5] * 3), np.array([0.5, 0.4, 0.5]), transform=T)
    pts = torch.from_numpy(coords.reshape(3, -1).astype(np.float32)).to(ctx.device)
    a = ctx.query(pts, case32.calib, *znum(case32), precision=_capi.PREC_FP32)
    b = ctx.eval_grid((16, 12, 20), [-0.5] * 3, [0.5, 0.4, 0.5], case32.calib, *znum(case32), transform=T, precision=_capi.PREC_FP32)
    assert torch.equal(a[0], b[0].reshape(-1)) and torch.equal(a[1], b[1].reshape(-1))


def test_column_kernels_match_reference_reconstruction_volumes(ctx, case32, golden_dir):
    """The 64^3 volumes captured from the UNMODIFIED reference's lib.mesh_util.reconstruction (real SuRSNet, CPU fp32;
    tests/golden/make_golden.py) against the column-factored kernels, one-pass and split-operand, dense and through
    the octree entry point (which at R = 64 = init_resolution evaluates every node, lib/sdf.py:66-79)."""
    from surs_b200 import _capi
    g = np.load(os.path.join(golden_dir, "recon_golden.npz"))
    args = ((64, 64, 64), [-0.5] * 3, [0.5] * 3, case32.calib) + znum(case32)
    for prec, tol in ((_capi.PREC_FP16, TOL_FP16_MAX), (_capi.PREC_FP16X3, TOL_X3_MAX), (_capi.PREC_FP32, TOL_FP32)):
        hr, lr = ctx.eval_grid(*args, precision=prec)
        ohr, olr, n_eval = ctx.eval_grid_octree(*args, threshold=0.05, init_resolution=64, precision=prec)
        assert n_eval == 64 ** 3
        for got, want in ((hr, g["oct64_hr"]), (lr, g["oct64_lr"]), (ohr.float(), g["oct64_hr"]), (olr.float(), g["oct64_lr"])):
            d = np.abs(got.cpu().numpy() - want)
            print("precision %d vs reference reconstruction volume: max|d| %.3g mean %.3g" % (prec, d.max(), d.mean()))
            assert d.max() < tol
        if prec != _capi.PREC_FP32:                       # dense and indexed column kernels: node for node identical
            assert torch.equal(hr, ohr.float()) and torch.equal(lr, olr.float())
    # the default precision (SURS_PREC_FP16R) against the same reference volumes, THE tolerance: every node marching
    # cubes reads a value from, and every inside / outside bit outside the reported near-threshold band
    hr, lr = ctx.eval_grid(*args, precision=_capi.PREC_DEFAULT)
    st = ctx.refine_stats
    print("default precision: %d nodes refined, max |one-pass - split| %.3g, band %.3g" % (st["nodes"], st["max_diff"], st["band"]))
    assert not st["fell_back"] and st["max_diff"] < 0.8 * st["band"]
    for got, want, name in ((hr, g["oct64_hr"], "HR"), (lr, g["oct64_lr"], "LR")):
        want = torch.from_numpy(want).to(got.device)
        helpers.parity_report(got, want, mask=helpers.mc_read_mask(want > 0.5), label="default precision dense 64^3 %s (nodes marching cubes reads)" % name)


def test_octree_blocks_match_reference_golden(ctx, golden_dir):
    """select / cells kernels driven by the analytic eval_func: bit exact vs lib/sdf.py's output."""
    g = np.load(os.path.join(golden_dir, "octree_golden.npz"))
    coords, _ = O.create_grid(64, 64, 64, np.array([-0.5] * 3), np.array([0.5] * 3))
    flat = coords.reshape(3, -1)
    for thr, init, kh, kl in ((0.05, 16, "oct64_hr", "oct64_lr"), (0.11, 8, "oct64b_hr", "oct64b_lr")):
        res = (64, 64, 64)
        hr = torch.zeros(res, device=ctx.device, dtype=torch.float64)
        lr = torch.zeros(res, device=ctx.device, dtype=torch.float64)
        dirty = torch.ones(res, device=ctx.device, dtype=torch.uint8)
        idx = torch.empty(64 ** 3, device=ctx.device, dtype=torch.int64)
        reso = 64 // init
        while reso > 0:
            n = ctx.octree_select(res, reso, dirty, idx)
            sel = idx[:n].cpu().numpy()
            a, b = helpers.analytic_eval_func(flat[:, sel])
            hr.view(-1)[idx[:n]] = torch.from_numpy(a.reshape(-1).astype(np.float64)).to(ctx.device)
            lr.view(-1)[idx[:n]] = torch.from_numpy(b.reshape(-1).astype(np.float64)).to(ctx.device)
            if reso <= 1:
                break
            ctx.octree_cells(res, reso, thr, hr, lr, dirty)
            reso //= 2
        assert np.array_equal(hr.cpu().numpy(), g[kh]), kh
        assert np.array_equal(lr.cpu().numpy(), g[kl]), kl


def test_fused_octree_equals_oracle_octree_on_same_occupancies(ctx, case32):
    from surs_b200 import _capi
    for prec in (_capi.PREC_FP32, _capi.PREC_FP16, _capi.PREC_FP16X3):
        res = (64, 64, 64)
        hr, lr, n_eval = ctx.eval_grid_octree(res, [-0.5] * 3, [0.5] * 3, case32.calib, *znum(case32), threshold=0.05,
                                              init_resolution=16, precision=prec)

        if prec == _capi.PREC_FP32:
            def eval_func(points):
                p = torch.from_numpy(np.ascontiguousarray(points, dtype=np.float32)).to(ctx.device)
                a, b = ctx.query(p, case32.calib, *znum(case32), precision=prec)
                return a.cpu().numpy(), b.cpu().numpy()
        else:
            # the tensor-core octree runs the indexed variant of the column kernel: node for node it must be
            # bit-identical to the dense column kernel (same table, same per-row arithmetic)
            dense = [v.cpu().numpy() for v in ctx.eval_grid(res, [-0.5] * 3, [0.5] * 3, case32.calib, *znum(case32), precision=prec)]

            def eval_func(points):
                ijk = np.rint((np.asarray(points, dtype=np.float64) + 0.5) * 64).astype(np.int64)
                return dense[0][ijk[0], ijk[1], ijk[2]], dense[1][ijk[0], ijk[1], ijk[2]]

        stats = []
        coords, _ = O.create_grid(*res, np.array([-0.5] * 3), np.array([0.5] * 3))
        ohr, olr = O.eval_grid_octree(0.05, coords, eval_func, init_resolution=16, num_samples=50000, stats=stats)
        assert np.array_equal(hr.cpu().numpy(), ohr) and np.array_equal(lr.cpu().numpy(), olr)
        assert n_eval == sum(s[1] for s in stats)
        assert n_eval < 64 ** 3       # something was pruned
    # R < init_resolution: zeros, as the reference
    z = ctx.eval_grid_octree((32, 32, 32), [-0.5] * 3, [0.5] * 3, case32.calib, *znum(case32), threshold=0.05, init_resolution=64)
    assert not z[0].any() and not z[1].any() and z[2] == 0


def _mc_check(ctx, vol, level=0.5, mat=None):
    v, f, n, val, st = mc_oracle.marching_cubes_lewiner(vol, level, return_stats=True)
    gv, gw, gf, gn, gval, gamb = ctx.marching_cubes(torch.from_numpy(vol).to(ctx.device), level, mat)
    assert gf.shape[0] == f.shape[0] and gv.shape[0] == v.shape[0]
    assert np.array_equal(gf.cpu().numpy(), f)                      # topology bit exact
    assert np.abs(gv.cpu().numpy() - v).max() <= 1e-5               # positions (bit exact in practice)
    assert np.array_equal(gv.cpu().numpy(), v)
    assert np.abs(gn.cpu().numpy() - n).max() < 1e-4
    assert np.array_equal(gval.cpu().numpy(), val)
    assert gamb == st["ambiguous_cells"]
    if mat is not None:
        assert np.allclose(gw.cpu().numpy(), O.verts_to_world(mat, v), rtol=0, atol=1e-12)
    return v, f


def test_marching_cubes_matches_cpu_twin(ctx):
    vol = helpers.sphere_volume(48, 15.3)
    _, mat = O.create_grid(48, 48, 48, np.array([-0.5] * 3), np.array([0.5] * 3))
    v, f = _mc_check(ctx, vol, mat=mat)
    assert helpers.mesh_euler_closed(v, f) == 2
    rng = np.random.default_rng(0)
    noisy = rng.random((24, 20, 28)).astype(np.float32)             # many ambiguous cells, ragged shape
    noisy[0] = noisy[-1] = 0; noisy[:, 0] = noisy[:, -1] = 0; noisy[:, :, 0] = noisy[:, :, -1] = 0
    v, f = _mc_check(ctx, noisy)
    helpers.mesh_euler_closed(v, f)
    _mc_check(ctx, np.ascontiguousarray(noisy[:, :, :27]))          # last axis not a multiple of 4: general kernels
    g = np.stack(np.meshgrid(*[np.arange(40)] * 3, indexing="ij")).astype(np.float64)
    wavy = (np.sin(g[0] * 0.7) + np.sin(g[1] * 0.9) + np.sin(g[2] * 0.8)).astype(np.float32)   # open surface at the border
    _mc_check(ctx, wavy, level=0.1)
    _mc_check(ctx, np.ascontiguousarray(vol[23:25]))                    # minimal thickness


def test_marching_cubes_passes_the_table_free_checker(ctx):
    """The kernels' meshes against the VOLUME alone (oracle/mc_check.py reads no table): vertex placement, one cell per
    triangle, closed oriented manifold, and the topology of the ambiguous cells (face deciders + interior / tunnel test)
    against the trilinear interpolant; interior statistics equal the CPU twin's."""
    from oracle import mc_check
    rng = np.random.default_rng(4)
    vol = rng.random((12, 9, 16)).astype(np.float32)
    gv, _, gf, _, _, namb = ctx.marching_cubes(torch.from_numpy(vol).to(ctx.device), 0.5)
    n_int, n_tun = ctx.mc_interior_stats()
    v, f, _, _, st = mc_oracle.marching_cubes_lewiner(vol, 0.5, return_stats=True)
    assert np.array_equal(gv.cpu().numpy(), v) and np.array_equal(gf.cpu().numpy(), f)
    assert (namb, n_int, n_tun) == (st["ambiguous_cells"], st["interior_ambiguous_cells"], st["tunnel_cells"])
    assert n_int > 50 and n_tun > 0
    rep = mc_check.check_mesh(vol, 0.5, gv.cpu().numpy(), gf.cpu().numpy())
    top = mc_check.check_cell_topology(vol, 0.5, gv.cpu().numpy(), gf.cpu().numpy(), max_cells=200)
    print(rep, {k: top[k] for k in top if k != "mismatches"})
    assert top["cells_checked"] > 100 and top["mismatch"] == 0, top["mismatches"]
    # general (non-fast) kernels: last axis not a multiple of 4
    vol2 = np.ascontiguousarray(vol[:, :, :15])
    gv, _, gf, _, _, _ = ctx.marching_cubes(torch.from_numpy(vol2).to(ctx.device), 0.5)
    assert ctx.mc_interior_stats() == tuple(mc_oracle.marching_cubes_lewiner(vol2, 0.5, return_stats=True)[4][k] for k in ("interior_ambiguous_cells", "tunnel_cells"))
    mc_check.check_mesh(vol2, 0.5, gv.cpu().numpy(), gf.cpu().numpy())
    # a smooth closed surface: every check, no tolerance on orientation
    sph = helpers.sphere_volume(32, 10.2)
    gv, _, gf, _, _, _ = ctx.marching_cubes(torch.from_numpy(sph).to(ctx.device), 0.5)
    rep = mc_check.check_mesh(sph, 0.5, gv.cpu().numpy(), gf.cpu().numpy())
    assert rep["orientation_wrong"] == 0 and rep["border_edges"] == 0


def test_marching_cubes_float64_source_and_value_range(ctx):
    """The bit pass also yields what skimage derives from every value: the float32 copy of a float64 volume (octree
    volumes stay float64, lib/sdf.py:60-61) and the value range the level is checked against -- fast and general path."""
    rng = np.random.default_rng(7)
    for shape in ((20, 24, 64), (9, 10, 27)):
        v64 = torch.from_numpy(rng.random(shape) * 1.3 - 0.2).to(ctx.device)
        v32 = v64.float()
        a = ctx.mc_count(v64, 0.5)
        ra = ctx.mc_value_range()
        va, _, _, _ = ctx.mc_emit_verts(a[0])
        fa = ctx.mc_emit_faces(a[1])
        assert torch.equal(ctx._mc_vol, v32)
        b = ctx.mc_count(v32, 0.5)
        rb = ctx.mc_value_range()
        vb, _, _, _ = ctx.mc_emit_verts(b[0])
        fb = ctx.mc_emit_faces(b[1])
        assert a == b and torch.equal(va, vb) and torch.equal(fa, fb)
        assert ra == rb == (float(v32.min()), float(v32.max()))
    # repeated volumes of very different surface size: the list-length estimate of the previous call must not matter
    big = torch.from_numpy(rng.random((32, 32, 64)).astype(np.float32)).to(ctx.device)
    small = torch.from_numpy(helpers.sphere_volume(64, 5.0)[:32, :32]).to(ctx.device).contiguous()
    want_big, want_small = None, None
    for vol in (small, big, small, big):
        got = ctx.marching_cubes(vol, 0.5)
        ref = mc_oracle.marching_cubes_lewiner(vol.cpu().numpy(), 0.5) if vol is big or got[0].shape[0] else None
        assert np.array_equal(got[2].cpu().numpy(), ref[1]) and np.array_equal(got[0].cpu().numpy(), ref[0])


def test_marching_cubes_slabs_reproduce_single_volume(ctx):
    """Three slabs with the seam protocol (SURS_MC_LOWER_FOREIGN + seam maps) == one volume."""
    from surs_b200 import _capi
    rng = np.random.default_rng(1)
    vol = helpers.sphere_volume(40, 13.1) + 0.2 * rng.random((40, 40, 40)).astype(np.float32)
    v, f, _, _ = mc_oracle.marching_cubes_lewiner(vol, 0.5)
    cuts = [0, 13, 27, 40]
    dev = ctx.device
    verts, faces, offset, seam = [], [], 0, None
    for r in range(3):
        lo, hi = cuts[r], min(cuts[r + 1] + 1, 40)                  # + 1 halo plane
        slab = torch.from_numpy(np.ascontiguousarray(vol[lo:hi])).to(dev)
        nv, nf, _ = ctx.mc_count(slab, 0.5, flags=_capi.MC_LOWER_FOREIGN if r > 0 else 0)
        seam_out = torch.empty((2, 40, 40), device=dev, dtype=torch.int32)
        vv, _, _, _ = ctx.mc_emit_verts(nv, None, vert_id_offset=offset, seam_out=seam_out, plane_offset=cuts[r])
        ff = ctx.mc_emit_faces(nf, seam_in=seam)
        verts.append(vv.cpu().numpy())
        faces.append(ff.cpu().numpy())
        offset += nv
        seam = seam_out
    assert np.array_equal(np.concatenate(verts), v)
    assert np.array_equal(np.concatenate(faces), f)


def test_marching_cubes_no_surface_gives_zero_counts(ctx):
    vol = torch.zeros((8, 8, 8), device=ctx.device)
    assert ctx.mc_count(vol, 0.5)[:2] == (0, 0)


def test_column_factored_dense_path(ctx, case32):
    """surs_eval_grid takes the column-factored kernels (query_col.cu) when (u,v) do not depend on
    the last grid axis; same occupancies as the fp32 mode / the generic tensor-core kernel."""
    from surs_b200 import _capi
    for res, bmax in (((64, 64, 64), [0.5, 0.5, 0.5]), ((5, 9, 200), [0.5, 0.4, 0.55]), ((3, 130, 128), [0.2, 0.5, 0.5])):
        args = (res, [-0.5] * 3, bmax, case32.calib) + znum(case32)
        col = ctx.eval_grid(*args, precision=_capi.PREC_FP16)
        ref = ctx.eval_grid(*args, precision=_capi.PREC_FP32)
        coords, _ = O.create_grid(*res, np.array([-0.5] * 3), np.array(bmax))
        pts = torch.from_numpy(coords.reshape(3, -1).astype(np.float32)).to(ctx.device)
        gen = ctx.query(pts, case32.calib, *znum(case32), precision=_capi.PREC_FP16)
        for a, b, g in zip(col, ref, gen):
            d = (a - b).abs()
            print("column path %s: max|d| vs fp32 %.3g mean %.3g; vs generic tc %.3g" %
                  (res, d.max().item(), d.mean().item(), (a.reshape(-1) - g).abs().max().item()))
            assert d.max().item() < TOL_FP16_MAX and d.mean().item() < TOL_FP16_MEAN
            assert np.array_equal((a == 0).cpu().numpy(), (b == 0).cpu().numpy())
        # a slab of the same grid is bit identical to the corresponding planes
        if res[0] >= 5:
            slab = ctx.eval_grid(*args, precision=_capi.PREC_FP16, plane_lo=1, plane_hi=4)
            assert torch.equal(col[0][1:4], slab[0]) and torch.equal(col[1][1:4], slab[1])


def test_split_operand_mode_dense_and_octree(ctx, case32):
    """SURS_PREC_FP16X3: A.W as A_hi.W_hi + A_lo.W_hi + A_hi.W_lo on the tensor cores (query_col.cu, P = 3).
    Occupancies within TOL_X3_MAX of the fp32 mode and of the float64 oracle; slabs bit identical; the octree
    node for node identical to the dense volume (covered by the fused-octree test as well)."""
    from surs_b200 import _capi
    for res, bmax in (((64, 64, 64), [0.5, 0.5, 0.5]), ((5, 9, 200), [0.5, 0.4, 0.55]), ((3, 130, 128), [0.2, 0.5, 0.5])):
        args = (res, [-0.5] * 3, bmax, case32.calib) + znum(case32)
        x3 = ctx.eval_grid(*args, precision=_capi.PREC_FP16X3)
        ref = ctx.eval_grid(*args, precision=_capi.PREC_FP32)
        one = ctx.eval_grid(*args, precision=_capi.PREC_FP16)
        for a, b, c in zip(x3, ref, one):
            d = (a - b).abs()
            flips = (a > 0.5) != (b > 0.5)
            print("split-operand path %s: max|d| vs fp32 %.3g mean %.3g (one-pass fp16: %.3g); flips %d" %
                  (res, d.max().item(), d.mean().item(), (c - b).abs().max().item(), int(flips.sum())))
            assert d.max().item() < TOL_X3_MAX and d.mean().item() < TOL_X3_MEAN
            assert np.array_equal((a == 0).cpu().numpy(), (b == 0).cpu().numpy())
            assert not flips.any() or ((b[flips] - 0.5).abs() < X3_FLIP_BAND).all()
        if res[0] >= 5:
            slab = ctx.eval_grid(*args, precision=_capi.PREC_FP16X3, plane_lo=1, plane_hi=4)
            assert torch.equal(x3[0][1:4], slab[0]) and torch.equal(x3[1][1:4], slab[1])
    # against the float64 oracle on a small grid
    res, bmax = (2, 8, 64), [0.5, 0.5, 0.5]
    x3 = ctx.eval_grid(res, [-0.5] * 3, bmax, case32.calib, *znum(case32), precision=_capi.PREC_FP16X3)
    coords, _ = O.create_grid(*res, np.array([-0.5] * 3), np.array(bmax))
    pts = coords.reshape(3, -1).astype(np.float32)
    ohr, olr = O.query(pts, case32.calib, case32.feat_lr, case32.feat_hr, case32.mlp_lr, case32.mlp_hr, load_size=case32.load_size)
    err = max(np.abs(x3[0].cpu().numpy().reshape(-1) - ohr).max(), np.abs(x3[1].cpu().numpy().reshape(-1) - olr).max())
    print("split-operand path vs float64 oracle: max|d| %.3g" % err)
    assert err < TOL_X3_MAX
    # point sources without a column structure (explicit points incl. out-of-image ones and ragged counts, transformed
    # grids, octree lists of a transformed grid, a calibration with shear): per-point tables through the same kernels
    for n in (1, 127, 129, 5000, 140000):
        p = torch.from_numpy(syn.random_points(n, seed=n, lo=-0.55, hi=0.55)).to(ctx.device)
        a = ctx.query(p, case32.calib, *znum(case32), precision=_capi.PREC_FP16X3)
        b = ctx.query(p, case32.calib, *znum(case32), precision=_capi.PREC_FP32)
        err = max((a[0] - b[0]).abs().max().item(), (a[1] - b[1]).abs().max().item())
        print("split-operand point query n=%d: max|d| vs fp32 %.3g" % (n, err))
        assert err < TOL_X3_MAX
        assert np.array_equal((a[0] == 0).cpu().numpy(), (b[0] == 0).cpu().numpy())
    T = np.array([[0.9, 0.1, 0.0, 0.01], [-0.1, 0.9, 0.05, -0.02], [0.0, -0.05, 1.1, 0.03], [0, 0, 0, 1.0]])
    args = ((64, 64, 64), [-0.5] * 3, [0.5] * 3, case32.calib) + znum(case32)
    a = ctx.eval_grid(*args, transform=T, precision=_capi.PREC_FP16X3)
    b = ctx.eval_grid(*args, transform=T, precision=_capi.PREC_FP32)
    assert max((a[0] - b[0]).abs().max().item(), (a[1] - b[1]).abs().max().item()) < TOL_X3_MAX
    oa = ctx.eval_grid_octree(*args, threshold=0.05, init_resolution=16, transform=T, precision=_capi.PREC_FP16X3)
    coords, _ = O.create_grid(64, 64, 64, np.array([-0.5] * 3), np.array([0.5] * 3), transform=T)

    def eval_func(points):                                   # the same per-point arithmetic through surs_query
        q = torch.from_numpy(np.ascontiguousarray(points, dtype=np.float32)).to(ctx.device)
        u, v = ctx.query(q, case32.calib, *znum(case32), precision=_capi.PREC_FP16X3)
        return u.cpu().numpy(), v.cpu().numpy()
    ohr, olr = O.eval_grid_octree(0.05, coords, eval_func, init_resolution=16, num_samples=50000)
    assert oa[2] < 64 ** 3 and np.array_equal(oa[0].cpu().numpy(), ohr) and np.array_equal(oa[1].cpu().numpy(), olr)
    shear = case32.calib.copy()
    shear[0, 2] = 0.05
    p = torch.from_numpy(syn.random_points(3000, seed=3, lo=-0.5, hi=0.5)).to(ctx.device)
    a = ctx.query(p, shear, *znum(case32), precision=_capi.PREC_FP16X3)
    b = ctx.query(p, shear, *znum(case32), precision=_capi.PREC_FP32)
    assert max((a[0] - b[0]).abs().max().item(), (a[1] - b[1]).abs().max().item()) < TOL_X3_MAX


def test_refined_mode_reproduces_the_split_operand_mesh(ctx, case32):
    """SURS_PREC_FP16R: one pass everywhere + split operands on the nodes the 0.5 iso-surface can depend on.  Every
    inside / outside bit equals the FP16X3 volume's, every node next to a sign change carries the FP16X3 value, so
    marching cubes gives the FP16X3 mesh bit for bit; untouched nodes keep the one-pass value."""
    from surs_b200 import _capi
    for res, bmax in (((64, 64, 64), [0.5, 0.5, 0.5]), ((40, 50, 128), [0.5, 0.4, 0.55])):
        args = (res, [-0.5] * 3, bmax, case32.calib) + znum(case32)
        x3 = ctx.eval_grid(*args, precision=_capi.PREC_FP16X3)
        one = ctx.eval_grid(*args, precision=_capi.PREC_FP16)
        ref = ctx.eval_grid(*args, precision=_capi.PREC_FP16R)
        n_ref = ctx.refined_nodes
        assert 0 < n_ref < 0.6 * res[0] * res[1] * res[2]
        changed = 0
        for r, x, o in zip(ref, x3, one):
            assert torch.equal(r > 0.5, x > 0.5)
            same_as_x3, same_as_one = r == x, r == o
            assert bool((same_as_x3 | same_as_one).all())
            changed += int((~same_as_one).sum())
            # nodes with an inside / outside change to a 6-neighbour must carry the split-operand value
            b = x > 0.5
            edge = torch.zeros_like(b)
            for d in range(3):
                diff = b.narrow(d, 1, b.shape[d] - 1) != b.narrow(d, 0, b.shape[d] - 1)
                edge.narrow(d, 1, b.shape[d] - 1).logical_or_(diff)
                edge.narrow(d, 0, b.shape[d] - 1).logical_or_(diff)
            assert bool(same_as_x3[edge].all())
            vr, _, fr, _, _, _ = ctx.marching_cubes(r, 0.5)
            vx, _, fx, _, _, _ = ctx.marching_cubes(x, 0.5)
            assert torch.equal(fr, fx) and torch.equal(vr, vx)
        assert changed <= 2 * n_ref
        print("refined mode %s: %d of %d nodes re-evaluated (%.1f %%)" % (res, n_ref, res[0] * res[1] * res[2], 100.0 * n_ref / (res[0] * res[1] * res[2])))
    # point sources and the octree: identical to FP16X3
    p = torch.from_numpy(syn.random_points(700, seed=5)).to(ctx.device)
    a = ctx.query(p, case32.calib, *znum(case32), precision=_capi.PREC_FP16R)
    b = ctx.query(p, case32.calib, *znum(case32), precision=_capi.PREC_FP16X3)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    args = ((64, 64, 64), [-0.5] * 3, [0.5] * 3, case32.calib) + znum(case32)
    oa = ctx.eval_grid_octree(*args, threshold=0.05, init_resolution=16, precision=_capi.PREC_FP16R)
    ob = ctx.eval_grid_octree(*args, threshold=0.05, init_resolution=16, precision=_capi.PREC_FP16X3)
    assert torch.equal(oa[0], ob[0]) and torch.equal(oa[1], ob[1]) and oa[2] == ob[2]


def test_refined_mode_widens_the_band_or_falls_back_when_the_check_fails(ctx, case32, monkeypatch):
    """The band of SURS_PREC_FP16R is verified on every call.  When max |one-pass - split| over the re-evaluated values
    is not safely inside it (driven here by a tiny band) the band is widened once and the selection repeated -- the mesh
    guarantee then holds again; without retries the whole slab is re-evaluated with split operands and the result IS
    the FP16X3 volume.  The statistics say which happened."""
    from surs_b200 import _capi
    args = ((48, 40, 64), [-0.5] * 3, [0.5, 0.4, 0.5], case32.calib) + znum(case32)
    x3 = ctx.eval_grid(*args, precision=_capi.PREC_FP16X3)
    monkeypatch.setenv("SURS_REFINE_BAND", "1e-4")
    r = ctx.eval_grid(*args, precision=_capi.PREC_FP16R)
    st = ctx.refine_stats
    assert st["attempts"] == 2 and not st["fell_back"] and st["band"] > 1e-3 and st["max_diff"] < 0.8 * st["band"]
    for a, x in zip(r, x3):
        assert torch.equal(a > 0.5, x > 0.5)
        edge = helpers.mc_read_mask(x > 0.5)
        assert torch.equal(a[edge], x[edge])
        va, _, fa, _, _, _ = ctx.marching_cubes(a, 0.5)
        vx, _, fx, _, _, _ = ctx.marching_cubes(x, 0.5)
        assert torch.equal(fa, fx) and torch.equal(va, vx)
    monkeypatch.setenv("SURS_REFINE_RETRIES", "0")
    r = ctx.eval_grid(*args, precision=_capi.PREC_FP16R)
    st = ctx.refine_stats
    assert st["fell_back"] and st["attempts"] == 1 and st["max_diff"] >= 0.8 * st["band"] and abs(st["band"] - 1e-4) < 1e-9
    assert torch.equal(r[0], x3[0]) and torch.equal(r[1], x3[1])
    monkeypatch.delenv("SURS_REFINE_BAND")
    monkeypatch.delenv("SURS_REFINE_RETRIES")
    r = ctx.eval_grid(*args, precision=_capi.PREC_FP16R)
    st = ctx.refine_stats
    assert not st["fell_back"] and st["attempts"] == 1 and 0 < st["nodes_lr_mlp_only"] < st["nodes"]


@pytest.mark.parametrize("seed,gain", [(3, 6.0), (4, 7.5), (5, 3.0), (6, 8.5)])
def test_default_precision_meets_the_tolerance_on_other_weights(seed, gain):
    """THE tolerance for the default precision must not depend on the one synthetic checkpoint the band was measured
    on: sharper (gain 7.5, 8.5: logits of +-60) and flatter (gain 3) MLPs, other features -- whatever the run-time band check decides
    (first band, widened band, or split operands everywhere), every value marching cubes reads is within 1e-3 of the
    fp32 mode and no inside / outside bit outside the 1e-3 band differs."""
    from surs_b200 import _capi
    c = _capi.Context("cuda:0")
    try:
        case = syn.SyntheticCase(S=64, seed=seed, gain=gain)
        load_case(c, case)
        args = ((64, 48, 128), [-0.5] * 3, [0.5, 0.25, 0.5], case.calib) + znum(case)
        got = c.eval_grid(*args)                                  # default precision
        st = c.refine_stats
        ref = c.eval_grid(*args, precision=_capi.PREC_FP32)
        one = c.eval_grid(*args, precision=_capi.PREC_FP16)
        print("seed %d gain %g: one-pass max err %.3g; refinement %s" % (seed, gain, max(float((one[0] - ref[0]).abs().max()), float((one[1] - ref[1]).abs().max())), st))
        for g, r, name in ((got[0], ref[0], "HR"), (got[1], ref[1], "LR")):
            helpers.parity_report(g, r, mask=helpers.mc_read_mask(r > 0.5), label="default precision, seed %d gain %g, %s" % (seed, gain, name))
    finally:
        c.close()


def test_feature_stripe_upload_for_slabs(ctx, case32):
    """surs_set_features_host with a u range uploads only the pixel columns a slab samples: the slab is bit-identical
    to the one computed from whole maps, and calls that would sample outside the stripe are refused."""
    from surs_b200 import _capi, parallel
    res, bmin, bmax = (64, 64, 64), [-0.5] * 3, [0.5] * 3
    args = (res, bmin, bmax, case32.calib) + znum(case32)
    full = {p: ctx.eval_grid(*args, precision=p) for p in (_capi.PREC_FP16, _capi.PREC_FP16X3, _capi.PREC_FP32)}
    f_lr, f_hr = torch.from_numpy(case32.feat_lr).pin_memory(), torch.from_numpy(case32.feat_hr).pin_memory()
    try:
        for lo, hi in ((0, 9), (20, 41), (55, 64)):
            # poison the resident maps, then upload the stripe only
            ctx.set_features(torch.full_like(f_lr, 7.0).to(ctx.device), torch.full_like(f_hr, -7.0).to(ctx.device))
            ctx.set_features_host(f_lr, f_hr, u_range=parallel.slab_u_range(res, bmin, bmax, case32.calib, lo, hi))
            for p, vols in full.items():
                s_hr, s_lr = ctx.eval_grid(*args, precision=p, plane_lo=lo, plane_hi=hi)
                assert torch.equal(s_hr, vols[0][lo:hi]) and torch.equal(s_lr, vols[1][lo:hi])
            if lo > 0:
                with pytest.raises(RuntimeError, match="stripe"):
                    ctx.eval_grid(*args, plane_lo=lo - 3, plane_hi=hi)
            with pytest.raises(RuntimeError, match="stripe"):
                ctx.query(torch.zeros(3, 4, device=ctx.device), case32.calib, *znum(case32))
            with pytest.raises(RuntimeError, match="stripe"):
                ctx.eval_grid_octree(*args, threshold=0.05, init_resolution=16)
        ctx.set_features_host(f_lr, f_hr)                    # whole maps from the host
        a = ctx.eval_grid(*args, precision=_capi.PREC_FP16)
        assert torch.equal(a[0], full[_capi.PREC_FP16][0])
    finally:
        load_case(ctx, case32)


def test_incremental_layer1_dense_path(ctx, case32, monkeypatch):
    """SURS_COL_INC=1: layer 1 updated incrementally along the column (query_inc.cu) instead of as a
    GEMM -- same occupancies within the fp16 tolerance, deterministic, slabs bit identical."""
    from surs_b200 import _capi
    for res, bmax in (((6, 7, 300), [0.5, 0.4, 0.55]), ((4, 33, 128), [0.2, 0.5, 0.5]), ((2, 3, 64), [0.5, 0.5, 0.5])):
        args = (res, [-0.5] * 3, bmax, case32.calib) + znum(case32)
        gemm = ctx.eval_grid(*args, precision=_capi.PREC_FP16)
        ref = ctx.eval_grid(*args, precision=_capi.PREC_FP32)
        monkeypatch.setenv("SURS_COL_INC", "1")
        inc = ctx.eval_grid(*args, precision=_capi.PREC_FP16)
        inc2 = ctx.eval_grid(*args, precision=_capi.PREC_FP16)
        slab = ctx.eval_grid(*args, precision=_capi.PREC_FP16, plane_lo=1, plane_hi=2)
        monkeypatch.delenv("SURS_COL_INC")
        for a, a2, b, g, sl in zip(inc, inc2, ref, gemm, slab):
            d = (a - b).abs()
            print("incremental path %s: max|d| vs fp32 %.3g mean %.3g; vs gemm kernel %.3g" %
                  (res, d.max().item(), d.mean().item(), (a - g).abs().max().item()))
            assert d.max().item() < TOL_FP16_MAX and d.mean().item() < TOL_FP16_MEAN
            assert torch.equal(a, a2) and torch.equal(a[1:2], sl)
            assert np.array_equal((a == 0).cpu().numpy(), (b == 0).cpu().numpy())


def test_cta_pair_dense_kernel_is_bit_identical(ctx, case32, monkeypatch, capfd):
    """SURS_COL_PAIR=1: the one-pass dense kernel on CTA pairs (query_col2.cu: tcgen05.mma.cta_group::2, each CTA holds
    half of every weight block) -- same K order and accumulation, so the volumes are bit identical to the one-CTA
    kernel, for even and odd tile counts, ragged columns and slabs; also under the refined default precision."""
    from surs_b200 import _capi
    for res, bmax in (((64, 64, 64), [0.5, 0.5, 0.5]), ((3, 5, 128), [0.5, 0.4, 0.55]), ((5, 9, 200), [0.5, 0.4, 0.55]), ((1, 1, 70), [0.2, 0.5, 0.5])):
        args = (res, [-0.5] * 3, bmax, case32.calib) + znum(case32)
        one = ctx.eval_grid(*args, precision=_capi.PREC_FP16)
        ref = ctx.eval_grid(*args, precision=_capi.PREC_FP16R)
        monkeypatch.setenv("SURS_COL_PAIR", "1")
        pair = ctx.eval_grid(*args, precision=_capi.PREC_FP16)
        pair_ref = ctx.eval_grid(*args, precision=_capi.PREC_FP16R)
        slab = ctx.eval_grid(*args, precision=_capi.PREC_FP16, plane_lo=res[0] // 2, plane_hi=res[0])
        monkeypatch.delenv("SURS_COL_PAIR")
        for a, b, c, d, sl in zip(one, pair, ref, pair_ref, slab):
            assert torch.equal(a, b), res
            assert torch.equal(c, d), res
            assert torch.equal(a[res[0] // 2:], sl), res
    # the pair kernel really is what ran: its profiling switch reports from inside query_col2.cu
    monkeypatch.setenv("SURS_COL_PAIR", "1")
    monkeypatch.setenv("SURS_COL_ABLATE", "128")
    capfd.readouterr()
    ctx.eval_grid((4, 4, 128), [-0.5] * 3, [0.5] * 3, case32.calib, *znum(case32), precision=_capi.PREC_FP16)
    monkeypatch.delenv("SURS_COL_ABLATE")
    monkeypatch.delenv("SURS_COL_PAIR")
    assert "[surs pair profile] pairs=8" in capfd.readouterr().err


def test_tensor_pipe_rate_probe(ctx):
    """scripts/umma_rate.py: one 128 x 256 x 16 MMA takes 128 cycles per SM with cta_group::1 and with cta_group::2
    (each CTA holding half of B), with 2 CTAs and with all SMs busy."""
    from surs_b200 import _capi
    for pair in (0, 1):
        for grid in (2, 148):
            c = _capi.selftest_umma_rate(ctx, pair, grid, 1024)
            assert len(c) == (grid // 2 if pair else grid)
            assert 127.0 < c.min() and c.max() < 140.0, (pair, grid, c.min(), c.max())


def test_marching_cubes_bitmask_word_boundaries(ctx):
    """Fast path (last axis % 4 == 0): rows that are not a multiple of 32 bits, surfaces crossing word
    boundaries, a single word per row, and the last cell of a row."""
    rng = np.random.default_rng(5)
    for shape in ((6, 7, 36), (5, 9, 100), (4, 4, 4), (3, 5, 64), (7, 3, 132)):
        vol = rng.random(shape).astype(np.float32)
        _mc_check(ctx, vol)
        ramp = np.broadcast_to(np.arange(shape[2], dtype=np.float32) / shape[2], shape).copy()      # one crossing per row ...
        ramp += 0.3 * rng.random(shape[:2]).astype(np.float32)[:, :, None]                           # ... at a row-dependent k
        _mc_check(ctx, ramp)
    edge = np.zeros((3, 3, 64), np.float32)
    edge[:, :, 31:33] = 1.0                                             # crossings exactly at bits 30|31 and 32|33
    edge[:, :, 63] = 1.0                                                # and at the last node of the row
    _mc_check(ctx, edge)


def test_transformed_and_short_grids_use_the_generic_kernel(ctx, case32):
    """A grid transform (lib/sdf.py:10-27) or a short last axis rules out the column kernels;
    both modes must still agree with the oracle on the same nodes."""
    from surs_b200 import _capi
    rot = np.eye(4)
    a = 0.3
    rot[:3, :3] = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]]) * 0.9
    rot[:3, 3] = [0.02, -0.03, 0.01]
    for res, transform in (((12, 10, 16), None), ((9, 8, 70), rot)):
        bmin, bmax = np.array([-0.5, -0.45, -0.4]), np.array([0.5, 0.5, 0.45])
        coords, _ = O.create_grid(*res, bmin, bmax, transform=transform)
        pts = coords.reshape(3, -1).astype(np.float32)
        ohr, olr = O.query(pts, case32.calib, case32.feat_lr, case32.feat_hr, case32.mlp_lr, case32.mlp_hr, load_size=case32.load_size)
        for prec, tol in ((_capi.PREC_FP32, TOL_FP32), (_capi.PREC_FP16, TOL_FP16_MAX)):
            hr, lr = ctx.eval_grid(res, bmin, bmax, case32.calib, *znum(case32), precision=prec, transform=transform)
            assert np.abs(hr.cpu().numpy().reshape(-1) - ohr).max() < tol
            assert np.abs(lr.cpu().numpy().reshape(-1) - olr).max() < tol


def test_non_square_feature_maps(ctx):
    """grid_sample normalises x by W - 1 and y by H - 1 (lib/geometry.py:4-12): maps with H != W."""
    from surs_b200 import _capi
    case = syn.SyntheticCase(S=32, seed=3)
    rng = np.random.default_rng(9)
    f_lr = rng.standard_normal((256, 6, 11)).astype(np.float32)
    f_hr = rng.standard_normal((64, 20, 13)).astype(np.float32)
    c2 = _capi.Context(ctx.device)
    try:
        t = lambda a: torch.from_numpy(a).to(c2.device)
        c2.set_weights([t(w) for w in case.mlp_lr[0]], [t(b) for b in case.mlp_lr[1]], [t(w) for w in case.mlp_hr[0]], [t(b) for b in case.mlp_hr[1]],
                       syn.MLP_DIM_LR, syn.MLP_DIM_HR, syn.RES_LAYERS)
        c2.set_features(t(f_lr), t(f_hr))
        pts = syn.random_points(3000, seed=4, lo=-0.6, hi=0.6)
        ohr, olr = O.query(pts, case.calib, f_lr, f_hr, case.mlp_lr, case.mlp_hr, load_size=case.load_size)
        for prec, tol in ((_capi.PREC_FP32, TOL_FP32), (_capi.PREC_FP16, TOL_FP16_MAX)):
            hr, lr = c2.query(t(pts), case.calib, *znum(case), precision=prec)
            assert np.abs(hr.cpu().numpy() - ohr).max() < tol and np.abs(lr.cpu().numpy() - olr).max() < tol
        col = c2.eval_grid((4, 5, 64), [-0.5] * 3, [0.5] * 3, case.calib, *znum(case), precision=_capi.PREC_FP16)
        ref = c2.eval_grid((4, 5, 64), [-0.5] * 3, [0.5] * 3, case.calib, *znum(case), precision=_capi.PREC_FP32)
        assert (col[0] - ref[0]).abs().max().item() < TOL_FP16_MAX
    finally:
        c2.close()


def test_error_paths_report_instead_of_falling_back(case32):
    """Unsupported configurations fail loudly with a message (no silent CPU / torch fallback in the C ABI)."""
    from surs_b200 import _capi
    c2 = _capi.Context("cuda:0")
    try:
        t = lambda a: torch.from_numpy(a).to(c2.device)
        pts = t(syn.random_points(64, seed=1))
        with pytest.raises(RuntimeError, match="set_weights|weights"):
            c2.query(pts, case32.calib, *znum(case32))
        bad = list(syn.MLP_DIM_LR)
        bad[1] = 512
        w_lr = [t(w) for w in case32.mlp_lr[0]]
        with pytest.raises(RuntimeError, match="mlp_dim"):
            c2.set_weights(w_lr, [t(b) for b in case32.mlp_lr[1]], [t(w) for w in case32.mlp_hr[0]], [t(b) for b in case32.mlp_hr[1]],
                           bad, syn.MLP_DIM_HR, syn.RES_LAYERS)
        with pytest.raises(RuntimeError, match="res_layers"):
            c2.set_weights(w_lr, [t(b) for b in case32.mlp_lr[1]], [t(w) for w in case32.mlp_hr[0]], [t(b) for b in case32.mlp_hr[1]],
                           syn.MLP_DIM_LR, syn.MLP_DIM_HR, [1, 2, 3])
        with pytest.raises(RuntimeError, match="2 nodes"):
            c2.mc_count(torch.zeros((1, 8, 8), device=c2.device), 0.5)
    finally:
        c2.close()


def test_perspective_and_image_space_transforms(ctx, case32, golden_dir):
    """perspective projection (lib/geometry.py:34-48) and the image-space `transforms` of query_mr /
    query_sr (:27-30) inside every kernel's projection prologue.  Perspective against vectors recorded from the
    unmodified reference; transforms against the oracle (the reference's own branch cannot run)."""
    from surs_b200 import _capi
    g = np.load(os.path.join(golden_dir, "variants_golden.npz"))
    pts_np = g["points"]
    pts = torch.from_numpy(pts_np).to(ctx.device)
    try:
        ctx.set_projection(True, None)
        for prec, tol in ((_capi.PREC_FP32, TOL_FP32), (_capi.PREC_DEFAULT, TOL), (_capi.PREC_FP16, TOL_FP16_MAX)):
            hr, lr = ctx.query(pts, g["persp_calib"], *znum(case32), precision=prec)
            assert np.abs(hr.cpu().numpy() - g["persp_hr"]).max() < tol and np.abs(lr.cpu().numpy() - g["persp_lr"]).max() < tol
            assert np.array_equal(hr.cpu().numpy() == 0, g["persp_hr"] == 0)
        # a perspective grid: generic kernels, equal to explicit points
        coords, _ = O.create_grid(8, 6, 64, np.array([-0.5] * 3), np.array([0.5] * 3))
        gp = torch.from_numpy(coords.reshape(3, -1).astype(np.float32)).to(ctx.device)
        a = ctx.query(gp, g["persp_calib"], *znum(case32), precision=_capi.PREC_FP32)
        b = ctx.eval_grid((8, 6, 64), [-0.5] * 3, [0.5] * 3, g["persp_calib"], *znum(case32), precision=_capi.PREC_FP32)
        assert torch.equal(a[0], b[0].reshape(-1)) and torch.equal(a[1], b[1].reshape(-1))
        # image-space affine on (u, v), orthogonal and perspective
        T = np.array([[0.9, 0.05, 0.02], [-0.04, 1.1, -0.03]], np.float32)
        for persp, calib in ((False, case32.calib), (True, g["persp_calib"])):
            ctx.set_projection(persp, T)
            ohr, olr = O.query(pts_np, calib, case32.feat_lr, case32.feat_hr, case32.mlp_lr, case32.mlp_hr, load_size=case32.load_size,
                               perspective=persp, uv_transform=T)
            for prec, tol in ((_capi.PREC_FP32, TOL_FP32), (_capi.PREC_DEFAULT, TOL)):
                hr, lr = ctx.query(pts, calib, *znum(case32), precision=prec)
                assert np.abs(hr.cpu().numpy() - ohr).max() < tol and np.abs(lr.cpu().numpy() - olr).max() < tol
        # the column-factored dense kernels with an image-space transform (orthogonal): node for node the generic result
        ctx.set_projection(False, T)
        coords, _ = O.create_grid(5, 6, 64, np.array([-0.5] * 3), np.array([0.5] * 3))
        gp = torch.from_numpy(coords.reshape(3, -1).astype(np.float32)).to(ctx.device)
        ref = ctx.query(gp, case32.calib, *znum(case32), precision=_capi.PREC_FP32)
        for prec, tol in ((_capi.PREC_FP16X3, TOL_X3_MAX), (_capi.PREC_FP16, TOL_FP16_MAX)):
            col = ctx.eval_grid((5, 6, 64), [-0.5] * 3, [0.5] * 3, case32.calib, *znum(case32), precision=prec)
            assert (col[0].reshape(-1) - ref[0]).abs().max().item() < tol and (col[1].reshape(-1) - ref[1]).abs().max().item() < tol
    finally:
        ctx.set_projection(False, None)
    # and the state is really reset
    hr, _ = ctx.query(pts, case32.calib, *znum(case32), precision=_capi.PREC_FP32)
    ohr, _ = O.query(pts_np, case32.calib, case32.feat_lr, case32.feat_hr, case32.mlp_lr, case32.mlp_hr, load_size=case32.load_size)
    assert np.abs(hr.cpu().numpy() - ohr).max() < TOL_FP32


def test_multi_view_query(ctx, case32, golden_dir):
    """opt.num_views = 2 (lib/model/SurfaceClassifier.py:70-76: mean over the views after layer 2) against vectors
    recorded from the unmodified reference; one and three views against the oracle."""
    g = np.load(os.path.join(golden_dir, "variants_golden.npz"))
    other = syn.SyntheticCase(S=32, seed=int(g["mv_other_seed"]))
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(ctx.device)
    pts = g["points"]
    ctx.set_features_views(t(np.stack([case32.feat_lr, other.feat_lr])), t(np.stack([case32.feat_hr, other.feat_hr])))
    hr, lr = ctx.query_views(t(np.stack([pts, pts])), g["mv_calibs"], *znum(case32))
    assert hr.shape == (2, pts.shape[1])
    assert np.abs(hr.cpu().numpy() - g["mv_hr"]).max() < TOL_FP32 and np.abs(lr.cpu().numpy() - g["mv_lr"]).max() < TOL_FP32
    assert np.array_equal(hr.cpu().numpy() == 0, g["mv_hr"] == 0)
    # three views, ragged point count, different points per view
    third = syn.SyntheticCase(S=32, seed=9)
    calibs = np.stack([g["mv_calibs"][0], g["mv_calibs"][1], g["mv_calibs"][0] * np.array([[1], [1], [-1], [1]], np.float32)])
    p3 = np.stack([syn.random_points(333, seed=s, lo=-0.55, hi=0.55) for s in (1, 2, 3)])
    fl, fh = [case32.feat_lr, other.feat_lr, third.feat_lr], [case32.feat_hr, other.feat_hr, third.feat_hr]
    ctx.set_features_views(t(np.stack(fl)), t(np.stack(fh)))
    hr, lr = ctx.query_views(t(p3), calibs, *znum(case32))
    ohr, olr = O.query_views(p3, calibs, fl, fh, case32.mlp_lr, case32.mlp_hr, load_size=case32.load_size)
    assert np.abs(hr.cpu().numpy() - ohr).max() < TOL_FP32 and np.abs(lr.cpu().numpy() - olr).max() < TOL_FP32
    # one view through the multi-view kernel = the single-view fp32 kernel
    from surs_b200 import _capi
    ctx.set_features_views(t(case32.feat_lr[None]), t(case32.feat_hr[None]))
    hr, lr = ctx.query_views(t(pts[None]), case32.calib[None], *znum(case32))
    shr, slr = ctx.query(t(pts), case32.calib, *znum(case32), precision=_capi.PREC_FP32)
    assert (hr[0] - shr).abs().max().item() < 1e-6 and (lr[0] - slr).abs().max().item() < 1e-6
