"""Pins marching cubes against the function the reference actually calls -- skimage.measure.marching_cubes_lewiner
(lib/mesh_util.py:40,45; scikit-image 0.17.2, environment.yml:333) -- WHEREVER scikit-image is importable.

It is not installed in this image (nor in the offline wheelhouse), so here these tests skip and the marching-cubes parity
stays UNPINNED; on any machine that has scikit-image (also a driver-provided baseline/_ref install, which
is put on sys.path below) they activate by themselves.  scikit-image >= 0.19 renamed the function:
measure.marching_cubes(volume, level, method='lewiner').
"""
import os
import sys

import numpy as np
import pytest

import helpers

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_REF = os.path.join(ROOT, "baseline", "_ref")
if os.path.isdir(_REF) and _REF not in sys.path:
    sys.path.append(_REF)

measure = pytest.importorskip("skimage.measure", reason="scikit-image is not installed: marching-cubes parity vs skimage stays unpinned")


def sk_mc(vol, level):
    if hasattr(measure, "marching_cubes_lewiner"):
        return measure.marching_cubes_lewiner(vol, level)
    return measure.marching_cubes(vol, level, method="lewiner")


def _volumes():
    yield "sphere", helpers.sphere_volume(24, 8.0)
    g = np.stack(np.meshgrid(*[np.linspace(-0.5, 0.5, 40, endpoint=False)] * 3, indexing="ij")).reshape(3, -1)
    hr, lr = helpers.analytic_eval_func(g)
    yield "analytic_hr", hr.reshape(40, 40, 40)
    yield "analytic_lr", lr.reshape(40, 40, 40)
    yield "noise", np.random.default_rng(0).random((9, 9, 9)).astype(np.float32)


def _canon_faces(verts, faces):
    """Faces as a set of rotation-normalised triples of vertex POSITIONS (independent of vertex numbering)."""
    key = [tuple(np.round(v.astype(np.float64), 4)) for v in verts]
    out = set()
    for a, b, c in faces:
        t = (key[a], key[b], key[c])
        k = t.index(min(t))
        out.add(t[k:] + t[:k])
    return out


def _compare(name, vol, ours):
    v, f, n, val = sk_mc(vol, 0.5)
    ov, of = ours[0], ours[1]
    # vertex sets agree within 1e-5 (the north_star's bound for exactly matching occupancy)
    a = np.array(sorted(map(tuple, np.round(v.astype(np.float64), 4))))
    b = np.array(sorted(map(tuple, np.round(ov.astype(np.float64), 4))))
    edge_only = len(a) == len(b)
    print("%s: skimage %d verts / %d faces, ours %d / %d" % (name, len(v), len(f), len(ov), len(of)))
    if name != "noise":                                   # smooth fields: no ambiguous cell can differ
        assert edge_only and np.abs(a - b).max() <= 1e-4
        assert len(f) == len(of)
        assert np.array_equal(v, ov) and np.array_equal(f, of), "vertex numbering / face order differ from skimage"
        assert _canon_faces(v, f) == _canon_faces(ov, of)
    else:                                                 # noise: report how far the derived tables are from Lewiner's
        same = len(_canon_faces(v, f) & _canon_faces(ov, of))
        print("noise: %d of %d skimage faces reproduced exactly" % (same, len(f)))


def test_cpu_twin_matches_skimage():
    from oracle import mc_oracle
    for name, vol in _volumes():
        _compare(name, vol, mc_oracle.marching_cubes_lewiner(vol, 0.5))


@pytest.mark.gpu
def test_cuda_kernels_match_skimage():
    import torch
    from surs_b200 import _capi
    ctx = _capi.Context("cuda:0")
    for name, vol in _volumes():
        gv, _, gf, gn, gval, _ = ctx.marching_cubes(torch.from_numpy(vol).to(ctx.device), 0.5)
        _compare(name, vol, (gv.cpu().numpy(), gf.cpu().numpy()))
    ctx.close()
