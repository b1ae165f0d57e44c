"""Marching cubes: the CPU twin (oracle/mc_oracle.c, which shares csrc/mc_tables.h with the CUDA kernels) against the
TABLE-FREE checker oracle/mc_check.py -- vertex placement, one cell per triangle, closed oriented 2-manifold, per-face
segment agreement, and the topology of every ambiguous cell (face deciders + interior / tunnel test) against a brute-force
count on the trilinear interpolant.  This is what guards the shared tables."""
import numpy as np
import pytest

import helpers
from oracle import mc_check, mc_oracle


def _mesh(vol):
    v, f, n, val, st = mc_oracle.marching_cubes_lewiner(vol, 0.5, return_stats=True)
    return v, f, st


def test_smooth_fields_pass_every_table_free_check():
    for vol in (helpers.sphere_volume(24, 8.0), helpers.sphere_volume(20, 6.5, sharp=0.4, centre=9.3)):
        v, f, st = _mesh(vol)
        rep = mc_check.check_mesh(vol, 0.5, v, f)
        assert rep["border_edges"] == 0 and rep["orientation_wrong"] == 0 and rep["max_position_error"] < 1e-5
        assert helpers.mesh_euler_closed(v, f) == 2
    # the analytic HR field of the octree goldens, cut by the box
    g = np.stack(np.meshgrid(*[np.linspace(-0.5, 0.5, 40, endpoint=False)] * 3, indexing="ij")).reshape(3, -1)
    vol = helpers.analytic_eval_func(g)[0].reshape(40, 40, 40)
    v, f, st = _mesh(vol)
    rep = mc_check.check_mesh(vol, 0.5, v, f)
    assert rep["orientation_wrong"] == 0


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_noise_volumes_topology_matches_the_trilinear_interpolant(seed):
    """White noise: a third of the cells are ambiguous, a sixth need the interior test -- every decision the tables and
    the run-time tests take must give the chamber structure of the trilinear interpolant."""
    vol = np.random.default_rng(seed).random((9, 9, 9)).astype(np.float32)
    v, f, st = _mesh(vol)
    assert st["ambiguous_cells"] > 100 and st["interior_ambiguous_cells"] > 50
    rep = mc_check.check_mesh(vol, 0.5, v, f)
    assert rep["orientation_wrong_fraction"] < 0.02        # flat triangles of a wrinkled cell, never a flipped patch
    top = mc_check.check_cell_topology(vol, 0.5, v, f, max_cells=250, seed=seed)
    print(st, {k: top[k] for k in top if k != "mismatches"})
    assert top["cells_checked"] > 150 and top["mismatch"] == 0, top["mismatches"]


def test_case_4_tunnel_is_taken_exactly_when_the_interpolant_has_one():
    """Two diagonally opposite positive corners (Lewiner's case 4): two separate caps, or a tunnel through the cell."""
    for others, want_tunnel in ((-1.0, False), (-0.15, True)):
        vol = np.full((2, 2, 2), 0.5 + others, np.float32)
        vol[0, 0, 0] = vol[1, 1, 1] = 1.5
        dc = vol.astype(np.float64) - 0.5
        truth = mc_check.trilinear_chambers(dc)
        assert truth == ((1, 1) if want_tunnel else (2, 1))
        v, f, st = _mesh(vol)
        assert st["interior_ambiguous_cells"] == 1 and st["tunnel_cells"] == int(want_tunnel)
        assert len(f) == (6 if want_tunnel else 2)
        top = mc_check.check_cell_topology(vol, 0.5, v, f)
        assert top["cells_checked"] == 1 and top["mismatch"] == 0 and top["tunnel_cells"] == int(want_tunnel)
        # the complementary configuration (two negative corners) by symmetry
        v2, f2, st2 = _mesh(1.0 - vol)
        assert st2["tunnel_cells"] == int(want_tunnel) and len(f2) == len(f)


def test_generator_reproduces_the_committed_tables(tmp_path):
    """csrc/mc_tables.h is the output of csrc/gen_mc_tables.py (product source next to the kernels)."""
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    csrc = os.path.join(root, "super-resolution-3d-human-shape-from-a-single-low-resolution-image_b200", "csrc")
    spec = importlib.util.spec_from_file_location("gen_mc_tables", os.path.join(csrc, "gen_mc_tables.py"))
    g = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(g)
    out = tmp_path / "mc_tables.h"
    nent, nidx, ntest = g.emit(str(out))
    with open(os.path.join(csrc, "mc_tables.h")) as f:
        assert out.read_text() == f.read()
    assert nent > 656 and ntest > 0
