"""CPU-side checks of the C-ABI boundary: the library loads, exports exactly what include/surs.h
declares, the host-side OBJ writer is byte-identical to the reference, and nothing in the product
package imports the oracle."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def capi():
    from surs_b200 import _capi
    if not os.path.exists(_capi.LIB_PATH):
        _capi.build()
    return _capi


def test_header_and_library_agree(capi):
    with open(os.path.join(ROOT, "include", "surs.h")) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    declared = set(re.findall(r"\b(surs_[a-z0-9_]+)\s*\(", text))
    assert declared == set(capi.SIGNATURES), declared ^ set(capi.SIGNATURES)
    lib = ctypes.CDLL(capi.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert capi.load().surs_version() >= 100


def test_create_fails_loudly_without_gpu(capi):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError):
        capi.Context()
    h = ctypes.c_void_p()
    assert capi.load().surs_create(ctypes.byref(h), 0) != 0
    assert b"surs_create" in capi.load().surs_last_error(None)


def test_obj_writer_matches_reference(capi, tmp_path, golden_dir):
    inp = np.load(os.path.join(golden_dir, "obj_golden_input.npz"))
    p = tmp_path / "m.obj"
    capi.save_obj_mesh(str(p), inp["verts"], inp["faces"])
    with open(os.path.join(golden_dir, "obj_golden.txt"), "rb") as f:
        assert p.read_bytes() == f.read()
    # a larger mesh against the oracle's restatement of lib/mesh_util.py:53-61
    from oracle import surs_oracle as O
    rng = np.random.default_rng(0)
    v = rng.standard_normal((5000, 3)) * np.array([1.0, 100.0, 1e-3])
    f = rng.integers(0, 5000, (9000, 3)).astype(np.int32)
    capi.save_obj_mesh(str(p), v, f)
    assert p.read_text() == O.obj_text(v, f)
    # the writer formats '%.4f' without printf except near rounding ties: values on and around the ties at the
    # fourth decimal, signed zeros, tiny / huge / non-finite values, several chunks of 64 K lines (all threads)
    n = 120000
    k = rng.integers(-2_000_000, 2_000_000, n).astype(np.float64)
    near = (k + 0.5) / 1e4 + rng.choice([0, 1e-12, -1e-12, 1e-9, -1e-9, 3e-7, -3e-7, 2e-6, -2e-6], n)
    vals = np.concatenate([rng.standard_normal(n) * rng.choice([1e-6, 1e-3, 1, 100, 1e4, 1e6, 1e12], n), near,
                           [0.0, -0.0, 1e5, -1e5, 99999.99995, -0.00004, -0.00005, 0.00005, np.inf, -np.inf, np.nan, 1e300, 5e-324,
                            0.12345, 0.12355, 2.5e-5]])
    v = np.concatenate([vals, np.zeros((-len(vals)) % 3)]).reshape(-1, 3)
    f = np.array([[0, 1, 2], [2147483646, 0, 5], [7, 2147483646, 1]], dtype=np.int32)
    capi.save_obj_mesh(str(p), v, f)
    assert p.read_text() == O.obj_text(v, f)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "super-resolution-3d-human-shape-from-a-single-low-resolution-image_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")) and fn != "mc_tables.h":
                with open(os.path.join(dirpath, fn)) as f:
                    src = f.read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), fn
                assert "/root/reference" not in src, fn


def test_cache_identity_helpers_are_sound_on_cpu():
    """lib/model/SuRSNet.py keys its caches on held tensors + version counters + a checksum, never on data_ptr alone."""
    import torch
    import importlib
    M = importlib.import_module("surs_b200.lib.model.SuRSNet")
    a, b = torch.zeros(4, 3), torch.ones(2)
    h = M._Held([a, b])
    assert h.same([a, b])
    assert not h.same([a.clone(), b]) and not h.same([b, a]) and not h.same([a])
    a.add_(1.0)                                   # in-place: version counter
    assert not h.same([a, b])
    h = M._Held([a, b])
    a.data = torch.zeros(4, 3)                    # storage swapped under the same object
    assert not h.same([a, b])
    w = torch.nn.Parameter(torch.randn(5, 5))
    c0 = M._checksum([w])
    v = w._version
    w.data.normal_()                              # bypasses the version counter: the checksum sees it
    assert w._version == v and M._checksum([w]) != c0
    # the net's feature lists are properties: assignment bumps a generation, PREC_FP16R is the default
    from helpers import make_opt
    net = M.SuRSNet(make_opt(), encoder=None)
    from surs_b200 import _capi
    assert net.precision == _capi.PREC_FP16R
    g = net._feat_gen
    net.im_feat_list_lr = [torch.zeros(1, 256, 4, 4)]
    net.im_feat_list_hr = [torch.zeros(1, 64, 8, 8)]
    assert net._feat_gen == g + 2 and net.im_feat_list_lr[0].shape[1] == 256
    net._w_dirty = False
    net.load_state_dict(net.state_dict())
    assert net._w_dirty
    net._w_dirty = False
    net.float()
    assert net._w_dirty
