"""ctypes binding of libsurs.so (include/surs.h).  PyTorch is used for device memory and
streams only; every computation below happens inside the hand-written sm_100a kernels.

The library is required: there is no CPU or eager fallback.  If it is missing or no B200 is
present the calls fail loudly."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(CSRC, "libsurs.so")

PREC_FP32 = 0       # CUDA cores, fp32: the exact mode
PREC_FP16 = 1       # one tensor-core pass, fp16 operands: explicit opt-in, does NOT meet the 1e-3 tolerance (measured 1.5e-2)
PREC_FP16X3 = 2     # split hi/lo fp16 operands on the tensor cores (three MMA passes), |d occ| < 1e-4
PREC_FP16R = 3      # DEFAULT: one pass everywhere + split operands on the nodes the 0.5 iso-surface can depend on
                    # (dense grids; identical to PREC_FP16X3 on octree / point paths)
PREC_DEFAULT = PREC_FP16R
MC_LOWER_FOREIGN = 1

_P = ctypes.c_void_p
_I64 = ctypes.c_int64
_lib = None

# name -> (restype, argtypes); kept in one place so tests can check the exported symbols
# against include/surs.h without a GPU.
SIGNATURES = {
    "surs_version": (ctypes.c_int, []),
    "surs_refined_nodes": (_I64, [_P]),
    "surs_refine_stats": (ctypes.c_int, [_P, _P, _P, _P, _P, _P, _P]),
    "surs_mc_seam_violations": (_I64, [_P]),
    "surs_create": (ctypes.c_int, [ctypes.POINTER(_P), ctypes.c_int]),
    "surs_destroy": (None, [_P]),
    "surs_last_error": (ctypes.c_char_p, [_P]),
    "surs_set_weights": (ctypes.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, ctypes.c_int, _P]),
    "surs_set_features": (ctypes.c_int, [_P, _P, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                         _P, ctypes.c_int, ctypes.c_int, ctypes.c_int, _P]),
    "surs_set_features_host": (ctypes.c_int, [_P, _P, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                         _P, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_float, _P]),
    "surs_set_projection": (ctypes.c_int, [_P, ctypes.c_int, _P]),
    "surs_set_features_views": (ctypes.c_int, [_P, ctypes.c_int, _P, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                               _P, ctypes.c_int, ctypes.c_int, ctypes.c_int, _P]),
    "surs_query_views": (ctypes.c_int, [_P, _P, _I64, _P, ctypes.c_float, ctypes.c_float, _P, _P, _P]),
    "surs_query": (ctypes.c_int, [_P, _P, _I64, _P, ctypes.c_float, ctypes.c_float, ctypes.c_int, _P, _P, _P]),
    "surs_query_host": (ctypes.c_int, [_P, _P, _I64, _P, ctypes.c_float, ctypes.c_float, ctypes.c_int, _P, _P, _P]),
    "surs_eval_grid": (ctypes.c_int, [_P, _P, _P, _P, _P, _P, ctypes.c_float, ctypes.c_float, ctypes.c_int,
                                      ctypes.c_int, ctypes.c_int, _P, _P, _P]),
    "surs_eval_grid_octree": (ctypes.c_int, [_P, _P, _P, _P, _P, _P, ctypes.c_float, ctypes.c_float, ctypes.c_int,
                                             ctypes.c_int, ctypes.c_double, _P, _P, _P, _P]),
    "surs_octree_stats": (ctypes.c_int, [_P, _P, _P]),
    "surs_octree_select": (ctypes.c_int, [_P, _P, ctypes.c_int, _P, _P, _P, _P]),
    "surs_octree_cells": (ctypes.c_int, [_P, _P, ctypes.c_int, ctypes.c_double, _P, _P, _P, _P]),
    "surs_mc_count": (ctypes.c_int, [_P, _P, _P, ctypes.c_float, ctypes.c_int, _P, _P, _P, _P]),
    "surs_mc_interior_stats": (ctypes.c_int, [_P, _P, _P]),
    "surs_mc_count_f64": (ctypes.c_int, [_P, _P, _P, _P, ctypes.c_float, ctypes.c_int, _P, _P, _P, _P]),
    "surs_mc_value_range": (ctypes.c_int, [_P, _P, _P]),
    "surs_mc_emit": (ctypes.c_int, [_P, _P, _P, _P, _P, _P, _P, _P]),
    "surs_mc_emit_verts": (ctypes.c_int, [_P, _P, _P, _P, _P, _P, _I64, ctypes.c_int, _P, _P]),
    "surs_mc_emit_faces": (ctypes.c_int, [_P, _P, _P, _P]),
    "surs_arena_create": (ctypes.c_int, [_P, _I64, _P, _P]),
    "surs_arena_open": (ctypes.c_int, [_P, _P, _P]),
    "surs_arena_close": (ctypes.c_int, [_P, _P]),
    "surs_arena_destroy": (ctypes.c_int, [_P, _P]),
    "surs_cast_f64_f32": (ctypes.c_int, [_P, _P, _P, _I64, _P]),
    "surs_save_obj_mesh": (ctypes.c_int, [ctypes.c_char_p, _P, _I64, _P, _I64]),
    "surs_selftest_umma": (ctypes.c_int, [_P, _P, _P, ctypes.c_int, ctypes.c_int, ctypes.c_int, _P, _P]),
    "surs_selftest_umma2": (ctypes.c_int, [_P, _P, _P, ctypes.c_int, ctypes.c_int, _P, _P]),
    "surs_selftest_umma_rate": (ctypes.c_int, [_P, ctypes.c_int, ctypes.c_int, ctypes.c_int, _P, _P]),
    "surs_launch_count": (_I64, [_P]),
}


def build(verbose=False):
    """Compiles csrc/*.cu for sm_100a with nvcc (cross-compiles without a GPU)."""
    out = None if verbose else subprocess.DEVNULL
    subprocess.check_call(["make", "-C", CSRC, "-j8", "libsurs.so"], stdout=out)
    return LIB_PATH


def load():
    """Loads libsurs.so; raises (never falls back) if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    # SURS_LIB: another build of the same library (A/B timing of kernel variants on one box: scripts/ab_grid.sh)
    path = os.environ.get("SURS_LIB") or LIB_PATH
    if not os.path.exists(path):
        raise RuntimeError(
            "libsurs.so not found at %s: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  There is no CPU fallback." % path)
    lib = ctypes.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def _stream(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _ptr(t):
    if t is None:
        return None
    return ctypes.c_void_p(t.data_ptr())


def _calib12(calib):
    """Upper 3x4 of a [4,4] / [3,4] / [1,4,4] calibration as a host float32[12]."""
    c = calib.detach().to("cpu", torch.float32).numpy() if isinstance(calib, torch.Tensor) else np.asarray(calib, np.float32)
    c = c.reshape(-1, c.shape[-1])[:3, :4]
    return np.ascontiguousarray(c, dtype=np.float32)


class Context:
    """One libsurs context = one device.  Holds the packed MLP weights and feature maps."""

    def __init__(self, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("surs_b200 needs a CUDA device (B200, sm_100a); none is visible")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.lib = load()
        h = _P()
        if self.lib.surs_create(ctypes.byref(h), self.device.index):
            raise RuntimeError(self.lib.surs_last_error(None).decode())
        self._h = h
        self._keep = []
        self.feature_generation = 0     # bumped by every set_features* call (lib.model.SuRSNet re-uploads when it lags)

    def close(self):
        if getattr(self, "_h", None):
            self.lib.surs_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc:
            raise RuntimeError(self.lib.surs_last_error(self._h).decode())

    @property
    def launches(self):
        return int(self.lib.surs_launch_count(self._h))

    # ---- parameters ---------------------------------------------------------------
    def set_weights(self, w_lr, b_lr, w_hr, b_hr, dims_lr, dims_hr, res_layers):
        """w_*: 5 tensors [Cout,Cin] or [Cout,Cin,1] (conv{l}.weight); b_*: 5 tensors [Cout]."""
        def prep(ts):
            out = [t.detach().to(self.device, torch.float32).reshape(t.shape[0], -1).contiguous() for t in ts]
            return out, (_P * 5)(*[t.data_ptr() for t in out])
        kw, pw = prep(w_lr)
        kb, pb = prep(b_lr)
        kw2, pw2 = prep(w_hr)
        kb2, pb2 = prep(b_hr)
        dl = (ctypes.c_int * 6)(*[int(v) for v in dims_lr])
        dh = (ctypes.c_int * 6)(*[int(v) for v in dims_hr])
        rl = (ctypes.c_int * len(res_layers))(*[int(v) for v in res_layers])
        with torch.cuda.device(self.device):
            self._check(self.lib.surs_set_weights(self._h, pw, pb, pw2, pb2, dl, dh, rl, len(res_layers), _stream(self.device)))
        del kw, kb, kw2, kb2      # the library made its own packed copies (the call is synchronous)

    def set_features(self, f_lr, f_hr):
        """f_lr [1,256,H,W] / [256,H,W], f_hr [1,64,H,W] / [64,H,W] fp32 NCHW on the device."""
        f_lr = f_lr.detach().to(self.device, torch.float32).contiguous()
        f_hr = f_hr.detach().to(self.device, torch.float32).contiguous()
        if f_lr.dim() == 4:
            if f_lr.shape[0] != 1 or f_hr.shape[0] != 1:
                raise RuntimeError("surs_b200 kernels handle one view (num_views == 1)")
            f_lr, f_hr = f_lr[0], f_hr[0]
        with torch.cuda.device(self.device):
            self._check(self.lib.surs_set_features(self._h, _ptr(f_lr), f_lr.shape[0], f_lr.shape[1], f_lr.shape[2],
                                                   _ptr(f_hr), f_hr.shape[0], f_hr.shape[1], f_hr.shape[2], _stream(self.device)))
            torch.cuda.current_stream(self.device).synchronize()   # inputs are borrowed until the repack ran
        self.feature_generation += 1

    # ---- projection variant / multi-view ------------------------------------------------
    def set_projection(self, perspective=False, uv_transform=None):
        """Projection used by the following query / grid calls: orthogonal (lib/geometry.py:15-31) or perspective
        (:34-48), plus the optional image-space `transforms` [2,3] (or [3,3]) of query_mr / query_sr."""
        tf = None
        if uv_transform is not None:
            t = uv_transform.detach().to("cpu", torch.float32).numpy() if isinstance(uv_transform, torch.Tensor) else np.asarray(uv_transform, np.float32)
            tf = np.ascontiguousarray(t.reshape(-1, t.shape[-1])[:2, :3], dtype=np.float32)
        self._check(self.lib.surs_set_projection(self._h, int(bool(perspective)), None if tf is None else tf.ctypes.data))

    def set_features_views(self, f_lr, f_hr):
        """Multi-view maps: f_lr [V,256,H,W], f_hr [V,64,H,W] fp32 NCHW on the device."""
        f_lr = f_lr.detach().to(self.device, torch.float32).contiguous()
        f_hr = f_hr.detach().to(self.device, torch.float32).contiguous()
        if f_lr.dim() != 4 or f_hr.dim() != 4 or f_lr.shape[0] != f_hr.shape[0]:
            raise RuntimeError("set_features_views takes [V,C,H,W] maps with the same number of views")
        with torch.cuda.device(self.device):
            self._check(self.lib.surs_set_features_views(self._h, f_lr.shape[0], _ptr(f_lr), f_lr.shape[1], f_lr.shape[2], f_lr.shape[3],
                                                         _ptr(f_hr), f_hr.shape[1], f_hr.shape[2], f_hr.shape[3], _stream(self.device)))
            torch.cuda.current_stream(self.device).synchronize()
        self.n_views = int(f_lr.shape[0])

    def query_views(self, points, calibs, z_num, z_den):
        """points [V,3,N] fp32 on the device, calibs [V,4,4] / [V,3,4] -> (pred_hr, pred_lr) fp32 [V,N] (fp32 kernel)."""
        pts = points.detach().to(self.device, torch.float32).contiguous()
        V, n = pts.shape[0], pts.shape[2]
        c = calibs.detach().to("cpu", torch.float32).numpy() if isinstance(calibs, torch.Tensor) else np.asarray(calibs, np.float32)
        c = np.ascontiguousarray(c.reshape(V, -1, 4)[:, :3, :4], dtype=np.float32)
        hr = torch.empty((V, n), device=self.device, dtype=torch.float32)
        lr = torch.empty((V, n), device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            self._check(self.lib.surs_query_views(self._h, _ptr(pts), n, c.ctypes.data, float(z_num), float(z_den), _ptr(hr), _ptr(lr),
                                                  _stream(self.device)))
        return hr, lr

    # ---- query ----------------------------------------------------------------------
    @property
    def refined_nodes(self):
        """Nodes re-evaluated with split operands by the last eval_grid(precision=PREC_FP16R)."""
        return int(self.lib.surs_refined_nodes(self._h))

    @property
    def refine_stats(self):
        """Run-time band check of the last eval_grid(precision=PREC_FP16R): nodes re-evaluated, max |one-pass - split|
        over them, the band, and whether the check failed (-> the slab was re-evaluated with PREC_FP16X3)."""
        n, nl, d, b, f, a = _I64(0), _I64(0), ctypes.c_float(0), ctypes.c_float(0), ctypes.c_int(0), ctypes.c_int(0)
        self.lib.surs_refine_stats(self._h, ctypes.byref(n), ctypes.byref(nl), ctypes.byref(d), ctypes.byref(b), ctypes.byref(f), ctypes.byref(a))
        return {"nodes": int(n.value), "nodes_lr_mlp_only": int(nl.value), "max_diff": float(d.value), "band": float(b.value),
                "fell_back": bool(f.value), "attempts": int(a.value)}

    def mc_interior_stats(self):
        """(cells with an interior ambiguity, cells that took the tunnel triangulation) of the last mc_count."""
        a, b = _I64(0), _I64(0)
        self._check(self.lib.surs_mc_interior_stats(self._h, ctypes.byref(a), ctypes.byref(b)))
        return int(a.value), int(b.value)

    def mc_seam_violations(self):
        """Seam edges without a vertex id in the last mc_emit_faces (0 unless two slabs disagree on a shared plane)."""
        return int(self.lib.surs_mc_seam_violations(self._h))

    def set_features_host(self, f_lr, f_hr, u_range=None):
        """Feature maps from HOST tensors (NCHW fp32, ideally pinned).  u_range = (u_lo, u_hi): upload only the pixel
        columns that image coordinates in that range sample (slab-parallel ranks); None: the whole maps."""
        f_lr = f_lr.detach().to(torch.float32).contiguous()
        f_hr = f_hr.detach().to(torch.float32).contiguous()
        if f_lr.is_cuda or f_hr.is_cuda:
            raise RuntimeError("set_features_host takes host tensors; use set_features for device tensors")
        if f_lr.dim() == 4:
            if f_lr.shape[0] != 1 or f_hr.shape[0] != 1:
                raise RuntimeError("surs_b200 kernels handle one view (num_views == 1)")
            f_lr, f_hr = f_lr[0], f_hr[0]
        u_lo, u_hi = (-1.0, 1.0) if u_range is None else (float(u_range[0]), float(u_range[1]))

        def stripe_bytes(f):                              # same pixel-column rule as stripe_columns() in csrc/surs_api.cu
            W = f.shape[2]
            x0 = max(0, int(np.floor((min(max(u_lo, -1.0), 1.0) + 1.0) * 0.5 * (W - 1))) - 1)
            x1 = min(W, int(np.floor((min(max(u_hi, -1.0), 1.0) + 1.0) * 0.5 * (W - 1))) + 3)
            return 4 * f.shape[0] * f.shape[1] * max(0, x1 - x0)
        self.last_upload_bytes = stripe_bytes(f_lr) + stripe_bytes(f_hr)
        with torch.cuda.device(self.device):
            self._check(self.lib.surs_set_features_host(self._h, _ptr(f_lr), f_lr.shape[0], f_lr.shape[1], f_lr.shape[2],
                                                        _ptr(f_hr), f_hr.shape[0], f_hr.shape[1], f_hr.shape[2], u_lo, u_hi,
                                                        _stream(self.device)))
            torch.cuda.current_stream(self.device).synchronize()   # the host buffers are borrowed until the copies ran
        self.feature_generation += 1

    def query(self, points, calib, z_num, z_den, precision=PREC_FP16R):
        """points [3,N] fp32 on the device -> (pred_hr, pred_lr) fp32 [N] device tensors."""
        pts = points.detach().to(self.device, torch.float32).contiguous()
        n = pts.shape[1]
        hr = torch.empty(n, device=self.device, dtype=torch.float32)
        lr = torch.empty(n, device=self.device, dtype=torch.float32)
        c = _calib12(calib)
        with torch.cuda.device(self.device):
            self._check(self.lib.surs_query(self._h, _ptr(pts), n, c.ctypes.data, float(z_num), float(z_den), int(precision),
                                            _ptr(hr), _ptr(lr), _stream(self.device)))
        return hr, lr

    def query_host(self, points_np, calib, z_num, z_den, precision=PREC_FP16R, out_hr=None, out_lr=None):
        """Host in / host out ([3,N] float32 numpy or pinned tensor), copies inside."""
        pts = np.ascontiguousarray(points_np, dtype=np.float32) if not isinstance(points_np, torch.Tensor) else points_np
        n = pts.shape[1]
        hr = np.empty(n, np.float32) if out_hr is None else out_hr
        lr = np.empty(n, np.float32) if out_lr is None else out_lr
        c = _calib12(calib)
        src = pts.data_ptr() if isinstance(pts, torch.Tensor) else pts.ctypes.data
        ph = hr.data_ptr() if isinstance(hr, torch.Tensor) else hr.ctypes.data
        pl = lr.data_ptr() if isinstance(lr, torch.Tensor) else lr.ctypes.data
        with torch.cuda.device(self.device):
            self._check(self.lib.surs_query_host(self._h, src, n, c.ctypes.data, float(z_num), float(z_den), int(precision),
                                                 ph, pl, _stream(self.device)))
        return hr, lr

    # ---- grids ----------------------------------------------------------------------
    @staticmethod
    def _grid_args(res, b_min, b_max, transform):
        r = (ctypes.c_int * 3)(*[int(v) for v in res])
        bmin = np.ascontiguousarray(np.asarray(b_min, dtype=np.float64).reshape(3))
        bmax = np.ascontiguousarray(np.asarray(b_max, dtype=np.float64).reshape(3))
        tr = None if transform is None else np.ascontiguousarray(np.asarray(transform, dtype=np.float64)[:3, :4])
        return r, bmin, bmax, tr

    def eval_grid(self, res, b_min, b_max, calib, z_num, z_den, transform=None, precision=PREC_FP16R,
                  plane_lo=0, plane_hi=None):
        """Dense evaluation of planes [plane_lo, plane_hi) -> two fp32 device volumes."""
        plane_hi = res[0] if plane_hi is None else plane_hi
        r, bmin, bmax, tr = self._grid_args(res, b_min, b_max, transform)
        shape = (plane_hi - plane_lo, int(res[1]), int(res[2]))
        hr = torch.empty(shape, device=self.device, dtype=torch.float32)
        lr = torch.empty(shape, device=self.device, dtype=torch.float32)
        c = _calib12(calib)
        with torch.cuda.device(self.device):
            self._check(self.lib.surs_eval_grid(self._h, r, bmin.ctypes.data, bmax.ctypes.data,
                                                None if tr is None else tr.ctypes.data, c.ctypes.data,
                                                float(z_num), float(z_den), int(precision), int(plane_lo), int(plane_hi),
                                                _ptr(hr), _ptr(lr), _stream(self.device)))
        return hr, lr

    def eval_grid_octree(self, res, b_min, b_max, calib, z_num, z_den, threshold, init_resolution=64,
                         transform=None, precision=PREC_FP16R):
        """Octree evaluation -> two float64 device volumes + number of network evaluations."""
        r, bmin, bmax, tr = self._grid_args(res, b_min, b_max, transform)
        shape = tuple(int(v) for v in res)
        hr = torch.empty(shape, device=self.device, dtype=torch.float64)
        lr = torch.empty(shape, device=self.device, dtype=torch.float64)
        c = _calib12(calib)
        n_eval = _I64(0)
        with torch.cuda.device(self.device):
            self._check(self.lib.surs_eval_grid_octree(self._h, r, bmin.ctypes.data, bmax.ctypes.data,
                                                       None if tr is None else tr.ctypes.data, c.ctypes.data,
                                                       float(z_num), float(z_den), int(precision), int(init_resolution),
                                                       float(threshold), _ptr(hr), _ptr(lr), ctypes.byref(n_eval),
                                                       _stream(self.device)))
        return hr, lr, int(n_eval.value)

    def octree_stats(self):
        """Phase times (ms) and counts of the last eval_grid_octree (include/surs.h: surs_octree_stats)."""
        ms = (ctypes.c_float * 5)()
        cnt = (_I64 * 6)()
        self._check(self.lib.surs_octree_stats(self._h, ms, cnt))
        out = dict(zip(("init_ms", "table_ms", "select_ms", "query_ms", "cells_ms"), [float(v) for v in ms]))
        out.update(zip(("candidates", "evaluated", "cells", "cells_read", "filled_hr", "filled_lr"), [int(v) for v in cnt]))
        return out

    def octree_select(self, res, reso, dirty, idx):
        r = (ctypes.c_int * 3)(*[int(v) for v in res])
        n = _I64(0)
        with torch.cuda.device(self.device):
            self._check(self.lib.surs_octree_select(self._h, r, int(reso), _ptr(dirty), _ptr(idx), ctypes.byref(n), _stream(self.device)))
        return int(n.value)

    def octree_cells(self, res, reso, threshold, sdf_hr, sdf_lr, dirty):
        r = (ctypes.c_int * 3)(*[int(v) for v in res])
        with torch.cuda.device(self.device):
            self._check(self.lib.surs_octree_cells(self._h, r, int(reso), float(threshold), _ptr(sdf_hr), _ptr(sdf_lr), _ptr(dirty),
                                                   _stream(self.device)))

    def cast_f64_f32(self, src):
        dst = torch.empty(src.shape, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            self._check(self.lib.surs_cast_f64_f32(self._h, _ptr(src), _ptr(dst), src.numel(), _stream(self.device)))
        return dst

    # ---- marching cubes ---------------------------------------------------------------
    def mc_count(self, vol, level, flags=0):
        """vol: contiguous device volume, fp32 -- or float64 (octree volumes): the float32 copy skimage would make is
        then written by the same pass and kept as ``self._mc_vol``.  Returns (n_verts, n_faces, n_ambiguous_cells);
        ``mc_value_range()`` gives the volume's (min, max) afterwards."""
        assert vol.dtype in (torch.float32, torch.float64) and vol.is_contiguous() and vol.device == self.device
        r = (ctypes.c_int * 3)(*[int(v) for v in vol.shape])
        nv, nf, na = _I64(0), _I64(0), _I64(0)
        with torch.cuda.device(self.device):
            if vol.dtype == torch.float64:
                vol32 = torch.empty(vol.shape, device=self.device, dtype=torch.float32)
                self._check(self.lib.surs_mc_count_f64(self._h, _ptr(vol), _ptr(vol32), r, float(level), int(flags), ctypes.byref(nv),
                                                       ctypes.byref(nf), ctypes.byref(na), _stream(self.device)))
                vol = vol32
            else:
                self._check(self.lib.surs_mc_count(self._h, _ptr(vol), r, float(level), int(flags), ctypes.byref(nv), ctypes.byref(nf),
                                                   ctypes.byref(na), _stream(self.device)))
        self._mc_vol = vol            # borrowed until the emit calls ran
        return int(nv.value), int(nf.value), int(na.value)

    def mc_value_range(self):
        lo, hi = ctypes.c_float(0), ctypes.c_float(0)
        self._check(self.lib.surs_mc_value_range(self._h, ctypes.byref(lo), ctypes.byref(hi)))
        return float(lo.value), float(hi.value)

    def mc_emit_verts(self, n_verts, mat=None, vert_id_offset=0, seam_out=None, want_normals=True, plane_offset=0, out_ptrs=None):
        """out_ptrs = (world, normals, values) raw device addresses (ints; e.g. offsets into a peer arena): the kernel
        writes there instead of into freshly allocated tensors, no index-space vertices are produced, returns None."""
        m = None if mat is None else np.ascontiguousarray(np.asarray(mat, dtype=np.float64)[:3, :4])
        if out_ptrs is not None:
            world_p, normals_p, values_p = (None if p in (None, 0) else ctypes.c_void_p(int(p)) for p in out_ptrs)
            with torch.cuda.device(self.device):
                self._check(self.lib.surs_mc_emit_verts(self._h, None if m is None else m.ctypes.data, None, world_p, normals_p, values_p,
                                                        int(vert_id_offset), int(plane_offset), _ptr(seam_out), _stream(self.device)))
            return None
        verts = torch.empty((n_verts, 3), device=self.device, dtype=torch.float32)
        normals = torch.empty((n_verts, 3), device=self.device, dtype=torch.float32) if want_normals else None
        values = torch.empty((n_verts,), device=self.device, dtype=torch.float32) if want_normals else None
        world = torch.empty((n_verts, 3), device=self.device, dtype=torch.float64) if mat is not None else None
        with torch.cuda.device(self.device):
            self._check(self.lib.surs_mc_emit_verts(self._h, None if m is None else m.ctypes.data, _ptr(verts), _ptr(world),
                                                    _ptr(normals), _ptr(values), int(vert_id_offset), int(plane_offset), _ptr(seam_out),
                                                    _stream(self.device)))
        return verts, world, normals, values

    def mc_emit_faces(self, n_faces, seam_in=None, out_ptr=None):
        if out_ptr is not None:
            with torch.cuda.device(self.device):
                self._check(self.lib.surs_mc_emit_faces(self._h, ctypes.c_void_p(int(out_ptr)), _ptr(seam_in), _stream(self.device)))
            return None
        faces = torch.empty((n_faces, 3), device=self.device, dtype=torch.int32)
        with torch.cuda.device(self.device):
            self._check(self.lib.surs_mc_emit_faces(self._h, _ptr(faces), _ptr(seam_in), _stream(self.device)))
        return faces

    # ---- peer arenas (multi-GPU emission into another rank's memory) ------------------------------
    def arena_create(self, nbytes):
        """-> (device address, 64-byte CUDA IPC handle) of `nbytes` of memory on this device."""
        p, h = _P(), (ctypes.c_ubyte * 64)()
        with torch.cuda.device(self.device):
            self._check(self.lib.surs_arena_create(self._h, int(nbytes), ctypes.byref(p), h))
        return int(p.value), bytes(h)

    def arena_open(self, handle):
        p, h = _P(), (ctypes.c_ubyte * 64).from_buffer_copy(handle)
        with torch.cuda.device(self.device):
            self._check(self.lib.surs_arena_open(self._h, h, ctypes.byref(p)))
        return int(p.value)

    def arena_close(self, peer_ptr):
        self._check(self.lib.surs_arena_close(self._h, ctypes.c_void_p(int(peer_ptr))))

    def arena_destroy(self, dev_ptr):
        self._check(self.lib.surs_arena_destroy(self._h, ctypes.c_void_p(int(dev_ptr))))

    def marching_cubes(self, vol, level, mat=None):
        """Single-device marching cubes: (verts f32 [V,3], world f64 [V,3] | None, faces i32 [F,3],
        normals, values, n_ambiguous)."""
        nv, nf, na = self.mc_count(vol, level)
        verts, world, normals, values = self.mc_emit_verts(nv, mat)
        faces = self.mc_emit_faces(nf)
        return verts, world, faces, normals, values, na

    def selftest_umma(self, A, B, tail16=False):
        A = A.to(self.device, torch.float32).contiguous()
        B = B.to(self.device, torch.float32).contiguous()
        D = torch.empty((128, B.shape[0]), device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            self._check(self.lib.surs_selftest_umma(self._h, _ptr(A), _ptr(B), B.shape[0], A.shape[1], int(bool(tail16)), _ptr(D),
                                                    _stream(self.device)))
        return D


class DeviceView:
    """A typed window into raw device memory (a peer arena) for torch.as_tensor through __cuda_array_interface__."""

    def __init__(self, ptr, shape, typestr, owner=None):
        self.__cuda_array_interface__ = {"shape": tuple(int(v) for v in shape), "typestr": typestr, "data": (int(ptr), False), "version": 3,
                                         "strides": None}
        self._owner = owner


def tensor_from_ptr(ptr, shape, dtype, device, owner=None):
    typestr = {torch.float64: "<f8", torch.float32: "<f4", torch.int32: "<i4", torch.uint8: "|u1", torch.int64: "<i8"}[dtype]
    if int(np.prod(shape)) == 0:
        return torch.empty(tuple(shape), dtype=dtype, device=device)
    return torch.as_tensor(DeviceView(ptr, shape, typestr, owner), device=device)


def save_obj_mesh(mesh_path, verts, faces):
    """lib/mesh_util.py:53-61, byte-identical output, written by the C writer."""
    v = np.ascontiguousarray(np.asarray(verts, dtype=np.float64).reshape(-1, 3))
    f = np.ascontiguousarray(np.asarray(faces, dtype=np.int32).reshape(-1, 3))
    if load().surs_save_obj_mesh(os.fsencode(mesh_path), v.ctypes.data, v.shape[0], f.ctypes.data, f.shape[0]):
        raise OSError("could not write %s" % mesh_path)


def selftest_umma2(ctx, A, B):
    """D[256,N] = A[256,K] . B[N,K]^T through tcgen05.mma.cta_group::2 (unit test of the CTA-pair path)."""
    A = A.to(ctx.device, torch.float32).contiguous()
    B = B.to(ctx.device, torch.float32).contiguous()
    D = torch.empty((256, B.shape[0]), device=ctx.device, dtype=torch.float32)
    with torch.cuda.device(ctx.device):
        ctx._check(ctx.lib.surs_selftest_umma2(ctx._h, _ptr(A), _ptr(B), B.shape[0], A.shape[1], _ptr(D), _stream(ctx.device)))
    return D


def selftest_umma_rate(ctx, pair, grid, reps=1024):
    """Cycles per MMA (128 x 256 x 16, or 256 x 256 x 16 on CTA pairs) seen by every issuing thread of a `grid`-CTA launch."""
    out = torch.zeros(grid, dtype=torch.int64, device=ctx.device)
    with torch.cuda.device(ctx.device):
        ctx._check(ctx.lib.surs_selftest_umma_rate(ctx._h, int(bool(pair)), int(grid), int(reps), _ptr(out), _stream(ctx.device)))
    torch.cuda.synchronize(ctx.device)
    c = out.cpu().numpy()
    return c[c > 0] / (4.0 * reps)
