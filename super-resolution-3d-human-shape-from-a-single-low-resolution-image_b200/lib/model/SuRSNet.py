"""Drop-in for the reference's ``lib/model/SuRSNet.py`` (+ ``BaseSuRSNet.py``) on the query side.

Same public surface: ``super_res``, ``filter_lr``, ``filter_hr``, ``query_mr``, ``query_sr``,
``get_preds``, attributes ``num_views``, ``preds_hr``, ``preds_lr``, ``im_feat_list_lr``,
``im_feat_list_hr`` -- plus the PIFu aliases ``filter(images)`` and ``query(points, calibs, ...)``
that are empty stubs in the reference (lib/model/BaseSuRSNet.py:50-56,66-78).

The two SurfaceClassifier MLPs keep the reference's state-dict keys
(``mlp_{lr,hr}.conv{0..4}.{weight,bias}``), so a reference checkpoint's MLP part loads unchanged.
The image encoder (SuRSSR_v3 + HGFilter, kept in PyTorch by design; ``lib/model/encoder.py`` holds a
state-dict-compatible one, built when ``encoder="auto"`` finds the encoder options in ``opt``) is
pluggable: pass ``encoder=`` (any module with ``super_res`` / ``filter_lr`` / ``filter_hr`` semantics).

Query semantics (reference lib/model/SuRSNet.py:131-187): ``query_mr`` computes the LR
prediction, ``query_sr`` appends the MASKED LR prediction as channel 321 and computes the HR one;
the fused kernel produces both in one launch, so ``query_mr`` runs it and ``query_sr`` on the
same points returns the cached HR result.  The stray ``print("z", z)`` of
lib/model/DepthNormalizer.py:17 is intentionally not reproduced.

Covered by the fused kernels as well: image-space ``transforms`` and the perspective projection
(lib/geometry.py:27-30,34-48; every kernel's projection prologue) and ``num_views > 1`` (the mean over
views after layer 2, lib/model/SurfaceClassifier.py:70-76; fp32 CUDA-core kernel).
Not covered by the kernels -> explicit torch path (a warning is issued once, never silent):
``--no_residual``, non-default ``mlp_dim`` / ``mlp_res_layers``, training mode (3 hourglass outputs).
"""
import warnings

import numpy as np
import torch
import torch.nn as nn

from .. import geometry
from ... import _capi
from .SurfaceClassifier import SurfaceClassifier

_DEFAULT_LR = [321, 1024, 512, 256, 128, 1]
_DEFAULT_HR = [322, 1024, 512, 256, 128, 1]


class _Held:
    """Identity of a set of tensors that stays sound while we hold it: the tensors themselves are kept (so the
    caching allocator cannot hand their memory, and CPython cannot hand their id, to a new tensor), together
    with the version counters in-place operations bump.  ``data_ptr`` catches ``p.data = other``."""

    def __init__(self, tensors):
        self.tensors = list(tensors)
        self.marks = [(t._version, t.data_ptr(), tuple(t.shape)) for t in self.tensors]

    def same(self, tensors):
        tensors = list(tensors)
        return (len(tensors) == len(self.tensors)
                and all(a is b for a, b in zip(tensors, self.tensors))
                and all(m == (t._version, t.data_ptr(), tuple(t.shape)) for m, t in zip(self.marks, tensors)))


def _checksum(tensors):
    """Two norms per tensor, one host read: catches writes that bypass the version counter
    (``p.data.normal_()``, ``init.normal_(m.weight.data)``).  20 small tensors -> ~0.1 ms."""
    n2 = torch._foreach_norm(tensors, 2)
    n1 = torch._foreach_norm(tensors, 1)
    return tuple(torch.stack(n2 + n1).double().tolist())


class SuRSNet(nn.Module):
    def __init__(self, opt, projection_mode="orthogonal", error_term=None, encoder="auto", precision=_capi.PREC_FP16R,
                 encoder_mode="eager"):
        """encoder_mode (built-in encoder only): "eager" = the reference's own execution (fp32 modules, whatever
        torch.backends.cudnn.allow_tf32 says); "fast" = the same modules with TF32 tensor-core convolutions (PyTorch's
        default for cuDNN, i.e. what the reference itself gets on a GPU) and the whole super_res + filter_hr + filter_lr
        forward captured in ONE CUDA graph per input shape (eval mode, CUDA only); "bf16" = "fast" with channels_last
        + bf16 autocast on top -- measured SLOWER on B200 (31 vs 20 ms at S = 512: layout conversions, GroupNorm and
        bicubic resampling dominate) and 10x less accurate (feature error 2e-2 vs 2e-3, occupancy up to 0.26 off), kept
        only so that the measurement can be repeated (scripts/encoder_bench.py)."""
        super().__init__()
        self.encoder_mode = encoder_mode
        self._enc_graphs = {}
        self.name = "surs_b200"
        self.opt = opt
        self.num_views = opt.num_views
        self.projection_mode = projection_mode
        self.projection = geometry.orthogonal if projection_mode == "orthogonal" else geometry.perspective
        self.index = geometry.index
        self.error_term = error_term if error_term is not None else nn.MSELoss()
        self.precision = precision
        self.mlp_lr = SurfaceClassifier(opt.mlp_dim_lr, opt.num_views, opt.no_residual, opt.mlp_res_layers_lr, nn.Sigmoid())
        self.mlp_hr = SurfaceClassifier(opt.mlp_dim_hr, opt.num_views, opt.no_residual, opt.mlp_res_layers_hr, nn.Sigmoid())
        # encoder: "auto" = the built-in PyTorch modules (reference names, so reference checkpoints load
        # with strict=True) when opt describes them; None = query side only; or any object with
        # super_res / filter_lr / filter_hr
        if encoder == "auto":
            encoder = "builtin" if hasattr(opt, "num_stack_lr") else None
        if encoder == "builtin":
            from .encoder import HGFilter, SuRSSR_v3, init_weights
            self.image_filter_lr = HGFilter(opt.num_stack_lr, opt.hg_depth, 256, opt.hg_dim, opt.norm, "low_res", False)
            self.image_filter_hr = HGFilter(opt.num_stack_hr, opt.hg_depth, 64, opt.hg_dim, opt.norm, "high_res", False)
            self.super_resolution = SuRSSR_v3(opt)
            init_weights(self)
            encoder = None
            self._builtin = True
        else:
            self._builtin = False
        self.encoder = encoder
        self._feat_gen = 0                 # bumped whenever im_feat_list_lr / im_feat_list_hr are (re)assigned
        self._im_feat_list_lr = []
        self._im_feat_list_hr = []
        self.intermediate_preds_list_lr = []
        self.intermediate_preds_list_hr = []
        self.im_SR = self.feature_lr = self.feature_hr = None
        self.preds_lr = self.preds_hr = None
        self.labels_lr = self.labels_hr = None
        self._ctx = None
        self._w_dirty = True               # MLP parameters must be (re)uploaded
        self._w_held = self._w_sum = None
        self._f_held = None                # feature tensors last uploaded (+ the generations they were uploaded at)
        self._f_gen_uploaded = self._ctx_gen_uploaded = -1
        self._q_held = None                # (points, calibs) of the last fused query_mr
        self._cached_hr = None
        self._warned = set()

    # ------------------------------------------------------------------ cache invalidation
    # The library holds packed copies of the MLP weights and repacked copies of the feature maps.  They are
    # refreshed when: the tensors are different objects / were modified in place (version counter), a cheap
    # checksum of the 20 MLP tensors changed (writes through ``.data``), ``load_state_dict`` / ``.to()`` /
    # ``.half()`` ran, ``im_feat_list_*`` were assigned, somebody else replaced the context's features, or the
    # user calls ``refresh()``.
    @property
    def im_feat_list_lr(self):
        return self._im_feat_list_lr

    @im_feat_list_lr.setter
    def im_feat_list_lr(self, value):
        self._im_feat_list_lr = value
        self._feat_gen += 1

    @property
    def im_feat_list_hr(self):
        return self._im_feat_list_hr

    @im_feat_list_hr.setter
    def im_feat_list_hr(self, value):
        self._im_feat_list_hr = value
        self._feat_gen += 1

    def refresh(self):
        """Forces the next query to re-upload the MLP parameters and the feature maps (needed only after writes
        the bookkeeping above cannot see, e.g. editing a feature map through ``.data``)."""
        self._w_dirty = True
        self._feat_gen += 1
        self._q_held = None

    def _apply(self, fn, *args, **kwargs):
        self._w_dirty = True
        self._enc_graphs = {}                      # captured graphs hold the old parameter storage
        return super()._apply(fn, *args, **kwargs)

    def load_state_dict(self, *args, **kwargs):
        self._w_dirty = True                       # (in-place copies: captured encoder graphs read the new values)
        return super().load_state_dict(*args, **kwargs)

    # ------------------------------------------------------------------ encoder side (PyTorch)
    def _need_encoder(self):
        if self._builtin:
            return self
        if self.encoder is None:
            raise RuntimeError("this SuRSNet was built without an image encoder; pass encoder= or set "
                               "im_feat_list_lr / im_feat_list_hr yourself")
        return self.encoder

    def super_res(self, images):
        """reference lib/model/SuRSNet.py:124-129."""
        enc = self._need_encoder()
        if enc is self and self._fast_encoder_ok(images):
            self.im_SR, self.feature_lr, self.feature_hr = self._encode_fast(images)
            return self.im_SR, self.feature_lr, self.feature_hr
        self._fast_out = None
        self.im_SR, self.feature_lr, self.feature_hr = self.super_resolution(images) if enc is self else enc.super_res(images)
        return self.im_SR, self.feature_lr, self.feature_hr

    def filter_lr(self, images):
        """reference lib/model/SuRSNet.py:101-110: keeps only the last hourglass output in eval mode."""
        enc = self._need_encoder()
        fast = getattr(self, "_fast_out", None)
        if fast is not None and images is self.feature_lr:       # already computed inside the captured graph
            self.im_feat_list_lr = [fast[0]]
            return
        self.im_feat_list_lr = self.image_filter_lr(images) if enc is self else enc.filter_lr(images)
        if not self.training:
            self.im_feat_list_lr = [self.im_feat_list_lr[-1]]

    def filter_hr(self, images):
        """reference lib/model/SuRSNet.py:112-122."""
        enc = self._need_encoder()
        fast = getattr(self, "_fast_out", None)
        if fast is not None and images is self.feature_hr:
            self.im_feat_list_hr = [fast[1]]
            return
        self.im_feat_list_hr = self.image_filter_hr(images) if enc is self else enc.filter_hr(images)
        if not self.training:
            self.im_feat_list_hr = [self.im_feat_list_hr[-1]]

    # ------------------------------------------------------------------ encoder, B200 configuration
    def _fast_encoder_ok(self, images):
        return (self.encoder_mode in ("fast", "bf16") and self._builtin and not self.training and images.is_cuda
                and not torch.is_grad_enabled())

    def _encoder_forward(self, images):
        """super_res + filter_hr + filter_lr of lib/train_util.py:57-59 in one function (what the graph captures)."""
        if self.encoder_mode == "bf16":
            with torch.autocast("cuda", dtype=torch.bfloat16):
                x = images.contiguous(memory_format=torch.channels_last)
                im_sr, feature_lr, feature_hr = self.super_resolution(x)
                f_hr = self.image_filter_hr(feature_hr)[-1]
                f_lr = self.image_filter_lr(feature_lr)[-1]
            # the consumers (libsurs repack, reference-style user code) take contiguous NCHW fp32
            return tuple(t.float().contiguous() for t in (im_sr, feature_lr, feature_hr, f_lr, f_hr))
        tf32 = torch.backends.cudnn.allow_tf32
        torch.backends.cudnn.allow_tf32 = True
        try:
            im_sr, feature_lr, feature_hr = self.super_resolution(images)
            f_hr = self.image_filter_hr(feature_hr)[-1]
            f_lr = self.image_filter_lr(feature_lr)[-1]
        finally:
            torch.backends.cudnn.allow_tf32 = tf32
        return im_sr, feature_lr, feature_hr, f_lr, f_hr

    def _encode_fast(self, images):
        key = (tuple(images.shape), images.dtype, images.device.index)
        entry = self._enc_graphs.get(key)
        if entry is None:
            if self.encoder_mode == "bf16" and not getattr(self, "_channels_last_done", False):
                for mod in (self.super_resolution, self.image_filter_lr, self.image_filter_hr):
                    mod.to(memory_format=torch.channels_last)
                self._channels_last_done = True
            static_in = images.detach().clone()
            side = torch.cuda.Stream(device=images.device)
            side.wait_stream(torch.cuda.current_stream(images.device))
            with torch.cuda.stream(side), torch.no_grad():
                for _ in range(3):                       # warm-up outside the capture (cuDNN plans, autocast caches)
                    self._encoder_forward(static_in)
            torch.cuda.current_stream(images.device).wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.no_grad(), torch.cuda.graph(graph):
                static_out = self._encoder_forward(static_in)
            entry = (graph, static_in, static_out)
            self._enc_graphs[key] = entry
        graph, static_in, static_out = entry
        static_in.copy_(images)
        graph.replay()
        # the graph's output buffers are overwritten by the next replay: hand out copies
        im_sr, feature_lr, feature_hr, f_lr, f_hr = (t.clone() for t in static_out)
        self._fast_out = (f_lr, f_hr)
        return im_sr, feature_lr, feature_hr

    def filter(self, images):
        """PIFu-style alias: the whole encoder (what gen_mesh does, lib/train_util.py:57-59)."""
        _, feature_lr, feature_hr = self.super_res(images)
        self.filter_hr(feature_hr)
        self.filter_lr(feature_lr)

    # ------------------------------------------------------------------ accelerated context
    def surs_context(self):
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("surs_b200 runs on a B200; move the network to a CUDA device (there is no CPU path)")
        if self._ctx is None or self._ctx.device != dev:
            self._ctx = _capi.Context(dev)
            self._w_dirty = True
            self._f_held = None
            self._q_held = None
        return self._ctx

    def depth_scale(self):
        """reference lib/model/DepthNormalizer.py:18: z * (loadSize // 2) / z_size."""
        return float(self.opt.loadSize // 2), float(self.opt.z_size)

    def _kernel_config_ok(self):
        return (self.projection_mode in ("orthogonal", "perspective") and not self.opt.no_residual
                and list(self.opt.mlp_dim_lr) == _DEFAULT_LR and list(self.opt.mlp_dim_hr) == _DEFAULT_HR
                and list(self.opt.mlp_res_layers_lr) == [2, 3, 4] and list(self.opt.mlp_res_layers_hr) == [2, 3, 4]
                and not self.training)

    def can_accelerate(self, calibs=None, transforms=None):
        """True when the fused CUDA path covers this call: default MLP shapes, eval mode; one view, or ``num_views``
        views of ONE subject (features / calibs / points with a leading dimension of num_views)."""
        ok = self._kernel_config_ok()
        if ok and transforms is not None and not (torch.is_tensor(transforms) and transforms.dim() == 2 and transforms.shape[0] >= 2
                                                  and transforms.shape[1] == 3):
            ok = False                                 # the reference indexes transforms[:2, :2] / [:2, 2:3]: one 2-D affine
        if ok and calibs is not None and torch.is_tensor(calibs) and calibs.dim() == 3 and calibs.shape[0] != self.num_views:
            ok = False
        if ok:
            self._sync()
        return ok

    def _sync(self):
        """Pushes MLP parameters / feature maps to the library when they changed."""
        ctx = self.surs_context()
        params = [c.weight for c in self.mlp_lr.layers()] + [c.bias for c in self.mlp_lr.layers()] + \
                 [c.weight for c in self.mlp_hr.layers()] + [c.bias for c in self.mlp_hr.layers()]
        if not self._w_dirty and (self._w_held is None or not self._w_held.same(params)):
            self._w_dirty = True
        csum = _checksum(params)
        if self._w_dirty or csum != self._w_sum:
            ctx.set_weights([c.weight for c in self.mlp_lr.layers()], [c.bias for c in self.mlp_lr.layers()],
                            [c.weight for c in self.mlp_hr.layers()], [c.bias for c in self.mlp_hr.layers()],
                            self.opt.mlp_dim_lr, self.opt.mlp_dim_hr, self.opt.mlp_res_layers_lr)
            self._w_held, self._w_sum, self._w_dirty = _Held(params), csum, False
            self._q_held = None
        if not self.im_feat_list_lr or not self.im_feat_list_hr:
            raise RuntimeError("image features missing: call filter_lr / filter_hr (or filter) before querying")
        feats = [self.im_feat_list_lr[-1], self.im_feat_list_hr[0]]
        if self.num_views > 1 and (feats[0].dim() != 4 or feats[0].shape[0] != self.num_views or feats[1].shape[0] != self.num_views):
            raise RuntimeError("num_views = %d: the feature maps must be [num_views, C, H, W] (one subject)" % self.num_views)
        if (self._f_held is None or self._f_gen_uploaded != self._feat_gen or self._ctx_gen_uploaded != ctx.feature_generation
                or not self._f_held.same(feats)):
            if self.num_views > 1:
                ctx.set_features_views(feats[0], feats[1])
            else:
                ctx.set_features(feats[0], feats[1])
            self._f_held = _Held(feats)
            self._f_gen_uploaded, self._ctx_gen_uploaded = self._feat_gen, ctx.feature_generation
            self._q_held = None

    def _warn_once(self, why):
        if why not in self._warned:
            self._warned.add(why)
            warnings.warn("surs_b200: %s is not covered by the fused CUDA query; using the torch path" % why)

    # ------------------------------------------------------------------ queries
    def _fused(self, points, calibs, transforms=None):
        ctx = self.surs_context()
        zn, zd = self.depth_scale()
        ctx.set_projection(self.projection_mode == "perspective", transforms)
        try:
            if self.num_views > 1:
                pts = points if points.dim() == 3 else points[None].expand(self.num_views, -1, -1)
                hr, lr = ctx.query_views(pts, calibs, zn, zd)
                return hr[:, None, :], lr[:, None, :]
            pts = points[0] if points.dim() == 3 else points
            hr, lr = ctx.query(pts, calibs, zn, zd, precision=self.precision)
            return hr.view(1, 1, -1), lr.view(1, 1, -1)
        finally:
            ctx.set_projection(False, None)

    def query_mr(self, points, calibs, transforms=None, labels=None):
        """reference lib/model/SuRSNet.py:131-159."""
        if labels is not None:
            self.labels_lr = labels
        if not self.can_accelerate(calibs, transforms):
            return self._query_mr_torch(points, calibs, transforms)
        hr, lr = self._fused(points, calibs, transforms)
        self.preds_lr = lr
        self.intermediate_preds_list_lr = [lr]
        self._cached_hr = hr
        held = [points, calibs] + ([transforms] if transforms is not None else [])
        self._q_held = _Held(held) if all(torch.is_tensor(t) for t in held) else None

    def query_sr(self, points, calibs, transforms=None, labels=None):
        """reference lib/model/SuRSNet.py:161-187 (requires query_mr on the same points first)."""
        if labels is not None:
            self.labels_hr = labels
        if not self.can_accelerate(calibs, transforms):
            return self._query_sr_torch(points, calibs, transforms)
        # the cached HR result is only valid for the very tensors query_mr saw (object identity while we hold them,
        # unmodified) and the weights / features it ran with (_sync above clears _q_held when either changed)
        if (self._q_held is not None and self._cached_hr is not None
                and self._q_held.same([points, calibs] + ([transforms] if transforms is not None else []))):
            hr = self._cached_hr
        else:
            # different points than query_mr saw: the reference would mix them; we recompute both consistently
            hr, _ = self._fused(points, calibs, transforms)
        self.preds_hr = hr
        self.intermediate_preds_list_hr = [hr]

    def query(self, points, calibs, transforms=None, labels=None):
        """PIFu-style alias: query_mr + query_sr; returns the HR prediction."""
        self.query_mr(points, calibs, transforms, labels)
        self.query_sr(points, calibs, transforms, labels)
        return self.preds_hr

    def get_preds(self):
        """reference lib/model/BaseSuRSNet.py:80-85 -- HR first."""
        return self.preds_hr, self.preds_lr

    # ------------------------------------------------------------------ torch path (uncovered variants)
    def _local_features(self, points, calibs, transforms):
        xyz = self.projection(points, calibs, transforms)
        xy, z = xyz[:, :2, :], xyz[:, 2:3, :]
        in_img = (xy[:, 0] >= -1.0) & (xy[:, 0] <= 1.0) & (xy[:, 1] >= -1.0) & (xy[:, 1] <= 1.0)
        zn, zd = self.depth_scale()
        z_feat = z * zn / zd
        feats = [torch.cat([self.index(f, xy), self.index(self.im_feat_list_hr[0], xy), z_feat], 1)
                 for f in self.im_feat_list_lr]
        return feats, in_img[:, None].float()

    def _query_mr_torch(self, points, calibs, transforms):
        self._warn_once("this configuration (non-default MLP shapes / --no_residual / training mode / batched subjects)")
        feats, mask = self._local_features(points, calibs, transforms)
        self.point_local_feat = feats
        self.intermediate_preds_list_lr = [mask * self.mlp_lr(f) for f in feats]
        self.preds_lr = self.intermediate_preds_list_lr[-1]

    def _query_sr_torch(self, points, calibs, transforms):
        feats, mask = self._local_features(points, calibs, transforms)
        self.intermediate_preds_list_hr = [mask * self.mlp_hr(torch.cat([f, p], 1))
                                           for f, p in zip(feats, self.intermediate_preds_list_lr)]
        self.preds_hr = self.intermediate_preds_list_hr[-1]

    def forward(self, points, images_lr, calibs, transforms=None):
        self.filter(images_lr)
        self.query_mr(points, calibs, transforms)
        self.query_sr(points, calibs, transforms)
        return self.get_preds()
