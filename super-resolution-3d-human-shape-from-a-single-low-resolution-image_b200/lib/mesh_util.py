"""Drop-in for the reference's ``lib/mesh_util.py``: ``reconstruction`` and the OBJ writers.

``reconstruction(opt, net, cuda, calib_tensor, resolution, b_min, b_max, use_octree=False,
num_samples=50000, transform=None)`` returns the reference's 8-tuple
``(verts_hr, faces_hr, normals_hr, values_hr, verts_lr, faces_lr, normals_lr, values_lr)``
(reference lib/mesh_util.py:8-49): verts float64 [V,3] in world coordinates, faces int32 [F,3],
normals float32 [V,3] and values float32 [V] in index space.  The PIFu call shape
``reconstruction(net, cuda, calib, resolution, b_min, b_max, ...)`` (no ``opt``) is accepted too.

With the accelerated network (``lib.model.SuRSNet``) nothing but the final mesh arrays leaves
the device: grid nodes are generated in-kernel, the (octree) volumes stay in HBM and marching
cubes runs on them.  With any other network object the generic path of the reference is
followed (``eval_func`` closure, numpy grids) and only marching cubes runs on the device
(scikit-image, which the reference calls, is replaced by the CUDA kernels in csrc/mc.cu).
"""
import numpy as np
import torch

from .sdf import create_grid, eval_grid, eval_grid_octree, grid_matrix
from .. import _capi

_MC_LEVEL = 0.5          # lib/mesh_util.py:40,45


def _mesh_from_volume(ctx, vol, mat):
    """Device marching cubes + world transform; mirrors skimage's error behaviour.  vol: fp32, or the octree's float64
    volume (the float32 copy skimage makes and the value range it checks are by-products of the bit pass)."""
    nv, nf, n_amb = ctx.mc_count(vol, _MC_LEVEL)
    vmin, vmax = ctx.mc_value_range()
    if not (vmin <= _MC_LEVEL <= vmax):
        raise ValueError("Surface level must be within volume data range.")
    if nv == 0 or nf == 0:
        raise RuntimeError("No surface found at the given iso value.")
    verts, world, normals, values = ctx.mc_emit_verts(nv, mat)
    faces = ctx.mc_emit_faces(nf)
    return world, faces, normals, values, n_amb


def _is_accelerated(net):
    return hasattr(net, "surs_context") and callable(getattr(net, "surs_context"))


_ARG_NAMES = ["opt", "net", "cuda", "calib_tensor", "resolution", "b_min", "b_max", "use_octree", "num_samples", "transform"]


def reconstruction(*args, **kwargs):
    """See the module docstring.  Accepts the SuRS signature (leading ``opt``) and PIFu's (no ``opt``)."""
    names = _ARG_NAMES
    if args and not hasattr(args[0], "threshold") and (hasattr(args[0], "query_mr") or hasattr(args[0], "query")):
        names = _ARG_NAMES[1:]
    p = dict(opt=None, use_octree=False, num_samples=50000, transform=None, precision=None, return_stats=False)
    p.update(zip(names, args))
    p.update(kwargs)
    if p["opt"] is None:
        p["opt"] = getattr(p["net"], "opt", None)
    return _reconstruction(**p)


def _reconstruction(opt, net, cuda, calib_tensor, resolution, b_min, b_max, use_octree, num_samples, transform,
                    precision, return_stats):
    import os, time
    _T = os.environ.get("SURS_TIMING") is not None
    _t = [time.perf_counter()]

    def _tick(label):
        if _T:
            torch.cuda.synchronize()
            now = time.perf_counter()
            print("[surs timing] %-22s %8.2f ms" % (label, (now - _t[0]) * 1e3))
            _t[0] = now
    b_min = np.asarray(b_min, dtype=np.float64).reshape(3)
    b_max = np.asarray(b_max, dtype=np.float64).reshape(3)
    stats = {}
    if _is_accelerated(net) and getattr(net, "num_views", 1) == 1 and net.can_accelerate(calib_tensor):
        ctx = net.surs_context()
        prec = net.precision if precision is None else precision
        mat = grid_matrix(resolution, b_min, b_max, transform)
        res = (resolution,) * 3
        zn, zd = net.depth_scale()
        persp = getattr(net, "projection_mode", "orthogonal") == "perspective"
        ctx.set_projection(persp, None)               # perspective grids take the generic kernels (no column structure)
        if use_octree:
            hr64, lr64, n_eval = ctx.eval_grid_octree(res, b_min, b_max, calib_tensor, zn, zd, float(opt.threshold),
                                                      init_resolution=64, transform=transform, precision=prec)
            vol_hr, vol_lr = hr64, lr64                 # float64: cast inside the marching-cubes bit pass
            stats["n_evaluated"] = n_eval
        else:
            vol_hr, vol_lr = ctx.eval_grid(res, b_min, b_max, calib_tensor, zn, zd, transform=transform, precision=prec)
            stats["n_evaluated"] = int(resolution) ** 3
        ctx.set_projection(False, None)
    else:
        # generic path of the reference (any net with query_mr / query_sr / get_preds)
        coords, mat = create_grid(resolution, resolution, resolution, b_min, b_max, transform=transform)

        def eval_func(points):
            points = np.expand_dims(points, axis=0)
            points = np.repeat(points, net.num_views, axis=0)
            samples = torch.from_numpy(points).to(device=cuda).float()
            net.query_mr(samples, calib_tensor)
            net.query_sr(samples, calib_tensor)
            pred_hr, pred_lr = net.get_preds()
            return pred_hr.detach().cpu().numpy(), pred_lr.detach().cpu().numpy()

        if use_octree:
            sdf_hr, sdf_lr = eval_grid_octree(opt, coords, eval_func, num_samples=num_samples)
        else:
            sdf_hr, sdf_lr = eval_grid(coords, eval_func, num_samples=num_samples)
        ctx = net.surs_context() if _is_accelerated(net) else _default_context(cuda)
        vol_hr = torch.from_numpy(sdf_hr.astype(np.float32)).to(ctx.device)
        vol_lr = torch.from_numpy(sdf_lr.astype(np.float32)).to(ctx.device)
    _tick("grid evaluation")
    m34 = mat[:3, :4]
    w_hr, f_hr, n_hr, v_hr, amb_hr = _mesh_from_volume(ctx, vol_hr, m34)
    w_lr, f_lr, n_lr, v_lr, amb_lr = _mesh_from_volume(ctx, vol_lr, m34)
    _tick("marching cubes x2")
    host = _to_host([w_hr, f_hr, n_hr, v_hr, w_lr, f_lr, n_lr, v_lr])
    _tick("device -> host")
    out_hr, out_lr = tuple(host[:4]), tuple(host[4:])
    if return_stats:
        stats["ambiguous_cells"] = (amb_hr, amb_lr)
        return out_hr + out_lr, stats
    return out_hr + out_lr


def _to_host(tensors):
    """Device -> host through pinned staging (torch's caching host allocator): a pageable
    .cpu() copy of a 512^3 mesh (350 MB) costs 170 ms, the pinned one 7 ms."""
    pinned = [torch.empty(t.shape, dtype=t.dtype, pin_memory=True) for t in tensors]
    for p, t in zip(pinned, tensors):
        p.copy_(t, non_blocking=True)
    torch.cuda.current_stream(tensors[0].device).synchronize()
    return [p.numpy() for p in pinned]


_CONTEXTS = {}


def _default_context(device):
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("marching cubes runs on the B200 (csrc/mc.cu); scikit-image is not used and there is no CPU fallback")
    key = dev.index if dev.index is not None else torch.cuda.current_device()
    if key not in _CONTEXTS:
        _CONTEXTS[key] = _capi.Context(torch.device("cuda", key))
    return _CONTEXTS[key]


def marching_cubes_lewiner(volume, level, device="cuda"):
    """Call-compatible stand-in for skimage.measure.marching_cubes_lewiner(volume, level) as the
    reference uses it (lib/mesh_util.py:40): numpy in, (verts, faces, normals, values) numpy out."""
    ctx = _default_context(device)
    vol = torch.from_numpy(np.ascontiguousarray(volume, dtype=np.float32)).to(ctx.device)
    verts, _, faces, normals, values, _ = ctx.marching_cubes(vol, level)
    vmin, vmax = ctx.mc_value_range()
    if not (vmin <= level <= vmax):
        raise ValueError("Surface level must be within volume data range.")
    if verts.shape[0] == 0 or faces.shape[0] == 0:
        raise RuntimeError("No surface found at the given iso value.")
    return verts.cpu().numpy(), faces.cpu().numpy(), normals.cpu().numpy(), values.cpu().numpy()


def save_obj_mesh(mesh_path, verts, faces):
    """reference lib/mesh_util.py:53-61 ('v %.4f %.4f %.4f', 1-based 'f a c b'); C writer."""
    _capi.save_obj_mesh(mesh_path, verts, faces)


def save_obj_mesh_with_color(mesh_path, verts, faces, colors):
    """reference lib/mesh_util.py:64-73."""
    with open(mesh_path, "w") as f:
        for v, c in zip(verts, colors):
            f.write("v %.4f %.4f %.4f %.4f %.4f %.4f\n" % (v[0], v[1], v[2], c[0], c[1], c[2]))
        for t in faces:
            f.write("f %d %d %d\n" % (t[0] + 1, t[2] + 1, t[1] + 1))


def save_obj_mesh_with_uv(mesh_path, verts, faces, uvs):
    """reference lib/mesh_util.py:76-89."""
    with open(mesh_path, "w") as f:
        for v, vt in zip(verts, uvs):
            f.write("v %.4f %.4f %.4f\n" % (v[0], v[1], v[2]))
            f.write("vt %.4f %.4f\n" % (vt[0], vt[1]))
        for t in faces:
            a, b, c = t[0] + 1, t[2] + 1, t[1] + 1
            f.write("f %d/%d %d/%d %d/%d\n" % (a, a, b, b, c, c))
