"""Multi-GPU reconstruction: the grid is cut into slabs along array axis 0, one per rank
(one process per GPU, ``torch.distributed`` / NCCL over NVLink).

The reference has no distributed code at all; this is new work.
Every query is independent and marching-cubes cells only need a one-plane halo, so the only
exchanges are
  1. an all-gather of four counts per rank (vertices / faces of the HR and LR mesh),
  2. the seam: ids of the vertices lying in the plane shared by two slabs travel from the lower
     rank to the upper one (2 x res^2 int32 per mesh), so that the concatenated mesh is
     bit-identical to the single-GPU one (same vertex numbering, no duplicates), and
  3. the gather of the vertex / face lists to rank 0
(each of the three once per step for BOTH meshes: the point-to-point operations are batched into one NCCL group).
Rank r evaluates planes [lo_r, hi_r + 1) (the halo plane is recomputed, not exchanged) and
meshes the cells whose lower plane it owns; axis 0 is the outermost scan axis of the marching
cubes, so concatenation in rank order *is* the single-volume order.
"""
import os
import time

import numpy as np
import torch
import torch.distributed as dist

from . import _capi

MC_LEVEL = 0.5


def slab_ranges(n_planes, world):
    """Contiguous, near-equal plane ranges [(lo, hi)] for each rank; every rank gets >= 1 plane
    when n_planes - 1 >= world (cells exist between planes)."""
    cells = n_planes - 1
    base, extra = divmod(cells, world)
    out, lo = [], 0
    for r in range(world):
        n = base + (1 if r < extra else 0)
        out.append((lo, lo + n))
        lo += n
    out[-1] = (out[-1][0], n_planes)       # the last rank keeps the final plane
    return out


def slab_u_range(res, b_min, b_max, calib, plane_lo, plane_hi):
    """Range of the image coordinate u = (calib . [x, y, z, 1])[0] (lib/geometry.py:15-31) over the grid nodes of
    planes [plane_lo, plane_hi): the extremes of an affine map over a box are at its corners."""
    c = np.asarray(calib, dtype=np.float64).reshape(-1, 4)[0]
    b_min, b_max = np.asarray(b_min, np.float64), np.asarray(b_max, np.float64)
    step = (b_max - b_min) / np.asarray(res, np.float64)
    lo = b_min + step * np.array([plane_lo, 0, 0])
    hi = b_min + step * (np.array([plane_hi, res[1], res[2]]) - 1)
    us = [c[0] * (hi[0] if k & 1 else lo[0]) + c[1] * (hi[1] if k & 2 else lo[1]) + c[2] * (hi[2] if k & 4 else lo[2]) + c[3] for k in range(8)]
    return float(min(us)), float(max(us))


def exclusive_offsets(counts):
    """counts [world, k] -> exclusive prefix sums along ranks, [world, k]."""
    c = np.asarray(counts, dtype=np.int64)
    return np.cumsum(c, axis=0) - c


def gather_rows(local, counts, dst=0, group=None):
    """Variable-length gather of row blocks: rank r contributes ``local`` [counts[r], ...];
    rank ``dst`` returns the concatenation in rank order, the others None."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if world == 1:
        return local
    counts = [int(c) for c in counts]
    if rank == dst:
        out = torch.empty((sum(counts),) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        offs = np.cumsum([0] + counts)
        out[offs[dst]:offs[dst + 1]] = local
        ops = [dist.P2POp(dist.irecv, out[offs[r]:offs[r + 1]], r, group) for r in range(world) if r != dst and counts[r] > 0]
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        return out
    if counts[rank] > 0:
        for w in dist.batch_isend_irecv([dist.P2POp(dist.isend, local.contiguous(), dst, group)]):
            w.wait()
    return None


def pass_up(tensor_out, tensor_in, group=None):
    """Rank r sends ``tensor_out`` to r + 1 and receives ``tensor_in`` from r - 1 (the seam maps)."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    ops = []
    if rank + 1 < world:
        ops.append(dist.P2POp(dist.isend, tensor_out, rank + 1, group))
    if rank > 0:
        ops.append(dist.P2POp(dist.irecv, tensor_in, rank - 1, group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()


def gather_rows_many(items, dst=0, group=None):
    """``gather_rows`` for several tensors in ONE batch of point-to-point operations (one NCCL group launch
    instead of one per tensor).  items: [(local, counts_per_rank), ...]; returns the list of concatenations on
    ``dst`` (None elsewhere; ``local`` itself when there is one rank)."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if world == 1:
        return [local for local, _ in items]
    ops, outs, keep = [], [], []
    for local, counts in items:
        counts = [int(c) for c in counts]
        if rank == dst:
            out = torch.empty((sum(counts),) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
            offs = np.cumsum([0] + counts)
            out[offs[dst]:offs[dst + 1]] = local
            ops += [dist.P2POp(dist.irecv, out[offs[r]:offs[r + 1]], r, group) for r in range(world) if r != dst and counts[r] > 0]
            outs.append(out)
        else:
            if counts[rank] > 0:
                keep.append(local.contiguous())
                ops.append(dist.P2POp(dist.isend, keep[-1], dst, group))
            outs.append(None)
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return outs


def pass_up_many(pairs, group=None):
    """``pass_up`` for several (tensor_out, tensor_in) pairs in one batch."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    ops = []
    for tensor_out, tensor_in in pairs:
        if rank + 1 < world:
            ops.append(dist.P2POp(dist.isend, tensor_out, rank + 1, group))
        if rank > 0:
            ops.append(dist.P2POp(dist.irecv, tensor_in, rank - 1, group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()


def _mc_contexts(ctx):
    """Marching-cubes state lives in the context between count and emit; a sibling context on the same device
    lets the HR and the LR volume go through count -> exchange -> emit together (one exchange round for both)."""
    sib = getattr(ctx, "_mc_sibling", None)
    if sib is None:
        sib = _capi.Context(ctx.device)
        ctx._mc_sibling = sib
    return ctx, sib


def reconstruct_slab(ctx, res, b_min, b_max, calib, z_num, z_den, mat, transform=None, precision=_capi.PREC_FP16R,
                     group=None, gather=True, want_normals=True):
    """Dense reconstruction of this rank's slab + the mesh exchange.  Returns on rank 0 the same
    8-tuple pieces as lib.mesh_util.reconstruction but as device tensors:
    ((world_hr, faces_hr, normals_hr, values_hr), (world_lr, ...)); other ranks get (None, None).
    Exchange rounds per step (both meshes together): one all-gather of 4 counts, one seam hand-over,
    one batched gather of the vertex / face lists."""
    distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    rank = dist.get_rank(group) if distributed else 0
    world = dist.get_world_size(group) if distributed else 1
    R0, R1, R2 = (int(v) for v in res)
    lo, hi = slab_ranges(R0, world)[rank]
    hi_halo = min(hi + 1, R0)
    vols = ctx.eval_grid(res, b_min, b_max, calib, z_num, z_den, transform=transform, precision=precision,
                         plane_lo=lo, plane_hi=hi_halo)
    flags = _capi.MC_LOWER_FOREIGN if rank > 0 else 0
    dev = ctx.device
    ctxs = _mc_contexts(ctx) if distributed else (ctx, ctx)
    timing = os.environ.get("SURS_TIMING") is not None and rank == 0
    t_last = [0.0]

    def tick(label):
        if timing:
            torch.cuda.synchronize(dev)
            now = time.perf_counter()
            if label:
                print("[surs timing rank0] %-28s %8.3f ms" % (label, (now - t_last[0]) * 1e3), flush=True)
            t_last[0] = now
    tick("grid evaluation" if timing and t_last[0] else None)
    if not distributed:
        results = []
        for vol in vols:                               # HR first, then LR (lib/mesh_util.py:40,45)
            nv, nf, _ = ctx.mc_count(vol, MC_LEVEL, flags)
            _, world_v, normals, values = ctx.mc_emit_verts(nv, mat, want_normals=want_normals, plane_offset=lo)
            results.append((world_v, ctx.mc_emit_faces(nf), normals, values))
        return tuple(results)
    counts = [c.mc_count(vol, MC_LEVEL, flags)[:2] for c, vol in zip(ctxs, vols)]
    tick("mc_count x2")
    mine = torch.tensor([counts[0][0], counts[0][1], counts[1][0], counts[1][1]], device=dev, dtype=torch.int64)
    allc = torch.empty((world, 4), device=dev, dtype=torch.int64)
    dist.all_gather_into_tensor(allc, mine, group=group)
    allc = allc.cpu().numpy()
    offs = exclusive_offsets(allc)
    tick("all-gather of the counts")
    emitted, seams = [], []
    for k, c in enumerate(ctxs):
        seam_out = torch.empty((2, R1, R2), device=dev, dtype=torch.int32) if rank + 1 < world else None
        seam_in = torch.empty((2, R1, R2), device=dev, dtype=torch.int32) if rank > 0 else None
        _, world_v, normals, values = c.mc_emit_verts(counts[k][0], mat, vert_id_offset=int(offs[rank, 2 * k]), seam_out=seam_out,
                                                      want_normals=want_normals, plane_offset=lo)
        emitted.append([world_v, None, normals, values])
        seams.append((seam_out, seam_in))
    tick("emit vertices x2")
    pass_up_many(seams, group)
    tick("seam hand-over")
    for k, c in enumerate(ctxs):
        emitted[k][1] = c.mc_emit_faces(counts[k][1], seam_in=seams[k][1])
    tick("emit faces x2")
    if not gather:
        return tuple(tuple(e) for e in emitted)
    items = []
    for k, e in enumerate(emitted):
        items += [(e[0], allc[:, 2 * k]), (e[1], allc[:, 2 * k + 1])]
        if want_normals:
            items += [(e[2], allc[:, 2 * k]), (e[3], allc[:, 2 * k])]
    got = gather_rows_many(items, 0, group)
    tick("gather of the mesh lists")
    if rank > 0:
        # the lower slab owns the vertices of the shared plane; a missing id means the two ranks disagree about an
        # inside / outside bit there (cannot happen while both evaluate the plane with the same arithmetic)
        bad = sum(c.mc_seam_violations() for c in ctxs)
        if bad:
            raise RuntimeError("slab seam mismatch on rank %d: %d face corners reference a vertex the lower slab does not have" % (rank, bad))
    if rank != 0:
        return (None, None)
    per = 4 if want_normals else 2
    results = []
    for k in range(2):
        g = got[per * k:per * (k + 1)]
        results.append((g[0], g[1], g[2], g[3]) if want_normals else (g[0], g[1], None, None))
    return tuple(results)


def reconstruction_from_host(ctx, feat_lr_host, feat_hr_host, res, b_min, b_max, calib, z_num, z_den, mat,
                             precision=_capi.PREC_FP16R, group=None):
    """The host-facing multi-GPU call: every rank uploads the two encoder feature maps (NCHW fp32 host tensors,
    ideally pinned) to its GPU, reconstructs its slab, and rank 0 returns the reference's 8-tuple
    (lib/mesh_util.py:8-49: verts float64 world coordinates, faces int32, normals, values -- HR then LR) as host
    numpy arrays; the other ranks return None."""
    from .lib.mesh_util import _to_host
    distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    rank = dist.get_rank(group) if distributed else 0
    world = dist.get_world_size(group) if distributed else 1
    # this rank only samples the image coordinates u of its slab (+ halo plane): upload that stripe of pixel columns
    lo, hi = slab_ranges(int(res[0]), world)[rank]
    hi = min(hi + 1, int(res[0]))
    ctx.set_features_host(feat_lr_host, feat_hr_host, u_range=slab_u_range(res, b_min, b_max, calib, lo, hi))
    hr, lr = reconstruct_slab(ctx, res, b_min, b_max, calib, z_num, z_den, mat, precision=precision, group=group)
    if hr is None:
        return None
    return tuple(_to_host(list(hr) + list(lr)))
