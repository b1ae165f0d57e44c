"""Multi-GPU reconstruction: the grid is cut into slabs along array axis 0, one per rank
(one process per GPU, ``torch.distributed`` / NCCL over NVLink).

The reference has no distributed code at all; this is new work.
Every query is independent and marching-cubes cells only need a one-plane halo, so the only
exchanges are
  1. an all-gather of four counts per rank (vertices / faces of the HR and LR mesh),
  2. the seam: ids of the vertices lying in the plane shared by two slabs travel from the lower
     rank to the upper one (2 x res^2 int32 per mesh), so that the concatenated mesh is
     bit-identical to the single-GPU one (same vertex numbering, no duplicates), and
  3. the gather of the vertex / face lists to rank 0.
Rank r evaluates planes [lo_r, hi_r + 1) (the halo plane is recomputed, not exchanged) and
meshes the cells whose lower plane it owns; axis 0 is the outermost scan axis of the marching
cubes, so concatenation in rank order *is* the single-volume order.
"""
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


def reconstruct_slab(ctx, res, b_min, b_max, calib, z_num, z_den, mat, transform=None, precision=_capi.PREC_FP16,
                     group=None, gather=True, want_normals=True):
    """Dense reconstruction of this rank's slab + the mesh exchange.  Returns on rank 0 the same
    8-tuple pieces as lib.mesh_util.reconstruction but as device tensors:
    ((world_hr, faces_hr, normals_hr, values_hr), (world_lr, ...)); other ranks get (None, None)."""
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
    results = []
    for vol in vols:                                   # HR first, then LR (lib/mesh_util.py:40,45)
        nv, nf, _ = ctx.mc_count(vol, MC_LEVEL, flags)
        if distributed:
            mine = torch.tensor([nv, nf], device=dev, dtype=torch.int64)
            allc = torch.empty((world, 2), device=dev, dtype=torch.int64)
            dist.all_gather_into_tensor(allc, mine, group=group)
            allc = allc.cpu().numpy()
        else:
            allc = np.array([[nv, nf]], dtype=np.int64)
        v_off = int(exclusive_offsets(allc)[rank, 0])
        seam_out = torch.empty((2, R1, R2), device=dev, dtype=torch.int32) if distributed and rank + 1 < world else None
        verts, world_v, normals, values = ctx.mc_emit_verts(nv, mat, vert_id_offset=v_off, seam_out=seam_out,
                                                            want_normals=want_normals, plane_offset=lo)
        seam_in = torch.empty((2, R1, R2), device=dev, dtype=torch.int32) if rank > 0 else None
        if distributed:
            pass_up(seam_out, seam_in, group)
        faces = ctx.mc_emit_faces(nf, seam_in=seam_in)
        if distributed and gather:
            world_v = gather_rows(world_v, allc[:, 0], 0, group)
            faces = gather_rows(faces, allc[:, 1], 0, group)
            if want_normals:
                normals = gather_rows(normals, allc[:, 0], 0, group)
                values = gather_rows(values, allc[:, 0], 0, group)
        results.append((world_v, faces, normals, values) if rank == 0 or not gather else None)
    return tuple(results)
