"""Multi-GPU reconstruction: the grid is cut into slabs along array axis 0, one per rank
(one process per GPU, ``torch.distributed`` / NCCL over NVLink).

The reference has no distributed code at all; this is new work.
Every query is independent and marching-cubes cells only need a one-plane halo, so the only
exchanges are
  1. an all-gather of four counts per rank (vertices / faces of the HR and LR mesh),
  2. the seam: ids of the vertices lying in the plane shared by two slabs travel from the lower
     rank to the upper one (2 x res^2 int32 per mesh), so that the concatenated mesh is
     bit-identical to the single-GPU one (same vertex numbering, no duplicates), and
  3. the gather of the vertex / face lists on rank 0 -- FUSED INTO THE EMISSION: rank 0 owns two
     ping-pong device arenas that every other rank maps through CUDA IPC (NVLink peer memory); the
     marching-cubes emit kernels of rank r write its vertices / faces straight at rank r's offset
     in rank 0's arrays, followed by one tiny all-reduce as the completion barrier.  No collective
     carries the payload (SURS_NCCL_GATHER=1 selects the NCCL send / recv gather instead).
For host results (`reconstruction_from_host`) nothing is gathered on a device at all: every rank
copies its part over ITS OWN PCIe link into a pinned shared-memory host arena at its offsets and
rank 0 returns zero-copy numpy views.
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


def weighted_slab_ranges(n_planes, shares):
    """Contiguous plane ranges whose CELL counts follow `shares` (one positive weight per rank) as closely as integer
    counts allow; every rank keeps at least one cell layer; the last rank keeps the final plane (as slab_ranges)."""
    world = len(shares)
    cells = n_planes - 1
    w = np.maximum(np.asarray(shares, dtype=np.float64), 1e-9)
    ideal = np.cumsum(w) / w.sum() * cells
    cuts = [0]
    for r in range(world - 1):
        lo_allowed, hi_allowed = cuts[-1] + 1, cells - (world - 1 - r)
        cuts.append(int(min(max(int(round(ideal[r])), lo_allowed), hi_allowed)))
    cuts.append(cells)
    out = [(cuts[r], cuts[r + 1]) for r in range(world)]
    out[-1] = (out[-1][0], n_planes)
    return out


def current_slab_ranges(ctx, n_planes, world):
    """The partition the next reconstruct_slab of this context will use: equal slabs, or -- after the first step and
    unless SURS_BALANCE=0 -- slabs sized by the per-rank grid-evaluation times measured in the previous step."""
    st = getattr(ctx, "_balance", None)
    if st is None or st["n_planes"] != n_planes or st["world"] != world or os.environ.get("SURS_BALANCE") == "0":
        return slab_ranges(n_planes, world)
    return weighted_slab_ranges(n_planes, st["shares"])


BALANCE_STEPS = 3      # calibration steps after which the partition is frozen (changing slab sizes re-allocates scratch)


def _update_balance(ctx, n_planes, world, ranges, times_us):
    """New shares from the measured per-rank times: a rank's speed is (cells it had) / (time it took); the next
    partition gives every rank cells in proportion to its speed (damped by one half, so that noise does not make the
    cuts oscillate).  Every rank calls this with the same all-gathered numbers, hence computes the same partition.
    Only the first BALANCE_STEPS steps of a (resolution, world size) calibrate; then the partition stays fixed, so that
    the slabs' scratch buffers and output tensors keep their sizes."""
    st = getattr(ctx, "_balance", None)
    same = st is not None and st["n_planes"] == n_planes and st["world"] == world
    if not same:
        # the very first step of a configuration pays one-off allocations: it only opens the calibration
        had0 = np.array([hi - lo for lo, hi in ranges], dtype=np.float64)
        had0[-1] -= 1.0
        ctx._balance = {"n_planes": n_planes, "world": world, "shares": had0 / had0.sum(), "times_us": None, "updates": 0}
        return
    if st["updates"] >= BALANCE_STEPS:
        return
    updates = st["updates"] + 1
    t = np.maximum(np.asarray(times_us, dtype=np.float64), 1.0)
    had = np.array([hi - lo for lo, hi in ranges], dtype=np.float64)
    had[-1] -= 1.0                                       # the last rank's range includes the final plane, not a cell layer
    speed = np.maximum(had, 1.0) / t
    target = speed / speed.sum()
    old = had / had.sum()
    shares = 0.5 * old + 0.5 * target
    ctx._balance = {"n_planes": n_planes, "world": world, "shares": shares, "times_us": [float(v) for v in t], "updates": updates}


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


# ---------------------------------------------------------------------------------------------
# peer arenas (device, owned by rank `dst`) and the pinned shared host arena
# ---------------------------------------------------------------------------------------------
def _align(n, a=256):
    return (int(n) + a - 1) // a * a


def mesh_layout(tot_v, tot_f, want_normals=True):
    """Byte offsets of the eight result arrays (HR then LR: world float64 [V,3], faces int32 [F,3], normals float32
    [V,3], values float32 [V]) inside an arena, and the total size."""
    off, lay = 0, []
    for k in range(2):
        d = {}
        for name, nbytes in (("world", 24 * int(tot_v[k])), ("faces", 12 * int(tot_f[k])),
                             ("normals", 12 * int(tot_v[k]) if want_normals else 0), ("values", 4 * int(tot_v[k]) if want_normals else 0)):
            d[name] = off
            off = _align(off + nbytes)
        lay.append(d)
    return lay, off


def _all_ok(ok, device, group):
    """Collective agreement: True only when every rank succeeded (so that all ranks take the same fallback)."""
    flag = torch.tensor([1 if ok else 0], device=device, dtype=torch.int32)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
    return bool(int(flag.item()))


def _peer_arenas(ctx, group, need, dst=0):
    """Two ping-pong device arenas on rank `dst`, mapped into every rank (collective: every rank passes the same need).
    Returns None -- on every rank -- when CUDA IPC is not available between the ranks (the caller then gathers over NCCL)."""
    st = getattr(ctx, "_peer_arenas", None)
    if st is not None and (st is False or st["capacity"] >= need):
        return st or None
    rank = dist.get_rank(group)
    if st is not None:
        for p in st["ptrs"]:
            (ctx.arena_destroy if st["owner"] else ctx.arena_close)(p)
    capacity = _align(int(need * 1.5) + (1 << 20), 1 << 20)
    local, handles, ok = [], [None, None], True
    if rank == dst:
        try:
            local = [ctx.arena_create(capacity) for _ in range(2)]
            handles = [h for _, h in local]
        except RuntimeError:
            ok = False
    box = [handles]
    dist.broadcast_object_list(box, src=dist.get_global_rank(group, dst) if group is not None else dst, group=group)
    ptrs = [p for p, _ in local]
    if rank != dst and box[0][0] is not None:
        try:
            ptrs = [ctx.arena_open(h) for h in box[0]]
        except RuntimeError:
            ok = False
    if not _all_ok(ok and box[0][0] is not None, ctx.device, group):
        import warnings
        warnings.warn("surs_b200: CUDA IPC peer arenas are not available between the ranks; gathering the meshes over NCCL instead")
        ctx._peer_arenas = False
        return None
    st = {"capacity": capacity, "ptrs": ptrs, "owner": rank == dst, "turn": 0}
    ctx._peer_arenas = st
    return st


class HostArena:
    """A POSIX shared-memory segment registered as pinned memory in every rank's process: rank r copies its slice of
    the result over its own PCIe link, rank 0 reads everything without another copy.  Two halves, used alternately."""

    def __init__(self, ctx, group, need, pin=True):
        from multiprocessing import shared_memory
        rank = dist.get_rank(group)
        self.capacity = _align(int(need * 1.5) + (1 << 20), 1 << 20)
        box = [None]
        if rank == 0:
            self.shm = shared_memory.SharedMemory(create=True, size=2 * self.capacity)
            box = [self.shm.name]
        dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        if rank != 0:
            self.shm = shared_memory.SharedMemory(name=box[0])
            try:                                      # only the creator unlinks the segment
                from multiprocessing import resource_tracker
                resource_tracker.unregister(self.shm._name, "shared_memory")
            except Exception:
                pass
        self.owner = rank == 0
        self.bytes = torch.frombuffer(self.shm.buf, dtype=torch.uint8, count=2 * self.capacity)
        self.pinned = bool(pin)
        if pin:
            rc = torch.cuda.cudart().cudaHostRegister(self.bytes.data_ptr(), 2 * self.capacity, 0)
            if int(rc) != 0:
                raise RuntimeError("cudaHostRegister of the shared host arena failed (%s)" % rc)
        self.turn = 0

    def half(self):
        v = self.bytes[self.turn * self.capacity:(self.turn + 1) * self.capacity]
        self.turn ^= 1
        return v

    def close(self):
        if self.pinned:
            try:
                torch.cuda.cudart().cudaHostUnregister(self.bytes.data_ptr())
            except Exception:
                pass
        if getattr(self, "closed", False):
            return
        self.closed = True
        del self.bytes
        try:
            self.shm.close()
        except BufferError:                                # numpy views handed to the caller are still alive
            pass
        if self.owner:
            try:
                self.shm.unlink()
            except FileNotFoundError:
                pass


_ROW_BYTES = {"world": 24, "faces": 12, "normals": 12, "values": 4}


def fill_host_arena(half, layout, emitted, offs, rank):
    """This rank's slices of the eight result arrays -> their place in the arena half (asynchronous for CUDA sources and
    a pinned arena).  emitted: per mesh (world, faces, normals, values); offs: exclusive row offsets [world, 4]."""
    for k, e in enumerate(emitted):
        for name, t, off_rows in (("world", e[0], offs[rank, 2 * k]), ("faces", e[1], offs[rank, 2 * k + 1]),
                                  ("normals", e[2], offs[rank, 2 * k]), ("values", e[3], offs[rank, 2 * k])):
            if t is None or t.numel() == 0:
                continue
            start = layout[k][name] + _ROW_BYTES[name] * int(off_rows)
            half[start:start + t.numel() * t.element_size()].view(t.dtype).view(t.shape).copy_(t, non_blocking=True)


def read_host_arena(half, layout, tot_v, tot_f):
    """Zero-copy numpy views of the eight arrays (the reference's 8-tuple order)."""
    out, buf = [], half.numpy()
    for k in range(2):
        L, nv, nf = layout[k], int(tot_v[k]), int(tot_f[k])
        out += [buf[L["world"]:L["world"] + 24 * nv].view(np.float64).reshape(nv, 3),
                buf[L["faces"]:L["faces"] + 12 * nf].view(np.int32).reshape(nf, 3),
                buf[L["normals"]:L["normals"] + 12 * nv].view(np.float32).reshape(nv, 3),
                buf[L["values"]:L["values"] + 4 * nv].view(np.float32).reshape(nv)]
    return tuple(out)


def _host_arena(ctx, group, need):
    """The shared pinned host arena of this context (collective).  None on every rank when it cannot be set up
    (e.g. /dev/shm too small): the caller then gathers on rank 0's device and copies from there."""
    import atexit
    st = getattr(ctx, "_host_arena", None)
    if st is False:
        return None
    if st is not None and st.capacity >= need:
        return st
    if st is not None:
        st.close()
    try:
        st = HostArena(ctx, group, need)
        ok = True
    except Exception:                                  # the barrier inside HostArena may not have been reached: agree below
        st, ok = None, False
    if not _all_ok(ok, ctx.device, group):
        if st is not None:
            st.close()
        import warnings
        warnings.warn("surs_b200: the pinned shared-memory host arena could not be set up; gathering the meshes on rank 0 instead")
        ctx._host_arena = False
        return None
    ctx._host_arena = st
    atexit.register(st.close)                          # unlink the shared-memory segment when the process ends
    return st


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
    ranges = current_slab_ranges(ctx, R0, world) if distributed else slab_ranges(R0, world)
    lo, hi = ranges[rank]
    hi_halo = min(hi + 1, R0)
    ev0, ev1 = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) if distributed else (None, None)
    if distributed:
        ev0.record()
    vols = ctx.eval_grid(res, b_min, b_max, calib, z_num, z_den, transform=transform, precision=precision,
                         plane_lo=lo, plane_hi=hi_halo)
    if distributed:
        ev1.record()
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
    # the counts travel together with this rank's grid-evaluation time (its events are complete: mc_count synchronised),
    # from which every rank derives the same partition for the NEXT step (adaptive slab balance, current_slab_ranges)
    grid_us = int(ev0.elapsed_time(ev1) * 1e3)
    mine = torch.tensor([counts[0][0], counts[0][1], counts[1][0], counts[1][1], grid_us], device=dev, dtype=torch.int64)
    allc = torch.empty((world, 5), device=dev, dtype=torch.int64)
    dist.all_gather_into_tensor(allc, mine, group=group)
    allc = allc.cpu().numpy()
    _update_balance(ctx, R0, world, ranges, allc[:, 4])
    allc = np.ascontiguousarray(allc[:, :4])
    offs = exclusive_offsets(allc)
    tick("all-gather of the counts")
    # gather == True: emit straight into rank 0's arena through the NVLink peer mapping (no payload collective)
    fused = gather is True and os.environ.get("SURS_NCCL_GATHER") is None
    layout = base = None
    if fused:
        tot_v, tot_f = allc[:, [0, 2]].sum(0), allc[:, [1, 3]].sum(0)
        layout, need = mesh_layout(tot_v, tot_f, want_normals)
        arenas = _peer_arenas(ctx, group, need)
        if arenas is None:
            fused = False
        else:
            base = arenas["ptrs"][arenas["turn"]]
            arenas["turn"] ^= 1
    emitted, seams = [], []
    for k, c in enumerate(ctxs):
        seam_out = torch.empty((2, R1, R2), device=dev, dtype=torch.int32) if rank + 1 < world else None
        seam_in = torch.empty((2, R1, R2), device=dev, dtype=torch.int32) if rank > 0 else None
        ov = int(offs[rank, 2 * k])
        if fused:
            L = layout[k]
            c.mc_emit_verts(counts[k][0], mat, vert_id_offset=ov, seam_out=seam_out, plane_offset=lo,
                            out_ptrs=(base + L["world"] + 24 * ov, (base + L["normals"] + 12 * ov) if want_normals else None,
                                      (base + L["values"] + 4 * ov) if want_normals else None))
            emitted.append(None)
        else:
            _, world_v, normals, values = c.mc_emit_verts(counts[k][0], mat, vert_id_offset=ov, seam_out=seam_out,
                                                          want_normals=want_normals, plane_offset=lo)
            emitted.append([world_v, None, normals, values])
        seams.append((seam_out, seam_in))
    tick("emit vertices x2")
    pass_up_many(seams, group)
    tick("seam hand-over")
    for k, c in enumerate(ctxs):
        if fused:
            c.mc_emit_faces(counts[k][1], seam_in=seams[k][1], out_ptr=base + layout[k]["faces"] + 12 * int(offs[rank, 2 * k + 1]))
        else:
            emitted[k][1] = c.mc_emit_faces(counts[k][1], seam_in=seams[k][1])
    tick("emit faces x2")
    if gather == "local":
        return tuple(tuple(e) for e in emitted), allc, offs
    if not gather:
        return tuple(tuple(e) for e in emitted)

    def check_seams():
        if rank > 0:
            # the lower slab owns the vertices of the shared plane; a missing id means the two ranks disagree about an
            # inside / outside bit there (cannot happen while both evaluate the plane with the same arithmetic)
            bad = sum(c.mc_seam_violations() for c in ctxs)
            if bad:
                raise RuntimeError("slab seam mismatch on rank %d: %d face corners reference a vertex the lower slab does not have" % (rank, bad))

    if fused:
        # completion barrier: stream-ordered behind every rank's emit kernels; when it has run on rank 0 all peer
        # writes have landed in its arena
        flag = torch.zeros(1, device=dev, dtype=torch.int32)
        dist.all_reduce(flag, group=group)
        tick("completion all-reduce")
        check_seams()
        if rank != 0:
            return (None, None)
        results = []
        for k in range(2):
            L, nv, nf = layout[k], int(tot_v[k]), int(tot_f[k])
            results.append((_capi.tensor_from_ptr(base + L["world"], (nv, 3), torch.float64, dev),
                            _capi.tensor_from_ptr(base + L["faces"], (nf, 3), torch.int32, dev),
                            _capi.tensor_from_ptr(base + L["normals"], (nv, 3), torch.float32, dev) if want_normals else None,
                            _capi.tensor_from_ptr(base + L["values"], (nv,), torch.float32, dev) if want_normals else None))
        return tuple(results)
    items = []
    for k, e in enumerate(emitted):
        items += [(e[0], allc[:, 2 * k]), (e[1], allc[:, 2 * k + 1])]
        if want_normals:
            items += [(e[2], allc[:, 2 * k]), (e[3], allc[:, 2 * k])]
    got = gather_rows_many(items, 0, group)
    tick("gather of the mesh lists")
    check_seams()
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
    numpy arrays; the other ranks return None.  With several ranks the arrays are zero-copy views of a pinned
    shared-memory arena that every rank filled over its own PCIe link; they stay valid until the call after next
    (two halves, used alternately).  SURS_NCCL_GATHER=1: gather on rank 0's device and copy from there instead."""
    from .lib.mesh_util import _to_host
    distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    rank = dist.get_rank(group) if distributed else 0
    world = dist.get_world_size(group) if distributed else 1
    # this rank only samples the image coordinates u of its slab (+ halo plane): upload that stripe of pixel columns
    lo, hi = (current_slab_ranges(ctx, int(res[0]), world) if distributed else slab_ranges(int(res[0]), world))[rank]
    hi = min(hi + 1, int(res[0]))
    ctx.set_features_host(feat_lr_host, feat_hr_host, u_range=slab_u_range(res, b_min, b_max, calib, lo, hi))
    if not distributed or os.environ.get("SURS_NCCL_GATHER") is not None:
        hr, lr = reconstruct_slab(ctx, res, b_min, b_max, calib, z_num, z_den, mat, precision=precision, group=group)
        if hr is None:
            return None
        return tuple(_to_host(list(hr) + list(lr)))
    emitted, allc, offs = reconstruct_slab(ctx, res, b_min, b_max, calib, z_num, z_den, mat, precision=precision, group=group, gather="local")
    tot_v, tot_f = allc[:, [0, 2]].sum(0), allc[:, [1, 3]].sum(0)
    layout, need = mesh_layout(tot_v, tot_f, True)
    arena = _host_arena(ctx, group, need)
    if arena is None:                                   # fallback: NCCL gather of what was emitted, host copy on rank 0
        items = []
        for k, e in enumerate(emitted):
            items += [(e[0], allc[:, 2 * k]), (e[1], allc[:, 2 * k + 1]), (e[2], allc[:, 2 * k]), (e[3], allc[:, 2 * k])]
        got = gather_rows_many(items, 0, group)
        return tuple(_to_host(got)) if rank == 0 else None
    half = arena.half()
    fill_host_arena(half, layout, emitted, offs, rank)
    torch.cuda.current_stream(ctx.device).synchronize()
    bad = sum(c.mc_seam_violations() for c in _mc_contexts(ctx)) if rank > 0 else 0
    if bad:
        raise RuntimeError("slab seam mismatch on rank %d: %d face corners reference a vertex the lower slab does not have" % (rank, bad))
    dist.barrier(group)                                   # every rank's copies have landed in the host arena
    if rank != 0:
        return None
    return read_host_arena(half, layout, tot_v, tot_f)
