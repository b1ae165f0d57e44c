"""Host-side logic of the slab-parallel reconstruction (surs_b200/parallel.py) on CPU:
plane partitioning, offsets, and the two exchanges (variable-length gather, seam pass-up) with
world_size 2 over gloo."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from surs_b200 import parallel


def test_slab_ranges_cover_all_cells_once():
    for planes in (2, 3, 17, 128, 512):
        for world in (1, 2, 3, 4, 8):
            if planes - 1 < world:
                continue
            r = parallel.slab_ranges(planes, world)
            assert r[0][0] == 0 and r[-1][1] == planes
            cells = []
            for k, (lo, hi) in enumerate(r):
                last = hi if k < world - 1 else hi - 1       # a cell is owned by the rank owning its lower plane
                cells.extend(range(lo, last))
                assert hi > lo
            assert cells == list(range(planes - 1))
            sizes = [hi - lo for lo, hi in r[:-1]] + [r[-1][1] - 1 - r[-1][0]]
            assert max(sizes) - min(sizes) <= 1


def test_slab_u_range():
    calib = np.diag([2.0, -2.0, 2.0, 1.0])
    lo, hi = parallel.slab_u_range((512, 512, 512), [-0.5] * 3, [0.5] * 3, calib, 64, 129)
    assert np.isclose(lo, 2 * (-0.5 + 64 / 512)) and np.isclose(hi, 2 * (-0.5 + 128 / 512))
    sheared = calib.copy()
    sheared[0, 1] = 0.5                                   # u also depends on y: the range widens by 0.5 * the y extent
    lo2, hi2 = parallel.slab_u_range((512, 512, 512), [-0.5] * 3, [0.5] * 3, sheared, 64, 129)
    assert lo2 < lo and hi2 > hi and np.isclose(hi2 - lo2, (hi - lo) + 0.5 * 511 / 512)


def test_exclusive_offsets():
    off = parallel.exclusive_offsets([[3, 5], [0, 2], [7, 1]])
    assert off.tolist() == [[0, 0], [3, 5], [3, 7]]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        counts = [5, 0, 3][:world] if world == 3 else [4, 7]
        local = torch.arange(counts[rank] * 3, dtype=torch.float32).reshape(counts[rank], 3) + 100 * rank
        got = parallel.gather_rows(local, counts, 0)
        seam_out = torch.full((2, 4, 4), rank, dtype=torch.int32)
        seam_in = torch.full((2, 4, 4), -7, dtype=torch.int32)
        parallel.pass_up(seam_out, seam_in)
        # the batched variants (one group of point-to-point operations for several tensors)
        counts2 = counts[::-1]
        local2 = torch.arange(counts2[rank], dtype=torch.int32) - 50 * rank
        many = parallel.gather_rows_many([(local, counts), (local2, counts2)], 0)
        seams = [(torch.full((2, 3), 10 * k + rank, dtype=torch.int32), torch.full((2, 3), -7, dtype=torch.int32)) for k in range(2)]
        parallel.pass_up_many(seams)
        if rank == 0:
            want = torch.cat([torch.arange(c * 3, dtype=torch.float32).reshape(c, 3) + 100 * r for r, c in enumerate(counts)])
            want2 = torch.cat([torch.arange(c, dtype=torch.int32) - 50 * r for r, c in enumerate(counts2)])
            ok = torch.equal(got, want) and bool((seam_in == -7).all())
            ok = ok and torch.equal(many[0], want) and torch.equal(many[1], want2) and all(bool((si == -7).all()) for _, si in seams)
        else:
            ok = got is None and bool((seam_in == rank - 1).all())
            ok = ok and many == [None, None] and all(bool((si == 10 * k + rank - 1).all()) for k, (_, si) in enumerate(seams))
        out.put((rank, ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_gather_and_seam_exchange_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res


def test_mesh_layout_is_aligned_and_disjoint():
    lay, total = parallel.mesh_layout([1000, 7], [1999, 11])
    spans = []
    for k, (nv, nf) in enumerate(((1000, 1999), (7, 11))):
        for name, n in (("world", 24 * nv), ("faces", 12 * nf), ("normals", 12 * nv), ("values", 4 * nv)):
            assert lay[k][name] % 256 == 0
            spans.append((lay[k][name], lay[k][name] + n))
    spans.sort()
    assert all(a[1] <= b[0] for a, b in zip(spans, spans[1:])) and spans[-1][1] <= total


def _arena_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # every rank holds its rows of the two meshes; the shared host arena receives them at the global offsets
        nv = np.array([[3, 2], [0, 4], [5, 1]][:world])           # [rank][mesh] vertices
        nf = 2 * nv + 1
        allc = np.stack([nv[:, 0], nf[:, 0], nv[:, 1], nf[:, 1]], 1)
        offs = parallel.exclusive_offsets(allc)
        tot_v, tot_f = allc[:, [0, 2]].sum(0), allc[:, [1, 3]].sum(0)
        layout, need = parallel.mesh_layout(tot_v, tot_f)
        arena = parallel.HostArena(None, None, need, pin=False)
        mk = lambda n, cols, dt, tag: (torch.arange(n * cols, dtype=torch.float64).reshape((n, cols) if cols > 1 else (n,)) + 1000 * rank + tag).to(dt)
        for rep in range(3):                                       # the two halves alternate
            emitted = [(mk(nv[rank, k], 3, torch.float64, 10 * k + rep), mk(nf[rank, k], 3, torch.int32, 20 * k), mk(nv[rank, k], 3, torch.float32, 30 * k),
                        mk(nv[rank, k], 1, torch.float32, 40 * k)) for k in range(2)]
            half = arena.half()
            parallel.fill_host_arena(half, layout, emitted, offs, rank)
            dist.barrier()
            ok = True
            if rank == 0:
                got = parallel.read_host_arena(half, layout, tot_v, tot_f)
                for k in range(2):
                    for j, (cols, dt, tag, cnt) in enumerate(((3, torch.float64, 10 * k + rep, nv), (3, torch.int32, 20 * k, nf), (3, torch.float32, 30 * k, nv),
                                                              (1, torch.float32, 40 * k, nv))):
                        want = torch.cat([(torch.arange(cnt[r, k] * cols, dtype=torch.float64).reshape((cnt[r, k], cols) if cols > 1 else (cnt[r, k],))
                                           + 1000 * r + tag).to(dt) for r in range(world)]).numpy()
                        ok = ok and np.array_equal(got[4 * k + j], want)
            dist.barrier()
            out.put((rank, ok))
        arena.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_shared_host_arena_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_arena_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(3 * world)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res


def test_weighted_slab_ranges_and_balance_update():
    class Ctx:
        pass
    ctx = Ctx()
    for planes, world in ((512, 8), (96, 3), (9, 4)):
        eq = parallel.slab_ranges(planes, world)
        assert parallel.current_slab_ranges(ctx, planes, world) == eq
        w = parallel.weighted_slab_ranges(planes, [1.0] * world)           # equal weights: the equal partition up to rounding
        assert all(abs(a[0] - b[0]) <= 1 and abs(a[1] - b[1]) <= 1 for a, b in zip(w, eq)) and w[0][0] == 0 and w[-1][1] == planes
    # rank 2 is 20 % slower: after a few updates it holds fewer cells, all cells are still covered exactly once
    planes, world = 512, 4
    ranges = parallel.slab_ranges(planes, world)
    frozen = None
    for it in range(7):
        if it == parallel.BALANCE_STEPS + 1:                       # (the first step only opens the calibration)
            frozen = ranges                                        # calibration is over: the partition no longer moves
        cells = [hi - lo - (1 if r == world - 1 else 0) for r, (lo, hi) in enumerate(ranges)]
        times = [c * (1.2 if r == 2 else 1.0) * 400 for r, c in enumerate(cells)]
        parallel._update_balance(ctx, planes, world, ranges, times)
        ranges = parallel.current_slab_ranges(ctx, planes, world)
        assert ranges[0][0] == 0 and ranges[-1][1] == planes and all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
        assert all(hi - lo >= 1 for lo, hi in ranges)
    assert ranges == frozen
    cells = [hi - lo - (1 if r == world - 1 else 0) for r, (lo, hi) in enumerate(ranges)]
    assert cells[2] < cells[0] and abs(cells[2] * 1.2 - cells[0]) <= 6.0 and sum(cells) == planes - 1
    # extreme weights still leave every rank a cell layer
    r = parallel.weighted_slab_ranges(9, [1000.0, 1.0, 1.0, 1.0])
    assert [hi - lo for lo, hi in r] == [5, 1, 1, 2] and r[-1][1] == 9
