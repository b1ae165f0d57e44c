"""Multi-GPU slab reconstruction == single-GPU reconstruction, bit for bit (needs >= 2 GPUs;
skipped on the single-GPU test box, run with `gpurun --gpus 2 -- pytest tests/test_gpu_multi.py -m gpu`)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, res, out, refined=False):
    from surs_b200 import _capi, parallel, synthetic as syn
    from surs_b200.lib import sdf as bsdf
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        dev = torch.device("cuda", rank)
        case = syn.SyntheticCase(S=32, seed=0)
        ctx = _capi.Context(dev)
        t = lambda a: torch.from_numpy(a).to(dev)
        ctx.set_weights([t(w) for w in case.mlp_lr[0]], [t(b) for b in case.mlp_lr[1]], [t(w) for w in case.mlp_hr[0]], [t(b) for b in case.mlp_hr[1]],
                        syn.MLP_DIM_LR, syn.MLP_DIM_HR, syn.RES_LAYERS)
        ctx.set_features(t(case.feat_lr), t(case.feat_hr))
        zn, zd = float(case.load_size // 2), float(case.z_size)
        bmin, bmax = np.array([-0.5] * 3), np.array([0.5] * 3)
        mat = bsdf.grid_matrix(res, bmin, bmax)[:3, :4]
        prec = _capi.PREC_FP16R if refined else _capi.PREC_FP16
        got = parallel.reconstruct_slab(ctx, (res,) * 3, bmin, bmax, case.calib, zn, zd, mat, precision=prec)
        ok = True
        # the three exchange paths give the same mesh: emission fused with the gather through the NVLink peer arena
        # (default, twice: the arenas ping-pong), the NCCL send / recv gather, and the pinned shared host arena
        got2 = parallel.reconstruct_slab(ctx, (res,) * 3, bmin, bmax, case.calib, zn, zd, mat, precision=prec)
        os.environ["SURS_NCCL_GATHER"] = "1"
        got_nccl = parallel.reconstruct_slab(ctx, (res,) * 3, bmin, bmax, case.calib, zn, zd, mat, precision=prec)
        del os.environ["SURS_NCCL_GATHER"]
        f_lr_h, f_hr_h = torch.from_numpy(case.feat_lr).pin_memory(), torch.from_numpy(case.feat_hr).pin_memory()
        host = parallel.reconstruction_from_host(ctx, f_lr_h, f_hr_h, (res,) * 3, bmin, bmax, case.calib, zn, zd, mat, precision=prec)
        host2 = parallel.reconstruction_from_host(ctx, f_lr_h, f_hr_h, (res,) * 3, bmin, bmax, case.calib, zn, zd, mat, precision=prec)
        ctx.set_features(t(case.feat_lr), t(case.feat_hr))               # whole maps again for the single-GPU check below
        if rank == 0:
            # vertices, faces and values are identical for ANY partition (the calls above ran with different slab sizes:
            # the first steps calibrate the adaptive balance); normals on a slab's first / last plane use a one-sided
            # difference, so they may differ where the cuts moved
            flat = lambda g: [x for mesh in g for k, x in enumerate(mesh) if k != 2]
            for other in (got2, got_nccl):
                ok = ok and all(torch.equal(a, b) for a, b in zip(flat(got), flat(other)))
            for h in (host, host2):
                hh = [h[i] for i in (0, 1, 3, 4, 5, 7)]
                ok = ok and all(np.array_equal(a.cpu().numpy(), b) for a, b in zip(flat(got), hh))
            print("rank0: fused / NCCL / host-arena paths identical:", ok, flush=True)
        else:
            ok = got == (None, None) and host is None
        if rank == 0:
            vols = ctx.eval_grid((res,) * 3, bmin, bmax, case.calib, zn, zd, precision=prec)
            for (w, f, n, v), vol in zip(got, vols):
                _, w1, f1, n1, v1, _ = ctx.marching_cubes(vol, 0.5, mat)
                # refined mode: a slab judges the nodes of its first / last plane by the neighbours inside the slab, so
                # volumes (and the per-vertex max-corner `values`) may differ there -- vertices and faces may not
                same = (w.shape == w1.shape and torch.equal(w, w1), f.shape == f1.shape and torch.equal(f, f1), refined or torch.equal(v, v1))
                # normals use the volume gradient: at a slab's first / last plane the one-sided difference of the
                # slab replaces the central difference of the full volume, so they agree everywhere else only
                nd = (n - n1).abs().amax(dim=1)
                print("rank0: verts/faces/values equal:", same, "normals differing:", int((nd > 1e-6).sum()), "of", n.shape[0], flush=True)
                ok = ok and all(same) and float((nd > 1e-6).float().mean()) < 0.1
        out.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,refined", [(2, False), (3, False), (2, True)])
def test_slab_reconstruction_equals_single_gpu(world, refined):
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, 96, q, refined)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res
