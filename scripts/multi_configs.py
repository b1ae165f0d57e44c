"""BASELINE.json configs 4 and 5 at N GPUs (run under torchrun; N = 1 works without it).

  C4  octree 512^3 reconstruction of N different synthetic 512x512 images, one per GPU (replicas, no collective):
      encoder (PyTorch, random init, super-resolution branch) + eval_grid_octree + marching cubes + D2H of the
      meshes, through lib.train_util's flow without the OBJ text writer (reported separately for rank 0).
  C5  raw query sweep: 2^20 .. 2^27 random points through surs_query, sharded N/G per GPU, no communication.
One JSON line per measurement on rank 0; times are the max over ranks (CUDA events / perf_counter after a barrier).
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from surs_b200 import _capi, synthetic as syn
from surs_b200.lib import mesh_util, train_util
from surs_b200.lib.model import SuRSNet
from surs_b200.lib.options import BaseOptions

FLOP = 4564998


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()

    def max_over_ranks(x):
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---------------- C4 ----------------
    res = int(os.environ.get("SURS_RES", "512"))
    opt = BaseOptions().parse(["--residual", "--resolution", str(res), "--loadSize", "1024", "--b_min", "-0.5", "-0.5", "-0.5",
                               "--b_max", "0.5", "0.5", "0.5", "--dataroot", "unused"])
    torch.manual_seed(0)
    net = SuRSNet(opt).to(dev).eval()
    for mlp in (net.mlp_lr, net.mlp_hr):                     # widen the occupancy range of the random-init MLPs
        for conv in mlp.layers():
            conv.weight.data *= 6.0
        mlp.conv4.bias.data += 0.3
    g = torch.Generator().manual_seed(1991 + rank)           # a different image on every GPU
    yy, xx = torch.meshgrid(torch.linspace(-1, 1, 512), torch.linspace(-1, 1, 512), indexing="ij")
    mask = ((xx / 0.45) ** 2 + (yy / 0.9) ** 2 < 1).float()
    img = ((torch.rand(1, 3, 8, 8, generator=g) * 2 - 1)[..., None, None].expand(1, 3, 8, 8, 64, 64).permute(0, 1, 2, 4, 3, 5).reshape(1, 3, 512, 512)
           * 0.5 + 0.5 * torch.sin(6.0 * xx + rank)[None, None]) * mask
    img = img.pin_memory()
    calib = train_util.make_calib(dev)
    b_min, b_max = np.array([-0.5] * 3), np.array([0.5] * 3)

    def one():
        t0 = time.perf_counter()
        with torch.no_grad():
            _, f_lr, f_hr = net.super_res(img.to(dev, non_blocking=True))
            net.filter_hr(f_hr)
            net.filter_lr(f_lr)
        torch.cuda.synchronize(dev)
        t1 = time.perf_counter()
        out, stats = mesh_util.reconstruction(opt, net, dev, calib, res, b_min, b_max, use_octree=True, return_stats=True)
        torch.cuda.synchronize(dev)
        return t1 - t0, time.perf_counter() - t1, out, stats

    for _ in range(3):          # warm-up while the previous result is still referenced, as a caller would hold it:
        held = one()            # two generations of pinned staging buffers end up in torch's host cache
    del held
    barrier()
    t0 = time.perf_counter()
    reps = 3
    enc = rec = 0.0
    for _ in range(reps):
        e, r, out, stats = one()
        enc += e / reps
        rec += r / reps
    barrier()
    per_image = max_over_ranks((time.perf_counter() - t0) / reps)
    enc, rec = max_over_ranks(enc), max_over_ranks(rec)
    if rank == 0:
        t0 = time.perf_counter()
        mesh_util.save_obj_mesh("/tmp/surs_c4_rank0_HR.obj", out[0], out[1])
        obj_s = time.perf_counter() - t0
        print(json.dumps({"config": "C4 octree %d^3, one synthetic 512x512 image per GPU (replicas)" % res, "n_gpus": world, "s_per_image": per_image,
                          "images_per_s": world / per_image, "encoder_s": enc, "reconstruction_to_host_s": rec,
                          "evaluated_fraction": stats["n_evaluated"] / res ** 3, "verts_hr": int(out[0].shape[0]), "faces_hr": int(out[1].shape[0]),
                          "obj_text_writer_s_hr_mesh_rank0": obj_s}), flush=True)
    del net, out
    torch.cuda.empty_cache()

    # ---------------- C5 ----------------
    if os.environ.get("SURS_SKIP_C5"):
        if world > 1:
            dist.destroy_process_group()
        return
    case = syn.SyntheticCase(S=512, seed=0)
    ctx = _capi.Context(dev)
    t = lambda a: torch.from_numpy(a).to(dev)
    ctx.set_weights([t(w) for w in case.mlp_lr[0]], [t(b) for b in case.mlp_lr[1]], [t(w) for w in case.mlp_hr[0]], [t(b) for b in case.mlp_hr[1]],
                    syn.MLP_DIM_LR, syn.MLP_DIM_HR, syn.RES_LAYERS)
    ctx.set_features(t(case.feat_lr), t(case.feat_hr))
    zn, zd = float(case.load_size // 2), float(case.z_size)
    for lg in (20, 22, 24, 26, 27):
        n = (1 << lg) // world
        pts = torch.rand(3, n, device=dev, generator=torch.Generator(device=dev).manual_seed(100 * lg + rank)) - 0.5
        ctx.query(pts, case.calib, zn, zd)
        ts = []
        for _ in range(3):
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            ctx.query(pts, case.calib, zn, zd)
            b.record()
            torch.cuda.synchronize(dev)
            ts.append(max_over_ranks(a.elapsed_time(b)))
        ms = sorted(ts)[1]
        if rank == 0:
            print(json.dumps({"config": "C5 query sweep", "n_gpus": world, "points": n * world, "ms": ms, "queries_per_s": n * world / ms * 1e3,
                              "algorithmic_tflops": n * world * FLOP / ms * 1e-9}), flush=True)
        del pts
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
