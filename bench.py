#!/usr/bin/env python
"""Benchmark of the SuRS reconstruction hot path on B200 (driver contract: see the task statement).

    python bench.py --gpus N --steps K --warmup W            # ours (N > 1: launched under torchrun)
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU path (torch port)

A step = one dense 512^3 reconstruction of one synthetic 512x512 input (BASELINE.json configs[2]):
occupancy query of all 134 217 728 grid nodes (projection + bilinear indexing + both MLPs) and
marching cubes of both volumes; at N > 1 the grid is slab-sharded over the ranks (fixed total
work: strong scaling) and the mesh lists are gathered on rank 0 over NCCL.  Features and weights
are resident in HBM when the timed region starts.  `value` = grid nodes / step time.

`e2e` = the same reconstruction through the public API lib.mesh_util.reconstruction(...) with
HOST buffers on both sides: the two feature maps come from pinned host memory every step
(H2D + repack inside the timed region) and the eight mesh arrays end in host numpy (D2H).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_QUERY = 4564998          # 2 x (1 140 545 + 1 141 954) MAC
METRIC = "occupancy_queries_per_s_512cubed_dense_recon"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    def __init__(self, index):
        self.rows = []
        self.stop = threading.Event()
        self.index = index
        self.thread = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([s.strip() for s in out.split(",")])
            except Exception:
                pass
            self.stop.wait(0.2)

    def __enter__(self):
        self.thread.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.thread.join(timeout=5)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        import statistics
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows if len(r) > 2 + i)]
        mx = float(self.rows[0][1]) if self.rows[0][1].replace(".", "").isdigit() else None
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": reasons, "samples": len(self.rows)}


def make_inputs(S, seed=0):
    from surs_b200 import synthetic as syn
    return syn.SyntheticCase(S=S, seed=seed)


# --------------------------------------------------------------------------------------------
# CPU baseline: the reference's query path restated with the same torch ops on the host cores
# --------------------------------------------------------------------------------------------
def cpu_query_rate(case, n_points, chunk=50000, threads=None):
    import numpy as np
    import torch
    from oracle import torch_port
    if threads:
        torch.set_num_threads(threads)
    net = torch_port.TorchPort(case)
    res = 512
    rng = np.random.default_rng(0)
    lin = rng.integers(0, res ** 3, n_points)                     # a bounded random sample of the 512^3 grid nodes
    step = 1.0 / res
    pts = np.stack([(lin // (res * res)) * step - 0.5, ((lin // res) % res) * step - 0.5, (lin % res) * step - 0.5]).astype(np.float32)
    net.query(pts[:, :2000])                                     # warm-up
    t0 = time.perf_counter()
    for s in range(0, n_points, chunk):
        net.query(pts[:, s:s + chunk])
    dt = time.perf_counter() - t0
    return n_points / dt, dt, torch.get_num_threads()


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path.  The reference is pure
    Python and cannot travel to the GPU box, so this is the oracle's torch port of it (same ops:
    baddbmm, grid_sample, Conv1d, leaky_relu, sigmoid), all host threads, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    case = make_inputs(args.size)
    n = args.cpu_points
    rates = []
    for i in range(args.warmup + args.steps):
        r, dt, threads = cpu_query_rate(case, n)
        if i >= args.warmup:
            rates.append((r, dt))
    value = sum(r for r, _ in rates) / len(rates)
    ms = 1e3 * sum(dt for _, dt in rates) / len(rates)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "queries/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "dense 512^3 reconstruction, S=%d input" % args.size, "sample": "%d random grid nodes per step, query only" % n},
        "cpu_baseline": {"value": value, "unit": "queries/s", "cores": threads, "kind": "port",
                         "sample": "%d random nodes of the 512^3 grid per step in 50 000-point chunks (torch CPU port of the reference query)" % n},
        "e2e": {"value": value, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from surs_b200 import _capi, parallel
    from surs_b200.lib import sdf as bsdf

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    res = args.resolution
    case = make_inputs(args.size)
    ctx = _capi.Context(dev)
    t = lambda a: torch.from_numpy(a).to(dev)
    ctx.set_weights([t(w) for w in case.mlp_lr[0]], [t(b) for b in case.mlp_lr[1]],
                    [t(w) for w in case.mlp_hr[0]], [t(b) for b in case.mlp_hr[1]],
                    [321, 1024, 512, 256, 128, 1], [322, 1024, 512, 256, 128, 1], [2, 3, 4])
    f_lr_host = torch.from_numpy(case.feat_lr).pin_memory()
    f_hr_host = torch.from_numpy(case.feat_hr).pin_memory()
    ctx.set_features(f_lr_host.to(dev), f_hr_host.to(dev))
    zn, zd = float(case.load_size // 2), float(case.z_size)
    b_min, b_max = np.array([-0.5] * 3), np.array([0.5] * 3)
    mat = bsdf.grid_matrix(res, b_min, b_max)
    prec = {"fp32": _capi.PREC_FP32, "fp16": _capi.PREC_FP16, "fp16x3": _capi.PREC_FP16X3, "fp16r": _capi.PREC_FP16R}[args.precision]
    n_queries = res ** 3

    def step():
        return parallel.reconstruct_slab(ctx, (res, res, res), b_min, b_max, case.calib, zn, zd, mat[:3, :4], precision=prec)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(args.warmup):
        out = step()
    barrier()
    def n_launches():                                    # the LR mesh goes through a sibling context at N > 1 (parallel.py)
        sib = getattr(ctx, "_mc_sibling", None)
        return ctx.launches + (sib.launches if sib is not None else 0)
    launches0 = n_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        e0.record()
        for _ in range(args.steps):
            out = step()
        e1.record()
        barrier()
        ms_total = e0.elapsed_time(e1)
    launches = n_launches() - launches0
    tt = torch.tensor([ms_total], device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms_step = float(tt.item()) / args.steps
    value = n_queries / (ms_step * 1e-3)

    # ---- dominant kernel (the fused query) timed alone with CUDA events on its stream ----------
    lo, hi = parallel.slab_ranges(res, world)[rank]
    hi_h = min(hi + 1, res)
    q_ms = []
    for i in range(3):
        torch.cuda.synchronize(dev)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        vols = ctx.eval_grid((res, res, res), b_min, b_max, case.calib, zn, zd, precision=prec, plane_lo=lo, plane_hi=hi_h)
        b.record()
        torch.cuda.synchronize(dev)
        q_ms.append(a.elapsed_time(b))
        del vols
    q_ms = sorted(q_ms)[1]
    peaks, peak_kind = load_peaks()
    n_slab = (hi_h - lo) * res * res
    achieved = n_slab * FLOP_PER_QUERY / (q_ms * 1e-3) / 1e12
    peak = float(peaks["bf16_tflops_sustained"])
    # dense grids with an axis-aligned calibration take the column-factored kernels (query_col.cu): the
    # products of the weights with the 320 image channels are computed once per (i,j) column, so the
    # EXECUTED tensor-core work is 2 x 1 376 256 MAC per point (layers 1-3 only) + the per-column table
    col_flop = 2 * 2 * (512 * 1024 + 256 * 512 + 128 * 256)
    # fp16x3: every product as A_hi.W_hi + A_lo.W_hi + A_hi.W_lo = three times the tensor-core work
    # (fp16r: the one-pass work; the refinement of a few percent of the nodes is not counted)
    executed_flop = {_capi.PREC_FP16: col_flop, _capi.PREC_FP16R: col_flop, _capi.PREC_FP16X3: 3 * col_flop}.get(prec, FLOP_PER_QUERY)
    executed = n_slab * executed_flop / (q_ms * 1e-3) / 1e12
    roofline = {"bound": "tensor", "kernel": "query_col_kernel (+ col_table_kernel)" if prec != _capi.PREC_FP32 else "query_simt_kernel",
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                # dram__bytes_read.sum + dram__bytes_write.sum of one query_col_kernel launch at 512^3, ncu --set full
                # (profiles/r1_ncu_full_final_col512.csv): 3.04 GB per-column table read + 1.06 GB volumes written
                "traffic": 4.107e9 if (world == 1 and res == 512 and prec == _capi.PREC_FP16) else None,
                "peak_kind": "%s sustained bf16 (kernel timed inside a 0.3-0.7 s step); burst = %.1f" % (peak_kind, float(peaks["bf16_tflops"])),
                "kernel_ms": q_ms, "algorithmic_flop_per_query": FLOP_PER_QUERY,
                "executed_flop_per_query": executed_flop, "executed_tflops": executed, "executed_frac": executed / peak,
                "note": "achieved = ALGORITHMIC FLOPs (SURVEY 8(d): 4 564 998 per query) / kernel time; it can exceed the peak because the "
                        "column factoring removes 40% of the MACs (exact refactoring, not skipped work). executed_* counts the MACs the "
                        "tensor cores really ran (fp16 operands, fp32 accumulate = the bf16 rate)"}

    # ---- the split-operand mode beside the headline (N = 1): same step, SURS_PREC_FP16X3 -----------------
    accurate = None
    if world == 1 and prec == _capi.PREC_FP16 and not args.no_x3:
        def step3():
            return parallel.reconstruct_slab(ctx, (res, res, res), b_min, b_max, case.calib, zn, zd, mat[:3, :4], precision=_capi.PREC_FP16X3)
        step3()
        torch.cuda.synchronize(dev)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(2):
            step3()
        b.record()
        torch.cuda.synchronize(dev)
        ms3 = a.elapsed_time(b) / 2
        # parity of both tensor modes against the fp32 CUDA-core mode on one plane of the grid (res^2 nodes)
        mid = res // 2
        v32 = ctx.eval_grid((res, res, res), b_min, b_max, case.calib, zn, zd, precision=_capi.PREC_FP32, plane_lo=mid, plane_hi=mid + 1)
        v16 = ctx.eval_grid((res, res, res), b_min, b_max, case.calib, zn, zd, precision=_capi.PREC_FP16, plane_lo=mid, plane_hi=mid + 1)
        vx3 = ctx.eval_grid((res, res, res), b_min, b_max, case.calib, zn, zd, precision=_capi.PREC_FP16X3, plane_lo=mid, plane_hi=mid + 1)
        dmax = lambda u, v: max(float((u[0] - v[0]).abs().max()), float((u[1] - v[1]).abs().max()))
        accurate = {"precision": "fp16x3", "value": n_queries / (ms3 * 1e-3), "unit": "queries/s", "ms_per_step": ms3,
                    "executed_tflops": n_queries * 3 * col_flop / (ms3 * 1e-3) / 1e12,
                    "max_abs_diff_vs_fp32_mode": {"fp16x3": dmax(vx3, v32), "fp16": dmax(v16, v32), "nodes": res * res,
                                                  "note": "pre-threshold occupancy, plane %d of the grid; north_star example tolerance 1e-3" % mid}}
        del v32, v16, vx3
        # the refined mode (fp16 everywhere + fp16x3 on the nodes the 0.5 iso-surface can depend on): same mesh as fp16x3
        def step_r():
            return parallel.reconstruct_slab(ctx, (res, res, res), b_min, b_max, case.calib, zn, zd, mat[:3, :4], precision=_capi.PREC_FP16R)
        step_r()
        torch.cuda.synchronize(dev)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(2):
            step_r()
        b.record()
        torch.cuda.synchronize(dev)
        msr = a.elapsed_time(b) / 2
        accurate["refined_mode"] = {"precision": "fp16r", "value": n_queries / (msr * 1e-3), "unit": "queries/s", "ms_per_step": msr,
                                    "refined_fraction": ctx.refined_nodes / float(n_queries),
                                    "note": "marching-cubes output bit-identical to fp16x3 (tests/test_gpu_fullsize.py)"}

    # ---- end to end through the public API with host buffers ------------------------------------
    e2e = None
    if rank == 0 and world == 1 and not args.no_e2e:
        from surs_b200.lib import mesh_util
        from surs_b200.lib.model import SuRSNet
        import types
        opt = types.SimpleNamespace(num_views=1, no_residual=False, mlp_dim_lr=[321, 1024, 512, 256, 128, 1],
                                    mlp_dim_hr=[322, 1024, 512, 256, 128, 1], mlp_res_layers_lr=[2, 3, 4],
                                    mlp_res_layers_hr=[2, 3, 4], loadSize=case.load_size, z_size=case.z_size, threshold=0.05)
        net = SuRSNet(opt, precision=prec).to(dev).eval()
        for mlp, wb in ((net.mlp_lr, case.mlp_lr), (net.mlp_hr, case.mlp_hr)):
            for conv, w, b in zip(mlp.layers(), wb[0], wb[1]):
                conv.weight.data = t(w)[:, :, None].contiguous()
                conv.bias.data = t(b)
        calib = torch.from_numpy(case.calib)[None].to(dev)

        def e2e_step():
            net.im_feat_list_lr = [f_lr_host.to(dev, non_blocking=True)[None]]
            net.im_feat_list_hr = [f_hr_host.to(dev, non_blocking=True)[None]]
            return mesh_util.reconstruction(opt, net, dev, calib, res, b_min, b_max, use_octree=False)

        for _ in range(3):          # warm-up: two generations of pinned staging buffers get cached
            r = e2e_step()
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        n_e2e = max(1, min(args.steps, 3))
        for _ in range(n_e2e):
            r = e2e_step()
        torch.cuda.synchronize(dev)
        dt = (time.perf_counter() - t0) / n_e2e
        d2h = sum(int(a.nbytes) for a in r)
        h2d = int(f_lr_host.numel() * 4 + f_hr_host.numel() * 4 + 64)
        e2e = {"value": n_queries / dt, "unit": "queries/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "s_per_mesh": dt, "verts_hr": int(r[0].shape[0]), "faces_hr": int(r[1].shape[0]),
               "verts_lr": int(r[4].shape[0]), "faces_lr": int(r[5].shape[0])}

    if world > 1 and not args.no_e2e:
        # N > 1: the host-facing distributed call -- every rank uploads the feature maps from pinned host memory,
        # reconstructs its slab, the meshes are gathered on rank 0 and copied to host numpy
        def e2e_step_n():
            return parallel.reconstruction_from_host(ctx, f_lr_host, f_hr_host, (res, res, res), b_min, b_max, case.calib, zn, zd,
                                                     mat[:3, :4], precision=prec)
        for _ in range(3):
            r = e2e_step_n()
        barrier()
        t0 = time.perf_counter()
        n_e2e = max(1, min(args.steps, 3))
        for _ in range(n_e2e):
            r = e2e_step_n()
        barrier()
        te = torch.tensor([(time.perf_counter() - t0) / n_e2e], device=dev, dtype=torch.float64)
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        dt = float(te.item())
        up = torch.tensor([ctx.last_upload_bytes], device=dev, dtype=torch.int64)
        dist.all_reduce(up)
        if rank == 0:
            e2e = {"value": n_queries / dt, "unit": "queries/s",
                   "h2d_bytes_per_step": int(up.item()),
                   "d2h_bytes_per_step": sum(int(a.nbytes) for a in r), "s_per_mesh": dt,
                   "note": "every rank uploads the stripe of the two feature maps its slab samples (h2d = sum over ranks); meshes gathered over NCCL, host copy on rank 0",
                   "verts_hr": int(r[0].shape[0]), "faces_hr": int(r[1].shape[0]), "verts_lr": int(r[4].shape[0]), "faces_lr": int(r[5].shape[0])}

    # ---- CPU baseline (rank 0, N = 1 only) --------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        rate, dt, threads = cpu_query_rate(case, args.cpu_points)
        cpu = {"value": rate, "unit": "queries/s", "cores": threads, "kind": "port",
               "sample": "%d random nodes of the 512^3 grid in 50 000-point chunks through the torch CPU port of the reference "
                         "query (%.1f s); marching cubes timed separately (mc_*)" % (args.cpu_points, dt)}
        # the other half of the reference's CPU path (skimage marching cubes, lib/mesh_util.py:40,45): the oracle's
        # scalar C twin on the 128^3 volumes of BASELINE config 1, one host thread like skimage's Cython loop
        from oracle import mc_oracle
        v128 = ctx.eval_grid((128, 128, 128), b_min, b_max, case.calib, zn, zd, precision=prec)
        mc_s = 0.0
        for v in v128:
            vh = v.cpu().numpy()
            t0 = time.perf_counter()
            mc_oracle.marching_cubes_lewiner(vh, 0.5)
            mc_s += time.perf_counter() - t0
        cpu["mc_port_s_both_volumes_128cubed"] = mc_s
        cpu["s_per_mesh_128cubed_estimate"] = 128 ** 3 / rate + mc_s
        cpu["s_per_mesh_512cubed_estimate"] = 512 ** 3 / rate + mc_s * 16      # surface cells scale with R^2

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "queries/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": {"fp16": "f16", "fp16x3": "f16x3", "fp16r": "f16+f16x3", "fp32": "f32"}[args.precision], "data": "synthetic",
            "config": {"workload": "dense %d^3 reconstruction (query + marching cubes of HR and LR volumes%s), one synthetic %dx%d input, "
                                   "random-init MLP weights" % (res, ", slab-sharded + NCCL mesh gather" if world > 1 else "", args.size, args.size),
                       "resolution": res, "input_side": args.size, "precision": args.precision,
                       "l2": "inputs+outputs per step exceed L2 (features %.0f MB fp16, volumes %.0f MB written); no explicit flush"
                             % ((case.feat_lr.size + case.feat_hr.size) * 2 / 1e6, 2 * n_queries * 4 / 1e6)},
            "s_per_mesh": ms_step * 1e-3,
            "gpu_launches": launches,
            "clocks": clocks.summary(),
            "roofline": roofline,
            "cpu_baseline": cpu,
            "e2e": e2e,
            "accurate_mode": accurate,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--resolution", type=int, default=512)
    ap.add_argument("--size", type=int, default=512, help="side of the synthetic low-res input image")
    ap.add_argument("--precision", default="fp16", choices=["fp16", "fp16x3", "fp16r", "fp32"],
                    help="fp16: one tensor-core pass (headline); fp16r: fp16 + fp16x3 on the nodes the iso-surface depends on; fp16x3: split hi/lo operands, three passes (|d occ| ~2e-5); fp32: CUDA cores")
    ap.add_argument("--cpu-points", type=int, default=800000, help="bounded CPU sample: ~10 s of the torch port on 16 cores")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-x3", action="store_true", help="skip the fp16x3 (split-operand) measurement beside the headline")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
