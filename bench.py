#!/usr/bin/env python
"""Benchmark of the SuRS reconstruction hot path on B200 (driver contract: see the task statement).

    python bench.py --gpus N --steps K --warmup W            # ours (N > 1: launched under torchrun)
    python bench.py --impl reference --steps K --warmup W    # the reference's own CPU path (oracle/_ref)

A step = one dense 512^3 reconstruction of one synthetic 512x512 input (BASELINE.json configs[2]):
occupancy query of all 134 217 728 grid nodes (projection + bilinear indexing + both MLPs) and
marching cubes of both volumes, in the DEFAULT precision (SURS_PREC_FP16R: one tensor-core pass +
split-operand refinement of every node the iso-surface depends on; |d occ| <= 1e-3 where marching
cubes reads, see `parity`).  At N > 1 the grid is slab-sharded over the ranks (fixed total work:
strong scaling) and the mesh lists are gathered on rank 0 over NCCL.  Features and weights are
resident in HBM when the timed region starts.  `value` = grid nodes / step time.

`e2e` = the same reconstruction through the public API lib.mesh_util.reconstruction(...) with
HOST buffers on both sides: the two feature maps come from pinned host memory every step
(H2D + repack inside the timed region) and the eight mesh arrays end in host numpy (D2H).

Beside the headline the line carries (N = 1 unless noted): `parity` (the tolerance check of this very
run), `roofline` (tensor pipe, the fused query kernel), `roofline_mc` and `roofline_octree` (HBM),
`configs` (BASELINE configs 2, 4, 5: octree 256^3 / 512^3 and the raw query sweep; config 5 at every N),
`mesh_sha256` (every N; at N > 1 checked against a single-GPU recomputation on rank 0), `cpu_baseline`.
"""
import argparse
import hashlib
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
PRECISIONS = ("fp16r", "fp16x3", "fp16", "fp32")


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def load_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full captures
    (profiles/traffic.json, written by scripts/ncu_summary.py from the .ncu-rep of the same command)."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f)
    return {}


class ClockSampler:
    """SM clock / throttle reasons sampled during the timed region: NVML in-process (every 20 ms, so that even the
    1-2 s timed regions at N >= 4 get dozens of samples), nvidia-smi as the fallback."""

    def __init__(self, index):
        self.rows = []
        self.stop = threading.Event()
        self.index = index
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index(index))
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    @staticmethod
    def _physical_index(index):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if index < len(ids) and ids[index].isdigit():
                return int(ids[index])
        return index

    def _run(self):
        if self.nvml is not None:
            n = self.nvml
            bits = {"hw_slowdown": getattr(n, "nvmlClocksEventReasonHwSlowdown", 0x8),
                    "hw_thermal_slowdown": getattr(n, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                    "sw_thermal_slowdown": getattr(n, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                    "sw_power_cap": getattr(n, "nvmlClocksEventReasonSwPowerCap", 0x4)}
            get_reasons = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
            while not self.stop.is_set():
                try:
                    sm = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
                    r = int(get_reasons(self.handle))
                    self.rows.append((sm, self.max_sm, [k for k, b in bits.items() if r & b]))
                except Exception:
                    pass
                self.stop.wait(0.02)
            return
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                r = [s.strip() for s in out.split(",")]
                if len(r) >= 6:
                    self.rows.append((float(r[0]), float(r[1]), [n for i, n in enumerate(names) if r[2 + i].lower().startswith("active")]))
            except Exception:
                pass
            self.stop.wait(0.1)

    def __enter__(self):
        self.thread.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.thread.join(timeout=5)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"], "samples": 0}
        import statistics
        reasons = sorted({r for row in self.rows for r in row[2]})
        return {"sm_mhz": statistics.median(r[0] for r in self.rows), "sm_max_mhz": self.rows[0][1], "reasons": reasons,
                "samples": len(self.rows), "source": "nvml" if self.nvml is not None else "nvidia-smi"}


def make_inputs(S, seed=0):
    from surs_b200 import synthetic as syn
    return syn.SyntheticCase(S=S, seed=seed)


def make_config(args):
    """The workload both arms are quoted on (identical dict in the reference arm's line)."""
    res, n = args.resolution, args.gpus
    feat_mb = (256 * (args.size // 2) ** 2 + 64 * (2 * args.size) ** 2) * 2 / 1e6
    return {"workload": "dense %d^3 reconstruction (query + marching cubes of HR and LR volumes%s), one synthetic %dx%d input, "
                        "random-init MLP weights" % (res, ", slab-sharded + NCCL mesh gather" if n > 1 else "", args.size, args.size),
            "resolution": res, "input_side": args.size, "precision": args.precision,
            "l2": "inputs+outputs per step exceed L2 (features %.0f MB fp16, volumes %.0f MB written); no explicit flush" % (feat_mb, 2 * res ** 3 * 4 / 1e6)}


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def mesh_sha256(arrays):
    import numpy as np
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


# --------------------------------------------------------------------------------------------
# CPU legs: the reference's own reconstruction (oracle/_ref, unmodified) or, without it, the torch port
# --------------------------------------------------------------------------------------------
def cpu_query_rate(case, n_points, chunk=50000, threads=None):
    """The oracle's torch port of the reference's query path on a bounded random sample of the 512^3 grid nodes."""
    import numpy as np
    import torch
    from oracle import torch_port
    torch.set_num_threads(int(threads) if threads else host_threads())
    net = torch_port.TorchPort(case)
    res = 512
    rng = np.random.default_rng(0)
    lin = rng.integers(0, res ** 3, n_points)
    step = 1.0 / res
    pts = np.stack([(lin // (res * res)) * step - 0.5, ((lin // res) % res) * step - 0.5, (lin % res) * step - 0.5]).astype(np.float32)
    net.query(pts[:, :2000])                                     # warm-up
    t0 = time.perf_counter()
    for s in range(0, n_points, chunk):
        net.query(pts[:, s:s + chunk])
    dt = time.perf_counter() - t0
    return n_points / dt, dt, torch.get_num_threads()


def reference_available():
    from oracle import ref_runner
    return ref_runner.available()


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on the box's host cores.

    When oracle/_ref (the verbatim copy of the reference's modules, oracle/make_ref.py) is present, a step is ONE call of
    the unmodified lib.mesh_util.reconstruction(opt, net, cpu, calib, R_s, b_min, b_max, use_octree=False,
    num_samples=50000) -- create_grid, batch_eval in 50 000-point chunks, query_mr + query_sr, marching cubes of both
    volumes, world transform -- at the bounded sample resolution R_s (--ref-resolution, default 64) so that K + W steps
    end within a few minutes; BASELINE config 1 (R = 128) is timed once on top and reported as `config1`.  Without
    oracle/_ref: the oracle's torch port of the query on random grid nodes (kind "port")."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    threads = host_threads()
    os.environ["OMP_NUM_THREADS"] = str(threads)                 # torchrun exports OMP_NUM_THREADS=1 for nproc > 1
    torch.set_num_threads(threads)
    case = make_inputs(args.size)
    config1 = None
    if reference_available():
        from oracle import ref_runner
        run = ref_runner.ReferenceRun(case, threads=threads)
        rs = args.ref_resolution
        n = rs ** 3
        for _ in range(min(args.warmup, 2)):                     # warm-up at a small size: thread pool, oneDNN primitives
            run.reconstruction(32)
        times = []
        t_budget = time.perf_counter()
        for i in range(args.steps):
            _, dt = run.reconstruction(rs)
            times.append(dt)
        ms = 1e3 * sum(times) / len(times)
        value = n / (ms * 1e-3)
        kind = "reference"
        sample = ("one call of the UNMODIFIED reference lib.mesh_util.reconstruction per step: dense %d^3 grid (%d queries in 50 000-point "
                  "chunks through SuRSNet.query_mr + query_sr on the CPU, S=%d features) + marching cubes of both volumes (%s) + world "
                  "transform" % (rs, n, args.size, "scikit-image" if run.mc == "skimage" else "scikit-image absent: oracle/mc_oracle.c, one thread"))
        if not args.no_config1 and time.perf_counter() - t_budget < 400:
            out, dt1 = run.reconstruction(128)
            config1 = {"workload": "BASELINE configs[0]: reference reconstruction on CPU, resolution 128 dense eval_grid + marching cubes",
                       "s_per_mesh": dt1, "queries_per_s": 128 ** 3 / dt1, "verts_hr": int(out[0].shape[0]), "faces_hr": int(out[1].shape[0]),
                       "marching_cubes": run.mc, "cores": run.threads}
        threads = run.threads
    else:
        n = args.cpu_points
        rates = []
        for i in range(min(args.warmup, 1) + args.steps):
            r, dt, threads = cpu_query_rate(case, n, threads=threads)
            if i >= min(args.warmup, 1):
                rates.append((r, dt))
        value = sum(r for r, _ in rates) / len(rates)
        ms = 1e3 * sum(dt for _, dt in rates) / len(rates)
        kind = "port"
        sample = "%d random nodes of the 512^3 grid per step in 50 000-point chunks (torch CPU port of the reference query; oracle/_ref absent)" % n
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "queries/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": make_config(args),
        "cpu_baseline": {"value": value, "unit": "queries/s", "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "config1": config1,
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------
def _timed(fn, dev, reps=1):
    """Median-free device timing of fn() with CUDA events on torch's current stream (where libsurs launches)."""
    import torch
    torch.cuda.synchronize(dev)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        out = fn()
    b.record()
    torch.cuda.synchronize(dev)
    return a.elapsed_time(b) / reps, out


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
    PREC = {"fp32": _capi.PREC_FP32, "fp16": _capi.PREC_FP16, "fp16x3": _capi.PREC_FP16X3, "fp16r": _capi.PREC_FP16R}
    prec = PREC[args.precision]
    n_queries = res ** 3
    peaks, peak_kind = load_peaks()
    peak_tc = float(peaks["bf16_tflops_sustained"])
    peak_hbm = float(peaks["hbm_gbs"])
    traffic = load_traffic()

    def step(p=prec):
        return parallel.reconstruct_slab(ctx, (res, res, res), b_min, b_max, case.calib, zn, zd, mat[:3, :4], precision=p)

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
    refine = ctx.refine_stats if prec == _capi.PREC_FP16R else None
    tt = torch.tensor([ms_total], device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms_step = float(tt.item()) / args.steps
    value = n_queries / (ms_step * 1e-3)

    # ---- mesh identity: sha256 of (verts_hr, faces_hr, verts_lr, faces_lr) of the timed step's result; at N > 1 rank 0
    # recomputes the whole grid alone (outside the timed region) and the gathered mesh must be bit-identical ------------
    mesh = None
    if rank == 0:
        arrs = [out[0][0], out[0][1], out[1][0], out[1][1]]
        mesh = {"sha256": mesh_sha256([a.cpu().numpy() for a in arrs]),
                "verts_hr": int(arrs[0].shape[0]), "faces_hr": int(arrs[1].shape[0]), "verts_lr": int(arrs[2].shape[0]), "faces_lr": int(arrs[3].shape[0])}
        if world > 1:
            vols = ctx.eval_grid((res, res, res), b_min, b_max, case.calib, zn, zd, precision=prec)
            single = []
            for vol in vols:
                nv, nf, _ = ctx.mc_count(vol, 0.5)
                _, wv, _, _ = ctx.mc_emit_verts(nv, mat[:3, :4], want_normals=False)
                single += [wv, ctx.mc_emit_faces(nf)]
            del vols
            mesh["single_gpu_sha256"] = mesh_sha256([a.cpu().numpy() for a in single])
            mesh["matches_single_gpu"] = mesh["single_gpu_sha256"] == mesh["sha256"]
            del single
    del out

    # ---- dominant kernel: the fused query of the slab, timed alone with CUDA events on its stream -------------------
    lo, hi = (parallel.current_slab_ranges(ctx, res, world) if world > 1 else parallel.slab_ranges(res, world))[rank]
    hi_h = min(hi + 1, res)
    n_slab = (hi_h - lo) * res * res

    def grid_ms(p):
        ts = []
        for _ in range(3):
            ms, vols = _timed(lambda: ctx.eval_grid((res, res, res), b_min, b_max, case.calib, zn, zd, precision=p, plane_lo=lo, plane_hi=hi_h), dev)
            del vols
            ts.append(ms)
        return sorted(ts)[1]
    q_ms = grid_ms(prec)
    # dense grids with an axis-aligned calibration take the column-factored kernels (query_col.cu): the products of the
    # weights with the 320 image channels are computed once per (i,j) column, so the EXECUTED tensor-core work is
    # 2 x 1 376 256 MAC per point (layers 1-3 only) + the per-column table; fp16x3 = three products per MAC
    col_flop = 2 * 2 * (512 * 1024 + 256 * 512 + 128 * 256)
    one_ms = grid_ms(_capi.PREC_FP16) if prec == _capi.PREC_FP16R else None
    refined_frac = (refine["nodes"] / float(n_slab)) if refine else 0.0
    lr_only_frac = (refine["nodes_lr_mlp_only"] / float(n_slab)) if refine else 0.0
    executed_flop = {_capi.PREC_FP16: col_flop, _capi.PREC_FP16R: col_flop * (1.0 + 3.0 * (refined_frac - 0.5 * lr_only_frac)),
                     _capi.PREC_FP16X3: 3 * col_flop}.get(prec, FLOP_PER_QUERY)
    achieved = n_slab * FLOP_PER_QUERY / (q_ms * 1e-3) / 1e12
    executed = n_slab * executed_flop / (q_ms * 1e-3) / 1e12
    tkey = "query_%s_%d" % (args.precision, res)
    roofline = {"bound": "tensor",
                "kernel": {"fp16r": "query_col_kernel<one pass> + refine_select + col_table_kernel<split> + query_col_kernel<indexed, split>",
                           "fp16": "query_col_kernel (+ col_table_kernel)", "fp16x3": "query_col_kernel<split> (+ col_table_kernel<split>)",
                           "fp32": "query_simt_kernel"}[args.precision],
                "achieved": achieved, "peak": peak_tc, "unit": "TFLOP/s", "frac": achieved / peak_tc,
                "traffic": traffic.get(tkey) if world == 1 else None,
                "peak_kind": "%s sustained bf16 (kernel timed inside a 0.3-0.7 s step); burst = %.1f" % (peak_kind, float(peaks["bf16_tflops"])),
                "kernel_ms": q_ms, "algorithmic_flop_per_query": FLOP_PER_QUERY,
                "executed_flop_per_query": executed_flop, "executed_tflops": executed, "executed_frac": executed / peak_tc,
                "note": "achieved = ALGORITHMIC FLOPs (SURVEY 8(d): 4 564 998 per query) x slab nodes / device time of the whole grid evaluation "
                        "(every kernel of surs_eval_grid in this precision). It can exceed the peak because the column factoring removes 40% of "
                        "the MACs (exact refactoring, not skipped work); executed_* counts the MACs the tensor cores really ran (fp16 operands, "
                        "fp32 accumulate = the bf16 rate; refined nodes three more products each)"}
    if world > 1:                                         # load balance of the slabs: every rank's grid-evaluation time
        allq = torch.zeros(world, device=dev)
        allq[rank] = q_ms
        dist.all_reduce(allq)
        roofline["per_rank_grid_ms"] = [round(float(v), 2) for v in allq.tolist()]
        roofline["slab_planes"] = [hi_ - lo_ for lo_, hi_ in parallel.current_slab_ranges(ctx, res, world)]
        roofline["slab_balance"] = ("adaptive: slab sizes follow the per-rank grid-evaluation times of the previous step (SURS_BALANCE=0: equal slabs); "
                                    "the gathered mesh is identical for any partition (mesh.matches_single_gpu)")
    if one_ms is not None:
        roofline["one_pass_kernel_ms"] = one_ms
        roofline["one_pass_executed_frac"] = n_slab * col_flop / (one_ms * 1e-3) / 1e12 / peak_tc
        roofline["refinement_ms"] = q_ms - one_ms
        roofline["refined_fraction"] = refined_frac

    # ---- marching cubes alone (HBM bound): both volumes of this rank's slab ------------------------------------------
    vols = ctx.eval_grid((res, res, res), b_min, b_max, case.calib, zn, zd, precision=prec, plane_lo=lo, plane_hi=hi_h)

    def mc_both():
        nb = 0
        for vol in vols:
            nv, nf, _ = ctx.mc_count(vol, 0.5)
            ctx.mc_emit_verts(nv, mat[:3, :4])
            ctx.mc_emit_faces(nf)
            nb += vol.numel() * 4 + nv * (12 + 24 + 12 + 4) + nf * 12
        return nb
    mc_both()
    mc_ms, mc_bytes = _timed(mc_both, dev)
    mc_gbs = mc_bytes / (mc_ms * 1e-3) / 1e9
    roofline_mc = {"bound": "hbm", "kernel": "mc_* (sign, bits, cell, scans, list_verts, list_faces), both volumes", "achieved": mc_gbs, "peak": peak_hbm,
                   "unit": "GB/s", "frac": mc_gbs / peak_hbm, "ms": mc_ms, "algorithmic_bytes": mc_bytes,
                   "traffic": traffic.get("mc_%d" % res) if world == 1 else None,
                   "note": "algorithmic bytes = 4 B x nodes read once + per vertex 12 B index coords + 24 B world coords (float64) + 12 B normal + 4 B "
                           "value + 12 B per face, both volumes; time includes the two host synchronisations of surs_mc_count"}
    del vols

    # ---- the other BASELINE configs (device-timed, default precision unless stated) ----------------------------------
    configs = {}
    if not args.no_configs:
        def octree(r):
            def run():
                a, b, n_eval = ctx.eval_grid_octree((r, r, r), b_min, b_max, case.calib, zn, zd, 0.05, precision=prec)
                st = ctx.octree_stats()
                m = bsdf.grid_matrix(r, b_min, b_max)[:3, :4]
                counts = []
                for vol in (a, b):                                  # float64 volumes: cast inside the marching-cubes bit pass
                    nv, nf, _ = ctx.mc_count(vol, 0.5)
                    ctx.mc_emit_verts(nv, m)
                    ctx.mc_emit_faces(nf)
                    counts.append((nv, nf))
                return n_eval, st, counts
            run()
            ms, (n_eval, st, counts) = _timed(run, dev)
            # HBM roofline of the bookkeeping (K3: select + cell pass + the initialisation the reference's semantics need)
            by = (r ** 3 * 17                                               # zero both float64 volumes, dirty = 1
                  + st["candidates"] * 1 + st["evaluated"] * (8 + 1)        # select: dirty read, index written, dirty cleared
                  + st["cells"] * 1 + st["cells_read"] * (16 * 8 + 8)       # decide: centre flag, 16 corners, parked values
                  + (st["filled_hr"] + st["filled_lr"]) * 8 + max(st["filled_hr"], st["filled_lr"]) * 1)   # fill
            k3_ms = st["init_ms"] + st["select_ms"] + st["cells_ms"]
            return {"ms_per_mesh": ms, "evaluated": n_eval, "evaluated_fraction": n_eval / float(r ** 3), "queries_per_s": n_eval / (st["query_ms"] * 1e-3),
                    "query_frac_of_tensor_peak": n_eval * FLOP_PER_QUERY / (st["query_ms"] * 1e-3) / 1e12 / peak_tc,
                    "phases_ms": {k: st[k] for k in ("init_ms", "table_ms", "select_ms", "query_ms", "cells_ms")},
                    "verts_hr": counts[0][0], "faces_hr": counts[0][1], "verts_lr": counts[1][0], "faces_lr": counts[1][1],
                    "bookkeeping": {"bound": "hbm", "achieved": by / (k3_ms * 1e-3) / 1e9, "peak": peak_hbm, "unit": "GB/s",
                                    "frac": by / (k3_ms * 1e-3) / 1e9 / peak_hbm, "ms": k3_ms, "algorithmic_bytes": by}}
        if world == 1:
            configs["c2_octree_256"] = octree(256)
            configs["c4_octree_512_one_image"] = octree(512)
        # config 5: raw query throughput, points sharded over the ranks (no communication), device resident
        sweep = {}
        for prec_name, sizes in ((args.precision, (20, 22, 24)), ("fp16", (20, 23, 26))):
            if prec_name == "fp32":
                continue
            for lg in sizes:
                n_all = 1 << lg
                n_mine = n_all // world
                g = torch.Generator(device=dev).manual_seed(1000 + rank)
                pts = torch.rand((3, n_mine), device=dev, generator=g) - 0.5
                ctx.query(pts, case.calib, zn, zd, precision=PREC[prec_name])          # warm-up at full size: scratch tables are sized by it
                ms, _ = _timed(lambda: ctx.query(pts, case.calib, zn, zd, precision=PREC[prec_name]), dev)
                tm = torch.tensor([ms], device=dev)
                if world > 1:
                    dist.all_reduce(tm, op=dist.ReduceOp.MAX)
                ms = float(tm.item())
                sweep["%s_2^%d" % (prec_name, lg)] = {"points": n_all, "ms": ms, "queries_per_s": n_all / (ms * 1e-3),
                                                      "frac_of_tensor_peak": n_all * FLOP_PER_QUERY / (ms * 1e-3) / 1e12 / (peak_tc * world)}
                del pts
        configs["c5_query_sweep"] = sweep
    roofline_octree = configs.get("c4_octree_512_one_image", {}).get("bookkeeping")

    # ---- parity of THIS run's precision against the fp32 mode (N = 1): the north_star's numbers ----------------------
    parity = None
    if world == 1 and rank == 0 and prec != _capi.PREC_FP32:
        planes = [res // 2 - 40, res // 2, res // 2 + 40] if res >= 128 else [res // 2]
        full = ctx.eval_grid((res, res, res), b_min, b_max, case.calib, zn, zd, precision=prec)
        rep = {"max_abs_all": 0.0, "max_abs_where_mc_reads": 0.0, "nodes": 0, "nodes_mc_reads": 0, "flips": 0, "near_threshold_1e-3": 0,
               "flips_outside_1e-3": 0}
        for pl in planes:
            a, b = max(pl - 1, 0), min(pl + 2, res)
            ref = ctx.eval_grid((res, res, res), b_min, b_max, case.calib, zn, zd, precision=_capi.PREC_FP32, plane_lo=a, plane_hi=b)
            for got, want in zip(full, ref):
                got = got[a:b]
                inside = want > 0.5
                edge = torch.zeros_like(inside)
                for d in range(3):
                    m = inside.shape[d] - 1
                    diff = inside.narrow(d, 1, m) != inside.narrow(d, 0, m)
                    edge.narrow(d, 1, m).logical_or_(diff)
                    edge.narrow(d, 0, m).logical_or_(diff)
                sel = slice(pl - a, pl - a + 1)                              # the middle plane has all six neighbours
                dd = (got[sel] - want[sel]).abs()
                fl = (got[sel] > 0.5) != inside[sel]
                near = (want[sel] - 0.5).abs() < 1e-3
                rep["max_abs_all"] = max(rep["max_abs_all"], float(dd.max()))
                if bool(edge[sel].any()):
                    rep["max_abs_where_mc_reads"] = max(rep["max_abs_where_mc_reads"], float(dd[edge[sel]].max()))
                rep["nodes"] += int(dd.numel())
                rep["nodes_mc_reads"] += int(edge[sel].sum())
                rep["flips"] += int(fl.sum())
                rep["near_threshold_1e-3"] += int(near.sum())
                rep["flips_outside_1e-3"] += int((fl & ~near).sum())
            del ref
        del full
        rep["tolerance"] = 1e-3
        rep["meets_tolerance"] = bool(rep["max_abs_where_mc_reads"] <= 1e-3 and rep["flips_outside_1e-3"] == 0)
        rep["note"] = ("pre-threshold occupancy of this run's precision vs the fp32 CUDA-core mode (itself within 2e-5 of the reference's fp32 "
                       "PyTorch path, tests/) on planes %s of the grid, HR and LR; `where_mc_reads` = nodes with an inside/outside change to a "
                       "6-neighbour, the only nodes whose VALUE marching cubes uses" % planes)
        if refine:
            rep["refinement"] = refine
        parity = rep

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
               "verts_lr": int(r[4].shape[0]), "faces_lr": int(r[5].shape[0]),
               "mesh_sha256": mesh_sha256([r[0], r[1], r[4], r[5]])}

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
                   "note": "every rank uploads the stripe of the two feature maps its slab samples (h2d = sum over ranks) and copies its part of the "
                           "meshes over its own PCIe link into a pinned shared-memory host arena (d2h = sum over ranks); rank 0 returns zero-copy views",
                   "verts_hr": int(r[0].shape[0]), "faces_hr": int(r[1].shape[0]), "verts_lr": int(r[4].shape[0]), "faces_lr": int(r[5].shape[0]),
                   "mesh_sha256": mesh_sha256([r[0], r[1], r[4], r[5]])}

    # ---- CPU baseline (rank 0, N = 1 only): bounded sample of the same workload on the host cores -------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        threads = host_threads()
        if reference_available():
            from oracle import ref_runner
            run = ref_runner.ReferenceRun(case, threads=threads)
            run.reconstruction(32)
            rs = args.cpu_resolution
            _, dt = run.reconstruction(rs)
            cpu = {"value": rs ** 3 / dt, "unit": "queries/s", "cores": run.threads, "kind": "reference",
                   "sample": "one call of the UNMODIFIED reference lib.mesh_util.reconstruction (oracle/_ref): dense %d^3 grid = %d queries in "
                             "50 000-point chunks on the CPU + marching cubes of both volumes (%s) + world transform, %.1f s"
                             % (rs, rs ** 3, "scikit-image" if run.mc == "skimage" else "scikit-image absent: oracle/mc_oracle.c", dt),
                   "s_per_mesh_sample": dt, "s_per_mesh_512cubed_estimate": dt * (512.0 / rs) ** 3}
        else:
            rate, dt, threads = cpu_query_rate(case, args.cpu_points, threads=threads)
            cpu = {"value": rate, "unit": "queries/s", "cores": threads, "kind": "port",
                   "sample": "%d random nodes of the 512^3 grid in 50 000-point chunks through the torch CPU port of the reference "
                             "query (%.1f s); oracle/_ref absent" % (args.cpu_points, dt)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "queries/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": {"fp16": "f16", "fp16x3": "f16x3", "fp16r": "f16+f16x3", "fp32": "f32"}[args.precision], "data": "synthetic",
            "config": make_config(args),
            "s_per_mesh": ms_step * 1e-3,
            "gpu_launches": launches,
            "clocks": clocks.summary(),
            "roofline": roofline,
            "roofline_mc": roofline_mc,
            "roofline_octree": roofline_octree,
            "parity": parity,
            "mesh": mesh,
            "cpu_baseline": cpu,
            "e2e": e2e,
            "configs": configs,
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
    ap.add_argument("--precision", default="fp16r", choices=list(PRECISIONS),
                    help="fp16r (default, meets the 1e-3 tolerance where marching cubes reads): one tensor-core pass + split-operand refinement; "
                         "fp16x3: split hi/lo operands everywhere (|d occ| < 1e-4); fp32: CUDA cores; fp16: one pass, opt-in, NOT a parity mode")
    ap.add_argument("--cpu-points", type=int, default=800000, help="torch-port CPU sample (only when oracle/_ref is absent)")
    ap.add_argument("--cpu-resolution", type=int, default=96, help="cpu_baseline: resolution of the one reference reconstruction() call (~10 s)")
    ap.add_argument("--ref-resolution", type=int, default=64, help="--impl reference: resolution of each step's reference reconstruction() call")
    ap.add_argument("--no-config1", action="store_true", help="--impl reference: skip the extra 128^3 reconstruction (BASELINE configs[0])")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip BASELINE configs 2, 4, 5 beside the headline")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
