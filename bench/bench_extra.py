"""bench_extra.py -- the non-headline configs of BASELINE.json (C2 k-d tree IDW, C3 esrg IDW, C5 bilinear) with
the same JSON contract as bench.py (which dispatches here for --workload c2|c3|c5).  Parity-test cases in the
first place; the lines printed here are for DESIGN.md / BASELINE.md, not the driver's headline."""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import workloads  # noqa: E402


def _kd_launches(n):
    """Kernel launches of one idw_kdtree call (mirrors kdtree.cu: two radix sorts + level-synchronous build)."""
    sort = 1 + 8 * 3 + 1  # keys + 8 passes x (histogram, scan, scatter) + tie detection
    levels = 0
    while (1 << levels) <= n:
        levels += 1
    scan = 1 if n <= 16 * 4096 else (3 if n <= 16 * 4096 * 4096 else 5)
    return 2 * sort + 1 + levels * (2 + scan) + 1 + 1 + 1  # init, levels, gather, query, flush


def _phases(ip):
    agg = {}
    for name, v in ip.last_timings():
        agg.setdefault(name, []).append(v)
    return {k: float(np.mean(v)) for k, v in agg.items()}


def run(args, rank, world, local_rank):
    import bench as B
    wl = args.workload
    sc = args.scale
    if wl == "c2":
        grid = max(int(4096 * sc) // 32 * 32, 64); n = max(int(1_000_000 * sc * sc), 1000)
        w = workloads.c2(n, grid)
    elif wl == "c3":
        grid = max(int(8192 * sc) // 32 * 32, 64); n = max(int(10_000_000 * sc * sc), 1000)
        w = workloads.c3(n, grid)
    else:
        grid = max(int(16384 * sc) // 32 * 32, 64); n = max(int(100_000_000 * sc * sc), 1000)
        w = workloads.c5(n, grid)
    k = 8

    # ------------------------------------------------------------------ CPU sample (oracle port)
    def cpu_sample(seconds):
        from oracle import oracle as orc
        cores = orc.get_max_threads()
        if wl == "c5":
            m = min(n, 2_000_000)
            pts = np.asfortranarray(np.stack([w.px[:m], w.py[:m]], 1))
            fld = np.asfortranarray(w.field)
            t0 = time.perf_counter(); orc.bilinear(w.x_axis, w.y_axis, fld, pts); dt = time.perf_counter() - t0
            m2 = int(min(n, max(m, m / dt * seconds)))
            pts = np.asfortranarray(np.stack([w.px[:m2], w.py[:m2]], 1))
            t0 = time.perf_counter(); orc.bilinear(w.x_axis, w.y_axis, fld, pts); dt = time.perf_counter() - t0
            return m2 / dt, cores, f"first {m2} of the {n} points ({dt:.1f} s)"
        rows = 4 * cores
        desc = ""
        for _ in range(2):
            rows = int(min(grid, max(rows, cores)))
            r0 = (grid - rows) // 2
            ys = w.y_axis[r0:r0 + rows]
            if wl == "c3":  # esrg needs an equally spaced grid: a contiguous band + the points within 64 cells of it
                sel = (w.points[:, 1] > ys[0] - 64) & (w.points[:, 1] < ys[-1] + 64)
                pts = np.asfortranarray(w.points[sel]); dat = w.data_pts[sel]
                t0 = time.perf_counter()
                orc.idw_esrg_nearest(pts, dat, w.x_axis, ys, 1.0, k)
            else:
                pts = np.asfortranarray(w.points.T)
                t0 = time.perf_counter()
                orc.idw_kdtree(pts, w.data_pts, w.x_axis, ys, k)
            dt = time.perf_counter() - t0
            desc = f"{rows} contiguous rows x {grid} columns around the grid centre ({rows * grid} targets, {dt:.1f} s)"
            rate = rows * grid / dt
            rows = int(rate * seconds / grid)
        return rate, cores, desc

    name = {"c2": f"C2 IDW k=8 via k-d tree: {n} points -> {grid} x {grid} grid",
            "c3": f"C3 IDW k=8 via esrg cell list: {n} points -> {grid} x {grid} grid",
            "c5": f"C5 bilinear: {grid} x {grid} grid -> {n} scattered points"}[wl]

    if args.impl == "reference":
        if rank != 0:
            return 0
        vals = []
        for i in range(args.warmup + args.steps):
            v, cores, desc = cpu_sample(min(args.cpu_seconds, 150.0 / max(args.steps, 1)) if i >= args.warmup else 1.0)
            if i >= args.warmup:
                vals.append(v)
        value = float(np.mean(vals))
        print(json.dumps({"metric": "target points interpolated/sec", "value": value, "unit": "targets/s",
                          "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": None,
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                          "data": "synthetic", "impl": "reference", "config": {"workload": name},
                          "cpu_baseline": {"value": value, "unit": "targets/s", "cores": cores, "kind": "port",
                                           "sample": desc},
                          "e2e": {"value": value, "unit": "targets/s", "h2d_bytes_per_step": 0,
                                  "d2h_bytes_per_step": 0}, "gpu_launches": 0}))
        return 0

    import torch
    import torch.distributed as dist
    from grid_interpolation_b200 import interpolation as ip
    from grid_interpolation_b200 import _lib
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), device=dev)

    if wl == "c5":
        lo, hi = rank * n // world, (rank + 1) * n // world          # contiguous point range per rank
        pts_soa = np.stack([w.px[lo:hi], w.py[lo:hi]], 0)            # x block, y block = the reference's (n,2) F-order
        pts_h = pts_soa.T                                             # F-ordered (n,2) view, no copy
        targets_rank = hi - lo
        d_x, d_y, d_f, d_p = t(w.x_axis), t(w.y_axis), t(w.field), torch.as_tensor(pts_soa, device=dev).t()
        d_out = torch.empty(targets_rank, dtype=torch.float64, device=dev)
        step_dev = lambda: ip.bilinear(d_x, d_y, d_f, d_p, out=d_out, sync=False)
        h_args = (ip.pinned_copy(w.x_axis), ip.pinned_copy(w.y_axis), ip.pinned_copy(w.field),
                  ip.pinned_copy(pts_h))
        h_out = ip.pinned_empty((targets_rank,), np.float64)
        step_host = lambda: ip.bilinear(*h_args, out=h_out)
        launches = 2
        dominant = "bilinear"
    else:
        band = (rank * grid // world, (rank + 1) * grid // world)
        rows = band[1] - band[0]
        targets_rank = rows * grid
        d_out = torch.empty((rows, grid), dtype=torch.float64, device=dev)
        h_out = ip.pinned_empty((rows, grid), np.float64, "C")
        if wl == "c2":
            pts = np.ascontiguousarray(w.points)                      # (n,2) C-order == interleaved == F-order (2,n)
            d_p, d_v, d_x, d_y = torch.as_tensor(pts, device=dev).t(), t(w.data_pts), t(w.x_axis), t(w.y_axis)
            step_dev = lambda: ip.idw_kdtree_ex(d_p, d_v, d_x, d_y, k, rows=band, out=d_out, sync=False)
            h_p = ip.pinned_copy(pts).T
            h_args = (h_p, ip.pinned_copy(w.data_pts), ip.pinned_copy(w.x_axis), ip.pinned_copy(w.y_axis))
            step_host = lambda: ip.idw_kdtree_ex(*h_args, k, rows=band, out=h_out)
            launches = _kd_launches(n)
            dominant = "kd_query"
        else:
            d_p, d_v, d_x, d_y = t(w.points), t(w.data_pts), t(w.x_axis), t(w.y_axis)
            step_dev = lambda: ip.idw_esrg_nearest_ex(d_p, d_v, d_x, d_y, 1.0, k, rows=band, out=d_out, sync=False)
            h_args = (ip.pinned_copy(w.points), ip.pinned_copy(w.data_pts), ip.pinned_copy(w.x_axis),
                      ip.pinned_copy(w.y_axis))
            step_host = lambda: ip.idw_esrg_nearest_ex(*h_args, 1.0, k, rows=band, out=h_out)
            cells = grid * grid
            passes = 1
            while passes < 4 and (cells - 1) >> (8 * passes):
                passes += 1
            # keys, radix passes x (histogram, scan, scatter), table, gather, tile query, exact-order query, flush
            launches = 1 + 3 * passes + 2 + 2 + 1
            dominant = "esrg_query"

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step_dev()
    sync_all()
    stream = torch.cuda.current_stream().cuda_stream
    _lib.check(_lib.lib().gi_stream_status(stream), "warm-up")
    clocks = B.ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    ip.set_profiling(True)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    ev0.record()
    for _ in range(args.steps):
        step_dev()
    ev1.record()
    sync_all()
    phase_ms = _phases(ip)
    ip.set_profiling(False)
    clk = clocks.stop() if rank == 0 else None
    _lib.check(_lib.lib().gi_stream_status(stream), "timed steps")
    ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    value = targets_rank * world * args.steps / (ms_total * 1e-3)

    # steady state with a plan (search structure prepared once; extension, not the headline)
    plan_line = None
    if wl != "c5":
        if wl == "c2":
            plan = ip.KdTreePlan(d_p, d_x, d_y, k, rows=band, order="C")
        else:
            plan = ip.EsrgPlan(d_p, d_x, d_y, 1.0, k, rows=band, order="C")
        p_out = torch.empty_like(d_out)
        for _ in range(3):
            plan(d_v, out=p_out, sync=False)
        sync_all()
        pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        pe0.record()
        for _ in range(args.steps):
            plan(d_v, out=p_out, sync=False)
        pe1.record()
        sync_all()
        pms = torch.tensor([pe0.elapsed_time(pe1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(pms, op=dist.ReduceOp.MAX)
        assert torch.equal(p_out, d_out), "plan result differs from the one-shot result"
        plan_line = {"ms_per_step": float(pms.item()) / args.steps,
                     "value": targets_rank * world * args.steps / (float(pms.item()) * 1e-3), "unit": "targets/s"}
        plan.close()

    # C5 with the points in the Morton order of their cells (north_star: Morton-ordered targets, Morton ranges per
    # GPU).  The order comes from the library (timed); the points are moved once, as an ingest step would (not
    # timed); the kernel is timed on the ordered points; the un-permutation of the results (torch index_put) is
    # timed separately.  Not the headline: `value` above is the reference's call on the points as they come.
    morton = None
    if wl == "c5":
        d_all = torch.as_tensor(np.stack([w.px, w.py], 0), device=dev).t()      # every rank orders the whole set
        def timed(fn, reps):
            fn(); torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(reps):
                r = fn()
            b.record(); torch.cuda.synchronize()
            return a.elapsed_time(b) / reps, r
        order_ms, order = timed(lambda: ip.bilinear_morton_order(d_x, d_y, d_all, sync=False), 3)
        mine = order[lo:hi].long()                                               # this rank's Morton range
        d_pm = d_all[mine].t().contiguous().t()                                  # (n, 2) in the reference's F order
        del d_all
        m_out = torch.empty(targets_rank, dtype=torch.float64, device=dev)
        sync_all()
        k_ms, _ = timed(lambda: ip.bilinear(d_x, d_y, d_f, d_pm, out=m_out, sync=False), args.steps)
        kms_t = torch.tensor([k_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(kms_t, op=dist.ReduceOp.MAX)
        back = torch.empty(n if world == 1 else targets_rank, dtype=torch.float64, device=dev)
        idx = mine if world == 1 else torch.argsort(mine)
        u_ms, _ = timed(lambda: back.index_copy_(0, idx, m_out) if world == 1 else back.copy_(m_out[idx]), 3)
        if world == 1:
            assert torch.equal(back, d_out), "Morton-ordered run, un-permuted, differs from the direct run"
        morton = {"order_ms": order_ms, "kernel_ms": float(kms_t.item()), "unpermute_ms": u_ms,
                  "value_kernel_only": targets_rank * world / (float(kms_t.item()) * 1e-3), "unit": "targets/s",
                  "note": "points pre-ordered along the Z curve of 16 x 16-node blocks (gi_bilinear_morton_order); "
                          "ranks take contiguous ranges of that order"}
        del d_pm, m_out, back, order, mine, idx
        torch.cuda.empty_cache()

    e2e = None
    if not args.no_e2e:
        e_steps = max(2, min(args.steps, 5))
        step_host()
        sync_all()
        t0 = time.perf_counter()
        for _ in range(e_steps):
            step_host()
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        h2d = sum(np.asarray(a).nbytes for a in h_args)
        e2e = {"value": targets_rank * world * e_steps / float(dt.item()), "unit": "targets/s",
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(h_out.nbytes), "steps": e_steps}
        assert np.array_equal(h_out, d_out.cpu().numpy()), "host-path result differs from device-path result"

    if rank == 0:
        peak, peak_src = B.peaks()
        kms = phase_ms.get(dominant)
        roof = None
        if kms:
            alg = B.ALG_BYTES[wl] * targets_rank
            ach = alg / (kms * 1e-3) / 1e9
            roof = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                    "kernel": dominant, "kernel_ms": kms, "peak_source": peak_src, "algorithmic_bytes_per_launch": alg}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            v, cores, desc = cpu_sample(args.cpu_seconds)
            cpu = {"value": v, "unit": "targets/s", "cores": cores, "kind": "port", "sample": desc}
        print(json.dumps({"metric": "target points interpolated/sec", "value": value, "unit": "targets/s",
                          "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                          "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong",
                          "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "ours",
                          "config": {"workload": name, "l2": "inputs + outputs per step exceed the 126 MB L2",
                                     "scale": sc},
                          "phase_ms": phase_ms, "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "plan": plan_line,
                          "morton": morton,
                          "clocks": clk, "gpu_launches": launches * args.steps}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0
