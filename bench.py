#!/usr/bin/env python
"""bench.py -- headline benchmark of the grid_interpolation hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c4|c2|c3|c5]
                    [--scaling strong|weak] [--scale F]

Workload = BASELINE.json's headline config C4: barycentric interpolation via the triangle walk from a synthetic
~10 M-triangle Delaunay mesh (5 M uniform points, scipy/Qhull, cached under $GI_CACHE) to a 16384 x 16384 regular
grid, hull fill included.  One "step" = one complete, stateless barycentric_interpolation of the WHOLE grid (mesh
packing + seed grid + walk / barycentric + hull fill), as the reference's routine is stateless.

  value     target points/s, inputs resident in HBM, CUDA events around exactly K steps after W warm-ups, max over
            ranks.  N = 1: one call.  N > 1 (torchrun, one rank per GPU; "scaling": "strong"): the SAME 16384^2 grid
            in N row bands (BASELINE.json: "row-band sharded over 8 GPUs"), mesh replicated, every rank packs only
            the triangles near its band, and the final gather of the field is INSIDE the timed step: the kernels
            store every result into the rank's own field and into rank 0's over NVLink (peer stores,
            grid_interpolation_b200.sharding.PeerField), followed by a stream-ordered signal; rank 0's step ends
            when all N bands have arrived in its copy.  Extra keys at N > 1: "gather_all" (the field gathered on
            EVERY rank), "bands_only" (no gather), "gather_nccl" (band + ncclAllGather, the plain-collective
            form), "weak" (round 1's headline: one 16384-row band per rank of a 16384 x 16384N grid),
            "nvlink_floor_ms" (bytes rank 0 must receive / 900 GB/s).
  parity    the run checks itself: N = 1 compares the whole field, the triangle ids, the outside-hull mask and the
            fill sources with the CPU oracle (the restatement of the Fortran, oracle/); N > 1 compares every
            rank's gathered field with a one-shot single-GPU run of the whole grid.
  e2e       the same metric through the reference-facing call with HOST (pinned) buffers.  N = 1: the numpy entry
            point, H2D + D2H inside the timed region.  N > 1: host inputs live on rank 0 only: H2D there, NCCL
            broadcast to the other ranks, band compute, every rank copies its band to its own pinned host buffer.
  roofline  the locate / evaluate kernel: algorithmic bytes (9.34 B/target for C4) / its CUDA-event
            duration, against the measured HBM copy bandwidth in MEASURED_PEAKS.json; "bound" names what limits
            the kernel according to ncu (profiles/).
  cpu_baseline  the CPU oracle (C++/OpenMP restatement of the Fortran, see oracle/) with all host threads on the
            whole grid of the same workload (a strided row sample when that would take longer than --cpu-seconds).

--impl reference times the reference's own CPU algorithm (the oracle port; the Fortran cannot be compiled in this
image: no gfortran) with all the host threads the process may use (not OMP_NUM_THREADS, which torchrun sets to 1).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import workloads  # noqa: E402

ALG_BYTES = {"c2": 9.43, "c3": 11.58, "c4": 9.34, "c5": 45.48}       # SURVEY.md section 8(d), FP64 + int32
DOMINANT = {"c4": "walk", "c5": "bilinear", "c3": "esrg_query", "c2": "kd_query"}
NVLINK_GBS = 900.0                                                     # per direction and GPU (NVLink 5)
# what limits locate_eval_kernel according to the committed ncu capture (profiles/r02_locate_eval_ncu.csv)
C4_LIMITER = {"unit": "dependent walk chain at 8 CTAs/SM: issue slots 55-60 %, L1 wavefronts 62 %, FP64 pipe 33 % (not HBM: DRAM 13 %)",
              "source": "ncu --set full, profiles/r02_final_locate_eval_ncu.csv"}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c4", choices=["c4", "c2", "c3", "c5"])
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"])
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the workload (debug only; <1 is not the headline)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="N > 1: skip bands_only / gather_nccl / weak")
    ap.add_argument("--dest-share", default="auto",
                    help="N > 1: share of the rows computed by the rank the field is gathered on: 'auto' (tuned before "
                         "the timed region), 'even' (1/N) or a number")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target CPU time of the cpu_baseline sample")
    ap.add_argument("--cpu-threads", type=int, default=0, help="threads of the CPU arms (0 = all the process may use)")
    return ap.parse_args()


def host_threads(args):
    """All hardware threads this process may run on -- not omp_get_max_threads(): torchrun exports
    OMP_NUM_THREADS=1, which made round 1's multi-GPU reference arm single-threaded."""
    if args.cpu_threads > 0:
        return args.cpu_threads
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
def make_c4(args, rank, world, dist):
    """The C4 mesh workload and the grid edge.  Rank 0 triangulates (about a minute at full size) and caches."""
    n = max(int(5_000_000 * args.scale ** 2), 1000)
    grid = max(int(16384 * args.scale) // 128 * 128, 128)
    if world > 1:
        if rank == 0:
            w = workloads.c4(n, grid)
        dist.barrier()
        if rank != 0:
            w = workloads.c4(n, grid)
    else:
        w = workloads.c4(n, grid)
    return w, grid


def cpu_oracle_c4(w, y_axis, seconds, threads, want_field=False):
    """Oracle barycentric_interpolation on this job's grid with `threads` OpenMP threads.  A probe on evenly strided
    rows (every row pays the reference's restart-from-the-start-triangle walk, as in the full run) sizes the sample;
    when the whole grid fits the budget the whole grid is computed -- and returned, for the parity check."""
    from oracle import oracle as orc
    orc.set_num_threads(threads)
    pts_f = np.asfortranarray(w.points); simp_f = np.asfortranarray(w.simplices); nbr_f = np.asfortranarray(w.neighbours)
    ny = len(y_axis)

    def run(ys, debug=False):
        t0 = time.perf_counter()
        res = orc.barycentric_interpolation(pts_f, w.data_pts, w.x_axis, ys, simp_f, nbr_f, debug=debug)
        return res, time.perf_counter() - t0

    rows = min(max(8 * threads, 64), ny)
    _, dt = run(np.ascontiguousarray(y_axis[:: max(ny // rows, 1)][:rows]))
    est_full = dt * ny / rows
    if est_full <= 2.0 * seconds:
        res, dt = run(y_axis, debug=want_field)
        tg = ny * w.x_axis.size
        return tg / dt, f"the whole {w.x_axis.size} x {ny} grid ({tg} targets, {dt:.1f} s)", res if want_field else None
    rows = int(min(ny, max(rows, rows * seconds / max(dt, 1e-3))))
    ys = np.ascontiguousarray(y_axis[:: max(ny // rows, 1)][:rows])
    _, dt = run(ys)
    tg = ys.size * w.x_axis.size
    return tg / dt, f"{ys.size} evenly strided rows x {w.x_axis.size} columns of the job's grid ({tg} targets, {dt:.1f} s)", None


def reference_arm(args):
    """--impl reference: the reference's CPU algorithm (oracle port) on the same config, all host threads."""
    w, grid = make_c4(args, 0, 1, None)
    y_axis = w.y_axis if args.scaling == "strong" else np.arange(grid * args.gpus, dtype=np.float64) / args.gpus
    threads = host_threads(args)
    vals, desc = [], ""
    per_step = min(max(args.cpu_seconds, 5.0), 150.0 / max(args.steps + args.warmup, 1))
    for i in range(args.warmup + args.steps):
        v, desc, _ = cpu_oracle_c4(w, y_axis, per_step if i >= args.warmup else 2.0, threads)
        if i >= args.warmup:
            vals.append(v)
    value = float(np.mean(vals))
    print(json.dumps({
        "metric": "target points interpolated/sec", "value": value, "unit": "targets/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": None, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
        "config": {"workload": f"C4 barycentric via triangle walk: {w.simplices.shape[0]} triangles -> "
                               f"{w.x_axis.size} x {len(y_axis)} grid (bounded sample per step)"},
        "cpu_baseline": {"value": value, "unit": "targets/s", "cores": threads, "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": "targets/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))
    return 0


class Timer:
    """K steps between two CUDA events on the current stream, barrier + synchronize on both sides, max over ranks."""

    def __init__(self, torch, dist, world, dev):
        self.torch, self.dist, self.world, self.dev = torch, dist, world, dev

    def sync_all(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def run(self, step, steps, warmup):
        torch = self.torch
        for _ in range(warmup):
            step()
        self.sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        self.sync_all()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(ms, op=self.dist.ReduceOp.MAX)
        return float(ms.item()) / steps


def phase_means(ip):
    agg = {}
    for name, v in ip.last_timings():
        agg.setdefault(name, []).append(v)
    return {k: float(np.mean(v)) for k, v in agg.items()}


def main():
    args = parse()
    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if args.workload != "c4":
        sys.path.insert(0, os.path.join(ROOT, "bench"))
        import bench_extra
        return bench_extra.run(args, rank, world, local_rank)
    if args.impl == "reference":
        return reference_arm(args) if rank == 0 else 0

    import torch
    import torch.distributed as dist
    from grid_interpolation_b200 import _lib, sharding
    from grid_interpolation_b200 import interpolation as ip
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    w, grid = make_c4(args, rank, world, dist if world > 1 else None)
    len_x = w.x_axis.size
    stream = torch.cuda.current_stream().cuda_stream
    tm = Timer(torch, dist, world, dev)
    warm = max(args.warmup, 3)
    halo = sharding.halo_for_fill() if world > 1 else 0
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), device=dev)
    status = lambda what: _lib.check(_lib.lib().gi_stream_status(stream), what)

    # device-resident inputs in the caller's (numpy C-order) layout, replicated on every rank
    d_pts, d_val, d_x = t(w.points), t(w.data_pts), t(w.x_axis)
    d_simp, d_nbr = t(w.simplices), t(w.neighbours)
    extras = {}
    weak = world > 1 and args.scaling == "weak"
    if weak:      # round 1's headline: per-GPU work fixed, no gather
        y_axis = np.arange(grid * world, dtype=np.float64) / world
        band = (rank * grid, (rank + 1) * grid)
    else:
        y_axis = w.y_axis
        band = sharding.row_band(grid, rank, world)
    d_y = t(y_axis)
    rows = band[1] - band[0]
    targets_total = len_x * len(y_axis)
    mesh = (d_pts, d_val, d_x, d_y, d_simp, d_nbr)

    field = None
    if world > 1 and not weak:
        field = sharding.PeerField(grid, len_x, order="C")
        d_out = field.tensor
        # The gathering rank's NVLink ingress is the floor of this step, not its SMs: let it compute a larger band
        # while the other bands arrive (sharding.gather_bands).  The share is tuned here, outside the timed region:
        # a few short runs per candidate, the fastest wins; the even split is reported next to it.
        bands_cur = [sharding.split_range(grid, world)]

        def step_dev():       # the final gather: band -> this rank's field and rank 0's (peer stores); rank 0 waits
            sharding.barycentric_gather(field, *mesh, halo=halo, dest=0, bands=bands_cur[0])
            if rank == 0:
                field.wait()

        even_share = 1.0 / world
        cands = [even_share] + ([] if args.dest_share == "even" else
                                [float(args.dest_share)] if args.dest_share != "auto" else
                                [even_share + j * (0.5 - even_share) / 6 for j in range(1, 7)] if world > 2 else [])
        sweep = []
        for c in cands:
            bands_cur[0] = sharding.gather_bands(grid, world, 0, c) if c != even_share else sharding.split_range(grid, world)
            sweep.append((tm.run(step_dev, 3, 2), c))
            status(f"dest share {c:.3f}")
        best_ms, best_share = min(sweep)
        if args.dest_share == "auto" and best_share != even_share:      # second pass: +- half a step around the best
            for c in (best_share - 0.03, best_share + 0.03):
                bands_cur[0] = sharding.gather_bands(grid, world, 0, c)
                sweep.append((tm.run(step_dev, 3, 2), c))
                status(f"dest share {c:.3f}")
            best_ms, best_share = min(sweep)
        bands_cur[0] = (sharding.gather_bands(grid, world, 0, best_share) if best_share != even_share
                        else sharding.split_range(grid, world))
        band = bands_cur[0][rank]
        rows = band[1] - band[0]
        extras["gather_bands"] = {"dest_share": best_share, "rows_per_rank": [hi - lo for lo, hi in bands_cur[0]],
                                  "sweep_ms": {f"{c:.3f}": ms for ms, c in sweep},
                                  "even_split_ms": sweep[0][0],
                                  "note": "rank 0 (the gather's destination) computes dest_share of the rows while the "
                                          "other bands arrive over NVLink; tuned outside the timed region"}

        def step_all():       # band -> EVERY rank's field, signal, every rank waits for all bands
            sharding.barycentric_gather(field, *mesh, halo=halo)
            field.wait()
    else:
        d_out = torch.empty((rows, len_x), dtype=torch.float64, device=dev)

        def step_dev():
            ip.barycentric_interpolation_ex(*mesh, rows=band, halo=halo, out=d_out, sync=False)

    for _ in range(warm):
        step_dev()
    tm.sync_all()
    status("warm-up")
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    ip.set_profiling(True)
    ms_step = tm.run(step_dev, args.steps, 0)
    phase_ms = phase_means(ip)
    ip.set_profiling(False)
    clk = clocks.stop() if rank == 0 else None
    status("timed steps")
    value = targets_total / (ms_step * 1e-3)
    walk_ms = phase_ms.get("walk")

    # ------------------------------------------------------------------ parity of THIS run's output
    parity = None
    if world > 1 and not weak:
        one = ip.barycentric_interpolation(*mesh, order="C")
        # rank 0 holds the gathered field; the other ranks hold (at least) their own band
        sel = slice(0, grid) if rank == 0 else slice(band[0], band[1])
        bad = int((one[sel] != d_out[sel]).sum().item())
        cnt = torch.tensor([bad], dtype=torch.int64, device=dev)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        parity = {"against": "a one-shot single-GPU run of the whole grid: the gathered field on rank 0, the own "
                             "band on the other ranks", "targets": targets_total, "value_mismatches": int(cnt.item())}
        # the all-to-all form: the field complete on EVERY rank
        ms_a = tm.run(step_all, args.steps, warm)
        status("gather_all")
        bad = int((one != d_out).sum().item())
        cnt = torch.tensor([bad], dtype=torch.int64, device=dev)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        extras["gather_all"] = {"ms_per_step": ms_a, "value": targets_total / (ms_a * 1e-3), "unit": "targets/s",
                                "value_mismatches": int(cnt.item()),
                                "note": "the field gathered on every rank (each receives (N-1)/N of it over NVLink)"}
        del one

    # ------------------------------------------------------------------ N > 1 extras
    if world > 1 and not weak and not args.no_extras:
        band_e = sharding.row_band(grid, rank, world)       # the extras use the even split (comparable across rounds)
        rows_e = band_e[1] - band_e[0]
        d_band = torch.empty((rows_e, len_x), dtype=torch.float64, device=dev)

        def step_band():
            ip.barycentric_interpolation_ex(*mesh, rows=band_e, halo=halo, out=d_band, band_pack=True, sync=False)

        ms_b = tm.run(step_band, args.steps, warm)
        status("bands_only")
        extras["bands_only"] = {"ms_per_step": ms_b, "value": targets_total / (ms_b * 1e-3), "unit": "targets/s",
                                "note": "N even bands without any gather (round 1's strong-scaling figure)"}
        assert torch.equal(d_band, d_out[band_e[0]:band_e[1]]), "band differs from the gathered field"
        gathered = torch.empty((grid, len_x), dtype=torch.float64, device=dev)
        even = all(hi - lo == rows_e for lo, hi in sharding.split_range(grid, world))

        def step_nccl():
            step_band()
            if even:
                dist.all_gather_into_tensor(gathered, d_band)
            else:
                gathered.copy_(sharding.all_gather_field(d_band, grid))

        ms_n = tm.run(step_nccl, args.steps, warm)
        status("gather_nccl")
        assert torch.equal(gathered, d_out), "NCCL-gathered field differs from the peer-store field"
        extras["gather_nccl"] = {"ms_per_step": ms_n, "value": targets_total / (ms_n * 1e-3), "unit": "targets/s",
                                 "note": "band kernels, then ncclAllGather of the bands (the plain-collective form)"}
        del gathered, d_band
        recv = (grid - rows) * len_x * 8
        extras["nvlink_floor_ms"] = recv / (NVLINK_GBS * 1e9) * 1e3
        extras["nvlink_floor_even_split_ms"] = (grid - rows_e) * len_x * 8 / (NVLINK_GBS * 1e9) * 1e3
        extras["gather_bytes_received_per_rank"] = int(recv)
        # weak scaling as in round 1: one grid-row band per rank of a `grid` x `grid`*N grid, no gather
        yw = t(np.arange(grid * world, dtype=np.float64) / world)
        bw = (rank * grid, (rank + 1) * grid)
        d_w = torch.empty((grid, len_x), dtype=torch.float64, device=dev)

        def step_weak():
            ip.barycentric_interpolation_ex(d_pts, d_val, d_x, yw, d_simp, d_nbr, rows=bw, halo=halo, out=d_w, sync=False)

        ms_w = tm.run(step_weak, max(args.steps // 2, 3), warm)
        status("weak")
        extras["weak"] = {"ms_per_step": ms_w, "value": len_x * grid * world / (ms_w * 1e-3), "unit": "targets/s",
                          "note": f"{len_x} x {grid * world} grid over the same mesh extent, one {grid}-row band per "
                                  "rank, no gather"}
        del d_w, yw

    # ------------------------------------------------------------------ steady state with a plan (extension)
    plan_line = None
    if world == 1:
        plan = ip.BarycentricPlan(d_pts, d_x, d_y, d_simp, d_nbr, order="C")
        p_out = torch.empty_like(d_out)
        ms_p = tm.run(lambda: plan(d_val, out=p_out, sync=False), args.steps, 3)
        assert torch.equal(p_out, d_out), "plan result differs from the one-shot result"
        plan_line = {"ms_per_step": ms_p, "value": targets_total / (ms_p * 1e-3), "unit": "targets/s",
                     "note": "mesh packed once (BarycentricPlan), one execute per step; not the headline"}
        plan.close()
        del p_out

    # ------------------------------------------------------------------ opt-in plane evaluation (extension)
    # GI_BARY_PLANE_VALUES: same triangles / mask / fill sources, values within 1e-12 * max|data| of the reference's
    # instead of bit-identical; NOT the headline (the headline is the bit-identical default)
    plane_line = None
    d_plane = None
    if world == 1:
        d_plane = torch.empty_like(d_out)
        ip.set_profiling(True)
        ms_pl = tm.run(lambda: ip.barycentric_interpolation_ex(*mesh, order="C", out=d_plane, plane_values=True,
                                                               sync=False), args.steps, 3)
        ph_pl = phase_means(ip)
        ip.set_profiling(False)
        scale_d = float(np.abs(w.data_pts).max())
        plane_line = {"ms_per_step": ms_pl, "value": targets_total / (ms_pl * 1e-3), "unit": "targets/s",
                      "phase_ms": ph_pl,
                      "max_abs_diff_vs_default_over_max_data": float((d_plane - d_out).abs().max().item()) / scale_d,
                      "tolerance": 1e-12,
                      "note": "plane_values=True (GI_BARY_PLANE_VALUES): value = d3 + gx*(x-x3) + gy*(y-y3) per triangle; "
                              "opt-in, not the headline"}

    # ------------------------------------------------------------------ mesh ingest (extension)
    # gi_mesh_reorder: the same mesh numbered along the Z curve of position (once, at ingest); the stateless call on
    # the ingested mesh must give the identical field.  NOT the headline (the headline mesh is in Qhull's order).
    ingest_line = None
    if world == 1:
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        ip.mesh_reorder(d_pts, d_simp, d_nbr, sync=False)          # warm-up (scratch, kernels)
        torch.cuda.synchronize()
        e0.record()
        ing = ip.mesh_reorder(d_pts, d_simp, d_nbr, sync=False)
        e1.record(); torch.cuda.synchronize()
        ms_ing = e0.elapsed_time(e1)
        val_ing = d_val[(ing.point_order - 1).long()]
        d_ing = torch.empty_like(d_out)
        ip.set_profiling(True)
        ms_im = tm.run(lambda: ip.barycentric_interpolation_ex(ing.points, val_ing, d_x, d_y, ing.simplices, ing.neighbours,
                                                               order="C", out=d_ing, sync=False), args.steps, 3)
        ph_im = phase_means(ip)
        ip.set_profiling(False)
        ingest_line = {"ms_per_step": ms_im, "value": targets_total / (ms_im * 1e-3), "unit": "targets/s",
                       "phase_ms": ph_im, "reorder_ms": ms_ing,
                       "value_mismatches_vs_default": int((d_ing != d_out).sum().item()),
                       "note": "stateless call on the mesh after gi_mesh_reorder (triangles and points along the Z "
                               "curve; reorder_ms once at ingest, outside the step); not the headline"}
        del d_ing, ing, val_ing

    # ------------------------------------------------------------------ e2e: host buffers through the public API
    e2e = None
    if not args.no_e2e:
        e_steps = max(2, min(args.steps, 5))
        host_in = [w.points, w.data_pts, w.x_axis, y_axis, w.simplices, w.neighbours]
        h2d = sum(a.nbytes for a in host_in)
        band_h = band if (world == 1 or weak) else sharding.row_band(grid, rank, world)   # e2e: even bands (every rank's PCIe link busy)
        h_out = ip.pinned_empty((band_h[1] - band_h[0], len_x), np.float64, "C")
        if world == 1:
            h_in = [ip.pinned_copy(a) for a in host_in]

            def step_host():
                ip.barycentric_interpolation_ex(*h_in, rows=band, halo=halo, out=h_out)
            note = ("pinned host buffers; H2D + pack + walk + fill + D2H inside the timed region; the call copies "
                    "output bands back while the next band is computed")
        else:
            h_in = [ip.pinned_copy(a) for a in host_in] if rank == 0 else None
            meta = [(a.shape, a.dtype) for a in host_in]
            t_out = torch.from_numpy(h_out)
            d_b = torch.empty((band_h[1] - band_h[0], len_x), dtype=torch.float64, device=dev)

            def step_host():   # rank 0 owns the host inputs: H2D there, NVLink broadcast, band, D2H per rank
                m = sharding.broadcast_inputs(h_in, src=0, device=dev, meta=meta)
                ip.barycentric_interpolation_ex(*m, rows=band_h, halo=halo, out=d_b, band_pack=not weak, sync=False)
                t_out.copy_(d_b, non_blocking=True)
                torch.cuda.synchronize()
            note = ("host inputs on rank 0 only: H2D there + NCCL broadcast of the mesh, band compute, every rank "
                    "copies its band into its own pinned host buffer (the field stays distributed over the ranks' "
                    "host memory)")
        step_host()
        tm.sync_all()
        t0 = time.perf_counter()
        for _ in range(e_steps):
            step_host()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        status("e2e")
        e2e = {"value": targets_total * e_steps / float(dt.item()), "unit": "targets/s",
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(targets_total * 8), "steps": e_steps,
               "note": note}
        ref_rows = d_out[band_h[0]:band_h[1]] if field is not None else d_out
        assert np.array_equal(h_out, ref_rows.cpu().numpy()), "host-path result differs from device-path result"

    # ------------------------------------------------------------------ CPU oracle: baseline + parity (N = 1)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        threads = host_threads(args)
        v, desc, res = cpu_oracle_c4(w, y_axis, args.cpu_seconds, threads, want_field=True)
        cpu = {"value": v, "unit": "targets/s", "cores": threads, "kind": "port", "sample": desc}
        if res is not None:
            want, o = res
            got, g = ip.barycentric_interpolation_ex(*mesh, order="C", debug=True)
            assert torch.equal(got, d_out), "debug call differs from the timed call"
            o_mask = np.ascontiguousarray(o.mask)
            ins = ~o_mask
            parity = {"against": "the CPU oracle (restatement of the Fortran) on the whole grid",
                      "targets": int(want.size),
                      "value_mismatches": int((got.cpu().numpy() != want).sum()),
                      "mask_mismatches": int((g.mask.cpu().numpy() != o_mask).sum()),
                      "triangle_mismatches": int((g.tri.cpu().numpy()[ins] != np.ascontiguousarray(o.tri)[ins]).sum()),
                      "fill_source_mismatches": int((g.fill_src.cpu().numpy() != np.ascontiguousarray(o.fill_src)).sum()),
                      "masked_targets": int(o_mask.sum())}
            if plane_line is not None:   # the opt-in form against the oracle: ids bit for bit, values to 1e-12 * max|data|
                gp, gg = ip.barycentric_interpolation_ex(*mesh, order="C", debug=True, plane_values=True)
                assert torch.equal(gp, d_plane), "debug call differs from the timed call (plane values)"
                err = np.abs(gp.cpu().numpy() - want).max() / float(np.abs(w.data_pts).max())
                plane_line["parity"] = {
                    "against": "the CPU oracle on the whole grid", "max_abs_diff_over_max_data": float(err),
                    "within_tolerance": bool(err <= 1e-12),
                    "mask_mismatches": int((gg.mask.cpu().numpy() != o_mask).sum()),
                    "triangle_mismatches": int((gg.tri.cpu().numpy()[ins] != np.ascontiguousarray(o.tri)[ins]).sum()),
                    "fill_source_mismatches": int((gg.fill_src.cpu().numpy() != np.ascontiguousarray(o.fill_src)).sum())}
                del gp, gg
            del want, o, got, g

    if field is not None:
        d_out = None
        field.close()
    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    peak, peak_src = peaks()
    alg = ALG_BYTES["c4"] * rows * len_x
    roof = None
    if walk_ms:
        achieved = alg / (walk_ms * 1e-3) / 1e9
        roof = {"bound": "issue", "roofline": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": None, "limited_by": C4_LIMITER,
                "kernel": "locate_eval_kernel<double>", "kernel_ms": walk_ms, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg}
        try:   # ncu figure of the same kernel / config; only quoted for the launch it was measured on
            with open(os.path.join(ROOT, "profiles", "walk_kernel_traffic.json")) as f:
                tj = json.load(f)
            if world == 1 and args.scale == 1.0:
                roof["traffic"] = tj.get("dram_bytes_per_launch")
                roof["traffic_source"] = tj.get("source")
        except Exception:
            pass

    # kernels of ours per step: pack (+ need), locate/evaluate, start triangle, fix-up, row walk, fill, status
    # flush (+ signal, wait and its flush in the gather form)
    launches = 7 + ((1 + 3) if field is not None else 0)
    band_txt = (f"{world} row bands, the field gathered on rank 0 by peer stores inside the step"
                if world > 1 and not weak else f"{world} row band(s) of {rows} rows")
    out = {
        "metric": "target points interpolated/sec", "value": value, "unit": "targets/s", "n_gpus": world,
        "steps": args.steps, "warmup": warm, "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "weak" if weak else "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "impl": "ours",
        "config": {"workload": f"C4 barycentric via triangle walk + hull fill: {w.points.shape[0]} points / "
                               f"{w.simplices.shape[0]} triangles (scipy Delaunay, default_rng(4)) -> "
                               f"{len_x} x {len(y_axis)} regular grid, {band_txt}",
                   "l2": "inputs (0.36 GB) and output (2.1 GB) per step exceed the 126 MB L2; no explicit flush",
                   "layout": "numpy C-order inputs and output, consumed in place (no transposition)",
                   "scale": args.scale},
        "phase_ms": phase_ms, "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "parity": parity, "plan": plan_line,
        "plane_values": plane_line, "ingested_mesh": ingest_line,
        "clocks": clk, "gpu_launches": launches * args.steps,
    }
    out.update(extras)
    print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
