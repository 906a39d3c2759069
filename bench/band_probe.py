"""Phase times of ONE row band of C4 on one GPU (what a rank of the strong-scaling run executes): whole-mesh pack
vs band-limited pack, per band position.  `python bench/band_probe.py [bands] [scale]`"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from grid_interpolation_b200 import interpolation as ip  # noqa: E402
import workloads

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 8
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
grid = int(16384 * scale) // 128 * 128
w = workloads.c4(int(5_000_000 * scale ** 2), grid)
t = lambda a: torch.as_tensor(np.ascontiguousarray(a), device="cuda:0")
mesh = [t(a) for a in (w.points, w.data_pts, w.x_axis, w.y_axis, w.simplices, w.neighbours)]
full = ip.barycentric_interpolation(*mesh, order="C")
for bp in (False, True):
    for b in sorted({0, nb // 2, nb - 1}):
        rows = (b * grid // nb, (b + 1) * grid // nb)
        out = torch.empty((rows[1] - rows[0], grid), dtype=torch.float64, device="cuda:0")
        for _ in range(3):
            ip.barycentric_interpolation_ex(*mesh, rows=rows, halo=8, out=out, band_pack=bp, sync=False)
        torch.cuda.synchronize()
        ip.set_profiling(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            ip.barycentric_interpolation_ex(*mesh, rows=rows, halo=8, out=out, band_pack=bp, sync=False)
        e1.record()
        torch.cuda.synchronize()
        agg = {}
        for name, v in ip.last_timings():
            agg.setdefault(name, []).append(v)
        ip.set_profiling(False)
        from grid_interpolation_b200 import _lib
        st = _lib.lib().gi_stream_status(torch.cuda.current_stream().cuda_stream)
        ok = torch.equal(out, full[rows[0]:rows[1]]) and st == 0
        print(f"band_pack={bp} band {b}/{nb} rows {rows}: {e0.elapsed_time(e1) / 10:.3f} ms/step  "
              + "  ".join(f"{k} {np.mean(v):.3f}" for k, v in agg.items()) + f"  equal={ok}", flush=True)
