"""Tuning helper (not part of the product): C4 barycentric phase times, round-1 kernel (every row tested,
GI_BARY_ROW_TESTS) against the segment kernel, both memory orders; outputs must be identical.
`python bench/walk_sweep.py [scale]`"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from grid_interpolation_b200 import interpolation as ip  # noqa: E402
import workloads  # noqa: E402

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
n = int(5_000_000 * scale ** 2); grid = int(16384 * scale) // 128 * 128
w = workloads.c4(n, grid)
dev = torch.device("cuda:0")
t = lambda a: torch.as_tensor(np.ascontiguousarray(a), device=dev)
args = (t(w.points), t(w.data_pts), t(w.x_axis), t(w.y_axis), t(w.simplices), t(w.neighbours))
ref = None
for order in ("C", "F"):
    for row_tests in (False,):
        out = torch.empty((grid, grid), dtype=torch.float64, device=dev)
        out = out if order == "C" else out.t()
        for _ in range(3):
            ip.barycentric_interpolation_ex(*args, out=out, sync=False)
        torch.cuda.synchronize()
        ip.set_profiling(True)
        for _ in range(8):
            ip.barycentric_interpolation_ex(*args, out=out, sync=False)
        torch.cuda.synchronize()
        ph = {}
        for name, ms in ip.last_timings():
            ph.setdefault(name, []).append(ms)
        ip.set_profiling(False)
        if ref is None:
            ref = out.clone()
        same = bool(torch.equal(ref, out))
        print(f"order {order} row_tests={row_tests}: " + "  ".join(f"{k} {np.mean(x):.3f} ms" for k, x in ph.items())
              + f"  identical={same}", flush=True)
