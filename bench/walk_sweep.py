"""Tuning helper (not part of the product): time the C4 barycentric phases for each GI_WALK_VARIANT."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from grid_interpolation_b200 import interpolation as ip  # noqa: E402
from grid_interpolation_b200 import workloads  # noqa: E402

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
variants = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0, 1, 2, 3, 4]
n = int(5_000_000 * scale ** 2); grid = int(16384 * scale) // 128 * 128
w = workloads.c4(n, grid)
dev = torch.device("cuda:0")
t = lambda a: torch.as_tensor(np.ascontiguousarray(a), device=dev)
args = (t(w.points), t(w.data_pts), t(w.x_axis), t(w.y_axis), t(w.simplices), t(w.neighbours))
out = torch.empty((grid, grid), dtype=torch.float64, device=dev)
ref = None
for order in ("C",):
    for v in variants:
        os.environ["GI_WALK_VARIANT"] = str(v)  # (the kernel no longer has variants; kept for re-use)
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
        print(f"variant {v}: " + "  ".join(f"{k} {np.mean(x):.3f} ms" for k, x in ph.items()) + f"  identical={same}",
              flush=True)
