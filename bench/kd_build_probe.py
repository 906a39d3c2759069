"""k-d tree build timings: rank-based (distinct coordinates) vs. the quickselect-faithful build (ties)."""
import time
import numpy as np
import torch
from grid_interpolation_b200 import interpolation as ip

def t(points, faithful, reps=3):
    p = torch.as_tensor(points, device="cuda")
    ip.kd_build(p, faithful=faithful); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        ip.kd_build(p, faithful=faithful)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3

rng = np.random.default_rng(0)
for n in (5_000, 100_000, 1_000_000):
    pts = rng.uniform(0, 4095.0, (2, n))
    print(f"n={n}: distinct rank-based {t(pts, False):.2f} ms, faithful(forced) {t(pts, True):.2f} ms, "
          f"float32-cast (auto) {t(pts.astype(np.float32).astype(np.float64), False):.2f} ms, "
          f"rounded to 0.01 (auto) {t(np.round(pts, 2), False):.2f} ms", flush=True)
for side in (100, 316, 1000):
    gx, gy = np.meshgrid(np.arange(side, dtype=np.float64), np.arange(side, dtype=np.float64))
    pts = np.stack([gx.ravel(), gy.ravel()])[:, rng.permutation(side * side)]
    print(f"lattice {side}x{side} (auto): {t(pts, False, 1):.1f} ms", flush=True)
