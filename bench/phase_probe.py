"""Tuning helper: time the esrg / kd / bilinear phases at a given scale (device-resident)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from grid_interpolation_b200 import interpolation as ip
import workloads
which = sys.argv[1]; scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
dev = torch.device("cuda:0"); t = lambda a: torch.as_tensor(np.ascontiguousarray(a), device=dev)
if which == "c3":
    g = int(8192 * scale) // 32 * 32; w = workloads.c3(int(10_000_000 * scale * scale), g)
    args = (t(w.points), t(w.data_pts), t(w.x_axis), t(w.y_axis), 1.0, 8)
    out = torch.empty((g, g), dtype=torch.float64, device=dev)
    fn = lambda: ip.idw_esrg_nearest_ex(*args, out=out, sync=False)
elif which == "c2":
    g = int(4096 * scale) // 32 * 32; w = workloads.c2(int(1_000_000 * scale * scale), g)
    args = (t(w.points).t(), t(w.data_pts), t(w.x_axis), t(w.y_axis), 8)
    out = torch.empty((g, g), dtype=torch.float64, device=dev)
    fn = lambda: ip.idw_kdtree_ex(*args, out=out, sync=False)
else:
    g = int(16384 * scale) // 32 * 32; n = int(100_000_000 * scale * scale); w = workloads.c5(n, g)
    args = (t(w.x_axis), t(w.y_axis), t(w.field), torch.as_tensor(np.stack([w.px, w.py], 0), device=dev).t())
    out = torch.empty(n, dtype=torch.float64, device=dev)
    fn = lambda: ip.bilinear(*args, out=out, sync=False)
for _ in range(2): fn()
torch.cuda.synchronize()
ip.set_profiling(True)
for _ in range(3): fn()
torch.cuda.synchronize()
ph = {}
for name, ms in ip.last_timings(): ph.setdefault(name, []).append(ms)
print(which, scale, "  ".join(f"{k} {np.mean(v):.3f} ms" for k, v in ph.items()), "checksum", float(out.double().sum()))
