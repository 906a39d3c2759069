import numpy as np, torch
from grid_interpolation_b200 import interpolation as ip
import workloads
dev = torch.device("cuda:0")
w = workloads.c4(); g = 16384
for R, dt in ((np.float32, torch.float32), (np.float64, torch.float64)):
    t = lambda a, d=dt: torch.as_tensor(np.ascontiguousarray(a), device=dev).to(d)
    a = (t(w.points), t(w.data_pts), t(w.x_axis), t(w.y_axis), t(w.simplices, torch.int32), t(w.neighbours, torch.int32))
    out = torch.empty((g, g), dtype=dt, device=dev)
    for _ in range(2): ip.barycentric_interpolation_ex(*a, out=out, sync=False, dtype=R)
    torch.cuda.synchronize()
    ip.set_profiling(True)
    for _ in range(3): ip.barycentric_interpolation_ex(*a, out=out, sync=False, dtype=R)
    torch.cuda.synchronize()
    ph = {}
    for name, ms in ip.last_timings(): ph.setdefault(name, []).append(ms)
    print(np.dtype(R).name, "  ".join(f"{k} {np.mean(v):.3f}" for k, v in ph.items()), flush=True)
    ip.set_profiling(False)
    res, d = ip.barycentric_interpolation_ex(*a, dtype=R, debug=True, rows=(4000, 4512), halo=8)
    print("  debug (512 rows + halo 8):", {k: v for k, v in vars(d).items() if not hasattr(v, "shape")}, flush=True)
    del a, out
