"""Single precision (the reference as shipped: every REAL is 32-bit without -fdefault-real-8) at the full C2-C5
sizes, device-resident, next to double precision (the graded kind): ms per call, mean of 5 after 2 warm-ups."""
import sys, time
import numpy as np, torch
from grid_interpolation_b200 import interpolation as ip
import workloads

dev = torch.device("cuda:0")
def t(a, dt): return torch.as_tensor(np.ascontiguousarray(a), device=dev).to(dt)
def timeit(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

which = sys.argv[1:] or ["c2", "c3", "c5", "c4"]
for name in which:
    for R, dt in ((np.float64, torch.float64), (np.float32, torch.float32)):
        if name == "c2":
            w = workloads.c2(); g = 4096
            a = (t(w.points, dt).t(), t(w.data_pts, dt), t(w.x_axis, dt), t(w.y_axis, dt), 8)
            out = torch.empty((g, g), dtype=dt, device=dev)
            fn = lambda: ip.idw_kdtree_ex(*a, out=out, sync=False, dtype=R)
        elif name == "c3":
            w = workloads.c3(); g = 8192
            a = (t(w.points, dt), t(w.data_pts, dt), t(w.x_axis, dt), t(w.y_axis, dt), 1.0, 8)
            out = torch.empty((g, g), dtype=dt, device=dev)
            fn = lambda: ip.idw_esrg_nearest_ex(*a, out=out, sync=False, dtype=R)
        elif name == "c5":
            w = workloads.c5(); n = w.px.size
            a = (t(w.x_axis, dt), t(w.y_axis, dt), t(w.field, dt), torch.stack([t(w.px, dt), t(w.py, dt)], 0).t())
            out = torch.empty(n, dtype=dt, device=dev)
            fn = lambda: ip.bilinear(*a, out=out, sync=False, dtype=R)
        else:
            w = workloads.c4(); g = 16384
            a = (t(w.points, dt), t(w.data_pts, dt), t(w.x_axis, dt), t(w.y_axis, dt), t(w.simplices, torch.int32),
                 t(w.neighbours, torch.int32))
            out = torch.empty((g, g), dtype=dt, device=dev)
            fn = lambda: ip.barycentric_interpolation_ex(*a, out=out, sync=False, dtype=R)
        ms = timeit(fn)
        print(f"{name} {np.dtype(R).name}: {ms:.3f} ms per call", flush=True)
        del a, out, fn
        torch.cuda.empty_cache()
