"""C1 small-call breakdown: wall time per call for numpy in/out, device tensors in/out, a plan; phase times of one
call; used with `ncu --metrics gpu__time_duration.sum` for the launch list of one call."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from grid_interpolation_b200 import interpolation as ip
import workloads
w = workloads.c1()
pf = np.asfortranarray(w.points); sf = np.asfortranarray(w.simplices); nf = np.asfortranarray(w.neighbours)
def t(fn, n=200):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3
one = len(sys.argv) > 1 and sys.argv[1] == "one"
field = ip.idw_esrg_nearest(w.points, w.data_pts, w.x_axis, w.y_axis, w.grid_spac, 6)
if one:
    ip.barycentric_interpolation(pf, w.data_pts, w.x_axis, w.y_axis, sf, nf)
    ip.bilinear(w.x_axis, w.y_axis, field, pf)
    sys.exit(0)
dev = [torch.as_tensor(np.ascontiguousarray(a), device="cuda:0") for a in (w.points, w.data_pts, w.x_axis, w.y_axis, w.simplices, w.neighbours)]
dfield = torch.as_tensor(field, device="cuda:0")
print("barycentric numpy   %.3f ms" % t(lambda: ip.barycentric_interpolation(pf, w.data_pts, w.x_axis, w.y_axis, sf, nf)))
print("barycentric device  %.3f ms" % t(lambda: ip.barycentric_interpolation(*dev)))
print("bilinear numpy      %.3f ms" % t(lambda: ip.bilinear(w.x_axis, w.y_axis, field, pf)))
print("bilinear device     %.3f ms" % t(lambda: ip.bilinear(dev[2], dev[3], dfield, dev[0])))
print("esrg numpy          %.3f ms" % t(lambda: ip.idw_esrg_nearest(pf, w.data_pts, w.x_axis, w.y_axis, w.grid_spac, 6)))
print("kdtree numpy        %.3f ms" % t(lambda: ip.idw_kdtree(w.points.T, w.data_pts, w.x_axis, w.y_axis, 6)))
ip.set_profiling(True)
ip.barycentric_interpolation(pf, w.data_pts, w.x_axis, w.y_axis, sf, nf)
print("phases numpy ", ip.last_timings())
ip.set_profiling(True)
ip.barycentric_interpolation(*dev); torch.cuda.synchronize()
print("phases device", ip.last_timings())
ip.set_profiling(False)
# pure python / ctypes overhead: time of an H2D + D2H of the same sizes through torch
h = torch.empty(132 * 79, dtype=torch.float64).pin_memory()
d = torch.empty(132 * 79, dtype=torch.float64, device="cuda:0")
def cp():
    d.copy_(h, non_blocking=True); h.copy_(d, non_blocking=True); torch.cuda.synchronize()
print("torch h2d+d2h+sync  %.3f ms" % t(cp))
