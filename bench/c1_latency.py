"""C1 (the reference's own test_interpolation.py case, 5 000 points -> 132 x 79 grid): per-call latency of the four
entry points through the numpy API (host buffers, H2D/D2H included) next to the CPU oracle.  Latency-bound."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from grid_interpolation_b200 import interpolation as ip
import workloads
from oracle import oracle as orc
w = workloads.c1()
def t(fn, n=50):
    fn(); fn()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    return (time.perf_counter() - t0) / n * 1e3
pf = np.asfortranarray(w.points); sf = np.asfortranarray(w.simplices); nf = np.asfortranarray(w.neighbours)
field = ip.idw_esrg_nearest(w.points, w.data_pts, w.x_axis, w.y_axis, w.grid_spac, 6)
rows = [
 ("barycentric_interpolation", lambda: ip.barycentric_interpolation(pf, w.data_pts, w.x_axis, w.y_axis, sf, nf),
  lambda: orc.barycentric_interpolation(pf, w.data_pts, w.x_axis, w.y_axis, sf, nf)),
 ("idw_esrg_nearest k=6", lambda: ip.idw_esrg_nearest(pf, w.data_pts, w.x_axis, w.y_axis, w.grid_spac, 6),
  lambda: orc.idw_esrg_nearest(pf, w.data_pts, w.x_axis, w.y_axis, w.grid_spac, 6)),
 ("idw_kdtree k=6", lambda: ip.idw_kdtree(w.points.T, w.data_pts, w.x_axis, w.y_axis, 6),
  lambda: orc.idw_kdtree(w.points.T, w.data_pts, w.x_axis, w.y_axis, 6)),
 ("bilinear", lambda: ip.bilinear(w.x_axis, w.y_axis, field, pf), lambda: orc.bilinear(w.x_axis, w.y_axis, field, pf)),
]
print(f"C1: {w.points.shape[0]} points, {w.simplices.shape[0]} triangles, grid {w.x_axis.size} x {w.y_axis.size}; oracle threads {orc.get_max_threads()}")
for name, g, c in rows:
    print(f"{name:28s} GPU (numpy in/out) {t(g):8.3f} ms/call   CPU oracle {t(c, 10):8.3f} ms/call")
