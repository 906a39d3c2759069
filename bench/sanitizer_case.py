import sys; sys.path.insert(0, '.')
import numpy as np
from grid_interpolation_b200 import interpolation as ip, workloads
w = workloads.c1(num_points=800)
a, d = ip.barycentric_interpolation_ex(w.points, w.data_pts, w.x_axis, w.y_axis, w.simplices, w.neighbours, debug=True)
b = ip.bilinear(w.x_axis, w.y_axis, a, w.points)
e = ip.idw_esrg_nearest(w.points, w.data_pts, w.x_axis, w.y_axis, w.grid_spac, 6)
e2 = ip.idw_esrg_nearest_ex(w.points, w.data_pts, w.x_axis, w.y_axis, w.grid_spac, 6, reference_order=True)
k = ip.idw_kdtree(w.points.T, w.data_pts, w.x_axis, w.y_axis, 6)
parts = [ip.barycentric_interpolation_ex(w.points, w.data_pts, w.x_axis, w.y_axis, w.simplices, w.neighbours, rows=r, halo=4, order="C") for r in ((0, 11), (11, w.y_axis.size))]
gx, gy = np.meshgrid(np.arange(0.0, 7.0), np.arange(0.0, 5.0)); pts = np.stack([gx.ravel(), gy.ravel()], 1)
s, n = workloads.delaunay(pts)
f = ip.barycentric_interpolation(pts, pts[:, 0] * 2.0, np.arange(-0.5, 7.0, 0.5), np.arange(-0.5, 5.0, 0.25), s, n)
print("ok", float(a.sum()), float(b.sum()), float(e.sum()), float(k.sum()), float(np.concatenate(parts).sum()), float(f.sum()), np.array_equal(e, e2))
