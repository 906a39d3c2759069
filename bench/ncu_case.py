"""One C4 barycentric call per kernel variant for an ncu capture: `ncu -k regex:locate_eval ... python bench/ncu_case.py [scale] [row_tests] [plane]`"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from grid_interpolation_b200 import interpolation as ip
import workloads
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
row_tests = len(sys.argv) > 2 and sys.argv[2] == "1"
plane = len(sys.argv) > 3 and sys.argv[3] == "1"
grid = int(16384 * scale) // 128 * 128
w = workloads.c4(int(5_000_000 * scale ** 2), grid)
t = lambda a: torch.as_tensor(np.ascontiguousarray(a), device="cuda:0")
mesh = [t(a) for a in (w.points, w.data_pts, w.x_axis, w.y_axis, w.simplices, w.neighbours)]
out = torch.empty((grid, grid), dtype=torch.float64, device="cuda:0")
for _ in range(3):
    ip.barycentric_interpolation_ex(*mesh, out=out, plane_values=plane, sync=False)
torch.cuda.synchronize()
print("done")
