"""One row band of C4 (what one rank of the strong-scaling run executes) for an ncu capture:
`ncu -k regex:locate_eval ... python bench/ncu_band.py <band> <nbands>`"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from grid_interpolation_b200 import interpolation as ip
import workloads
b, nb = int(sys.argv[1]), int(sys.argv[2])
grid = 16384
w = workloads.c4()
t = lambda a: torch.as_tensor(np.ascontiguousarray(a), device="cuda:0")
mesh = [t(a) for a in (w.points, w.data_pts, w.x_axis, w.y_axis, w.simplices, w.neighbours)]
rows = (b * grid // nb, (b + 1) * grid // nb)
out = torch.empty((rows[1] - rows[0], grid), dtype=torch.float64, device="cuda:0")
for _ in range(3):
    ip.barycentric_interpolation_ex(*mesh, rows=rows, halo=8, out=out, band_pack=True, sync=False)
torch.cuda.synchronize()
print("done")
