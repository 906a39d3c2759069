"""Three single-precision C4 barycentric calls (for an ncu launch list of the f32 path)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from grid_interpolation_b200 import interpolation as ip
import workloads
w = workloads.c4(); g = 16384
t = lambda a, d=torch.float32: torch.as_tensor(np.ascontiguousarray(a), device="cuda:0").to(d)
a = (t(w.points), t(w.data_pts), t(w.x_axis), t(w.y_axis), t(w.simplices, torch.int32), t(w.neighbours, torch.int32))
out = torch.empty((g, g), dtype=torch.float32, device="cuda:0")
for _ in range(3):
    ip.barycentric_interpolation_ex(*a, out=out, sync=False, dtype=np.float32)
torch.cuda.synchronize()
print("done")
