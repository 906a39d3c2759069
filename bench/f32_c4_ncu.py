import numpy as np, torch
from grid_interpolation_b200 import interpolation as ip, workloads
dev = torch.device("cuda:0")
w = workloads.c4(1_250_000, 8192); g = 8192
R, dt = np.float32, torch.float32
t = lambda a, d=dt: torch.as_tensor(np.ascontiguousarray(a), device=dev).to(d)
a = (t(w.points), t(w.data_pts), t(w.x_axis), t(w.y_axis), t(w.simplices, torch.int32), t(w.neighbours, torch.int32))
out = torch.empty((g, g), dtype=dt, device=dev)
for _ in range(3): ip.barycentric_interpolation_ex(*a, out=out, sync=False, dtype=R)
torch.cuda.synchronize()
