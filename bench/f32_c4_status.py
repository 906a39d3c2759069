import numpy as np, torch
from grid_interpolation_b200 import interpolation as ip, workloads
dev = torch.device("cuda:0")
w = workloads.c4(); g = 16384
t = lambda a, d=torch.float32: torch.as_tensor(np.ascontiguousarray(a), device=dev).to(d)
a = (t(w.points), t(w.data_pts), t(w.x_axis), t(w.y_axis), t(w.simplices, torch.int32), t(w.neighbours, torch.int32))
out = torch.empty((g, g), dtype=torch.float32, device=dev)
try:
    ip.barycentric_interpolation_ex(*a, out=out, dtype=np.float32)
    print("f32 full C4: ok")
except RuntimeError as e:
    print("f32 full C4 raised:", e)
for r0 in range(0, g, 2048):
    try:
        res, d = ip.barycentric_interpolation_ex(*a, dtype=np.float32, debug=True, rows=(r0, r0 + 2048), halo=8)
        print(r0, "ok", {k: v for k, v in vars(d).items() if not hasattr(v, "shape")})
    except RuntimeError as e:
        print(r0, "raised:", e)
