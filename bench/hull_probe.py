"""Walk iterations per target for single-row bands of C4 (each target then walks from the seed grid): how long
are the walks of the rows outside / near the hull?"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from grid_interpolation_b200 import interpolation as ip
import workloads
w = workloads.c4()
t = lambda a: torch.as_tensor(np.ascontiguousarray(a), device="cuda:0")
mesh = [t(a) for a in (w.points, w.data_pts, w.x_axis, w.y_axis, w.simplices, w.neighbours)]
print("min/max y of points", w.points[:, 1].min(), w.points[:, 1].max(), "x", w.points[:, 0].min(), w.points[:, 0].max())
for r in (0, 1, 2, 3, 5, 8, 16, 31, 32, 8192, 16380, 16382, 16383):
    out, g = ip.barycentric_interpolation_ex(*mesh, rows=(r, r + 1), halo=0, order="C", debug=True, sync=False)
    torch.cuda.synchronize()
    from grid_interpolation_b200 import _lib
    _lib.lib().gi_stream_status(torch.cuda.current_stream().cuda_stream)     # (a 1-row band cannot fill: ignore)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    o2 = torch.empty_like(out)
    e0.record()
    ip.barycentric_interpolation_ex(*mesh, rows=(r, r + 1), halo=0, out=o2, sync=False)
    e1.record(); torch.cuda.synchronize()
    _lib.lib().gi_stream_status(torch.cuda.current_stream().cuda_stream)
    print(f"row {r}: iterations/target {g.iters_tot / 16384:.1f}  masked {int(g.mask.sum())}  call {e0.elapsed_time(e1):.3f} ms", flush=True)
