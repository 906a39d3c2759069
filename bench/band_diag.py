"""Diagnostic: one band of C4 with the band-limited pack against the same rows of the one-shot run."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from grid_interpolation_b200 import interpolation as ip, _lib
import workloads
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
grid = int(16384 * scale) // 128 * 128
w = workloads.c4(int(5_000_000 * scale ** 2), grid)
t = lambda a: torch.as_tensor(np.ascontiguousarray(a), device="cuda:0")
mesh = [t(a) for a in (w.points, w.data_pts, w.x_axis, w.y_axis, w.simplices, w.neighbours)]
full, gf = ip.barycentric_interpolation_ex(*mesh, order="C", debug=True)
for b in (4, 7):
    rows = (b * grid // 8, (b + 1) * grid // 8)
    for bp in (False, True):
        out, g = ip.barycentric_interpolation_ex(*mesh, rows=rows, halo=8, order="C", debug=True, band_pack=bp, sync=False)
        torch.cuda.synchronize()
        st = _lib.lib().gi_stream_status(torch.cuda.current_stream().cuda_stream)
        print("band", b, "band_pack", bp, "status", st, flush=True)
        sl = slice(rows[0], rows[1])
        dv = (out != full[sl]); dm = (g.mask != gf.mask[sl]); dt = (g.tri != gf.tri[sl]) & ~gf.mask[sl]
        print("  value diffs", int(dv.sum()), "mask diffs", int(dm.sum()), "tri diffs (inside)", int(dt.sum()),
              "masked", int(g.mask.sum()), int(gf.mask[sl].sum()), "fill level", g.fill_level_max, "iters", g.iters_tot)
        if dm.any():
            idx = dm.nonzero()[:12].cpu().numpy()
            print("  first mask diffs (row, col):", [(int(r) + rows[0], int(c)) for r, c in idx])
            print("  mask-diff cols histogram:", np.histogram(dm.nonzero()[:, 1].cpu().numpy(), bins=8, range=(0, grid))[0])
            print("  mask-diff rows histogram:", np.histogram(dm.nonzero()[:, 0].cpu().numpy(), bins=8, range=(0, rows[1]-rows[0]))[0])
        if dv.any():
            idx = dv.nonzero()[:10].cpu().numpy()
            print("  first value diffs (row, col):", [(int(r) + rows[0], int(c)) for r, c in idx])
            print("  cols histogram:", np.histogram(dv.nonzero()[:, 1].cpu().numpy(), bins=8, range=(0, grid))[0])
