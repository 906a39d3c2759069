"""What would a spatially ordered mesh buy?  C4 with (a) the mesh as Qhull delivers it, (b) triangles re-ordered along
the Z curve of their centroids, (c) triangles and points re-ordered (ids relabelled; vertex order inside a triangle
kept, so every value keeps its bits).  Host-side numpy re-ordering: a probe for the GPU ingest pass, not the pass."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from grid_interpolation_b200 import interpolation as ip
import workloads

def zkey(x, y, cell):
    def spread(v):
        v = v.astype(np.uint64)
        v = (v | (v << 16)) & 0x0000ffff0000ffff; v = (v | (v << 8)) & 0x00ff00ff00ff00ff
        v = (v | (v << 4)) & 0x0f0f0f0f0f0f0f0f; v = (v | (v << 2)) & 0x3333333333333333
        v = (v | (v << 1)) & 0x5555555555555555
        return v
    return spread((x / cell).astype(np.int64)) | (spread((y / cell).astype(np.int64)) << 1)

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
grid = int(16384 * scale) // 128 * 128
w = workloads.c4(int(5_000_000 * scale ** 2), grid)
dev = torch.device("cuda:0")
t = lambda a: torch.as_tensor(np.ascontiguousarray(a), device=dev)

def run(tag, points, data, simp, nbr, ref=None):
    args = (t(points), t(data), t(w.x_axis), t(w.y_axis), t(simp), t(nbr))
    out = torch.empty((grid, grid), dtype=torch.float64, device=dev)
    for _ in range(3):
        ip.barycentric_interpolation_ex(*args, out=out, sync=False)
    torch.cuda.synchronize()
    ip.set_profiling(True)
    for _ in range(8):
        ip.barycentric_interpolation_ex(*args, out=out, sync=False)
    torch.cuda.synchronize()
    ph = {}
    for name, ms in ip.last_timings():
        ph.setdefault(name, []).append(ms)
    ip.set_profiling(False)
    same = "" if ref is None else f"  identical={bool(torch.equal(ref, out))}"
    print(f"{tag:34s} " + "  ".join(f"{k} {np.mean(v):.3f} ms" for k, v in ph.items()) + same, flush=True)
    return out

ref = run("as delivered (Qhull order)", w.points, w.data_pts, w.simplices, w.neighbours)
# (b) triangles along the Z curve of their centroids
cx = w.points[w.simplices - 1, 0].mean(1); cy = w.points[w.simplices - 1, 1].mean(1)
tperm = np.argsort(zkey(cx, cy, 4.0), kind="stable")
tinv = np.empty_like(tperm); tinv[tperm] = np.arange(tperm.size)
simp_b = w.simplices[tperm]
nb = w.neighbours[tperm]
nbr_b = np.where(nb > 0, tinv[np.maximum(nb, 1) - 1] + 1, 0).astype(np.int32)
run("triangles Z-ordered", w.points, w.data_pts, simp_b, nbr_b, ref)
# (c) points Z-ordered as well
pperm = np.argsort(zkey(w.points[:, 0], w.points[:, 1], 4.0), kind="stable")
pinv = np.empty_like(pperm); pinv[pperm] = np.arange(pperm.size)
simp_c = (pinv[simp_b - 1] + 1).astype(np.int32)
run("triangles and points Z-ordered", w.points[pperm], w.data_pts[pperm], simp_c, nbr_b, ref)
